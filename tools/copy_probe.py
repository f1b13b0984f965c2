#!/usr/bin/env python
"""How fast is a plain device copy of the SAME number of bytes the step kernel moves?  (context for roofline.frac at
N = 2^20: MEASURED_PEAKS.json's 6.57 TB/s is a 2 GiB copy; a 109 MB kernel pays a fixed launch/ramp/drain cost.)"""
import torch

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
for label, nbytes in (("f32 step @2^20 (54.5 MB moved)", 54_525_952), ("f64 step @2^20 (109 MB moved)", 109_051_904),
                      ("f64 step @2^22 (436 MB)", 436_207_616), ("f64 step @2^24 (1.74 GB)", 1_744_830_464)):
    half = nbytes // 2
    # rotate over enough buffers to exceed L2 between reuses, like bench.py does
    nbuf = max(2, int(2 * 126e6 // half) + 1)
    src = [torch.empty(half, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    dst = [torch.empty(half, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    for i in range(nbuf):
        dst[i].copy_(src[i])
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
    for k, (a, b) in enumerate(evs):
        a.record(stream)
        dst[k % nbuf].copy_(src[k % nbuf])
        b.record(stream)
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    med = ms[len(ms) // 2]
    print(f"{label:36s} copy kernel median {med*1e3:8.2f} us  -> {nbytes/med/1e6:8.1f} GB/s")
# empty-kernel floor with the same event bracketing
x = torch.zeros(1, device=dev)
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
for a, b in evs:
    a.record(stream); x.add_(1); b.record(stream)
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in evs)
print(f"1-element kernel between two events: median {ms[50]*1e3:.2f} us")
