#!/usr/bin/env python
"""Host<->device copy ceiling of THIS box under the e2e path's traffic pattern, with 1..N GPUs active at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/pcie_probe_multi.py [--out gpurun_out/pcie_probe.json]

Every rank owns one GPU and moves, per "step", what one `TradingEnvironment.step(numpy)` of the BASELINE workload moves
(AS, N = 2^20, f64): 16.8 MB host->device (actions) and 41.9 MB device->host (observations + rewards), from / to pinned
host memory, with plain cudaMemcpyAsync on two streams -- no kernels, no library code.  For k = 1, 2, 4, ... active ranks
(the others idle at the barrier) it reports the aggregate GB/s of H2D alone, D2H alone and both directions at once:
the ceiling `bench.py`'s e2e number is a fraction of.  Variants: pinned pages allocated with / without a per-rank CPU
affinity (cores split evenly across ranks), and write-combined pages for the H2D source.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

N = 1 << 20
H2D_BYTES = N * 2 * 8
D2H_BYTES = N * 5 * 8

cudart = None


def rt():
    global cudart
    if cudart is None:
        cudart = C.CDLL("libcudart.so.12")
        cudart.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
        cudart.cudaFreeHost.argtypes = [C.c_void_p]
        cudart.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    return cudart


def host_alloc(nbytes, flags):
    p = C.c_void_p()
    rc = rt().cudaHostAlloc(C.byref(p), nbytes, flags)
    assert rc == 0, f"cudaHostAlloc rc={rc}"
    C.memset(p, 1, nbytes)  # first touch here, under the current affinity
    return p


def rank_cpus(rank, world):
    allowed = sorted(os.sched_getaffinity(0))
    per = max(1, len(allowed) // world)
    return set(allowed[rank * per:(rank + 1) * per]) or set(allowed)


def measure(rank, world, k_active, mode, h_src, h_dst, d_src, d_dst, s_in, s_out, reps):
    active = rank < k_active

    def enqueue():
        if mode in ("h2d", "duplex"):
            rt().cudaMemcpyAsync(d_dst.data_ptr(), h_src, H2D_BYTES, 1, s_in.cuda_stream)
        if mode in ("d2h", "duplex"):
            rt().cudaMemcpyAsync(h_dst, d_src.data_ptr(), D2H_BYTES, 2, s_out.cuda_stream)

    if active:
        enqueue()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if active:
        for _ in range(reps):  # a step waits for its own copies, like mbt_step does
            enqueue()
            s_in.synchronize()
            s_out.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt if active else 0.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per_step = float(t[0]) / reps
    nbytes = (H2D_BYTES if mode in ("h2d", "duplex") else 0) + (D2H_BYTES if mode in ("d2h", "duplex") else 0)
    return {"ms_per_step": 1e3 * per_step, "aggregate_gbs": k_active * nbytes / per_step / 1e9,
            "per_gpu_gbs": nbytes / per_step / 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--reps", type=int, default=40)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl" if world > 1 else "gloo", device_id=torch.device("cuda", local) if world > 1 else None,
                            init_method=None if "MASTER_ADDR" in os.environ else "tcp://127.0.0.1:29512",
                            rank=rank, world_size=world)
    d_dst = torch.empty(H2D_BYTES, dtype=torch.uint8, device="cuda")
    d_src = torch.ones(D2H_BYTES, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    results = []
    ks = [k for k in (1, 2, 4, 8) if k <= world]
    default_aff = os.sched_getaffinity(0)
    for variant in ("default", "affinity", "affinity+wc"):
        if "affinity" in variant:
            os.sched_setaffinity(0, rank_cpus(rank, world))
        else:
            os.sched_setaffinity(0, default_aff)
        h_src = host_alloc(H2D_BYTES, 4 if "wc" in variant else 0)  # cudaHostAllocWriteCombined = 4
        h_dst = host_alloc(D2H_BYTES, 0)
        for k in ks:
            for mode in ("h2d", "d2h", "duplex"):
                r = measure(rank, world, k, mode, h_src, h_dst, d_src, d_dst, s_in, s_out, args.reps)
                r.update(variant=variant, active_gpus=k, mode=mode)
                results.append(r)
                if rank == 0:
                    print(f"{variant:12s} k={k} {mode:6s} {r['ms_per_step']:.3f} ms/step  aggregate {r['aggregate_gbs']:.1f} GB/s"
                          f"  per GPU {r['per_gpu_gbs']:.1f} GB/s", file=sys.stderr, flush=True)
        rt().cudaFreeHost(h_src)
        rt().cudaFreeHost(h_dst)
    os.sched_setaffinity(0, default_aff)
    if rank == 0:
        info = {"world": world, "cpus_allowed": len(default_aff), "h2d_bytes": H2D_BYTES, "d2h_bytes": D2H_BYTES,
                "gpu": torch.cuda.get_device_name(0), "results": results}
        text = json.dumps(info, indent=1)
        if args.out:
            os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
            with open(args.out, "w") as f:
                f.write(text)
        print(json.dumps(info))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
