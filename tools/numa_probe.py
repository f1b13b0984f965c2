#!/usr/bin/env python
"""Is the slow H2D rate a NUMA effect?  Measure pinned H2D / D2H with the process bound to each NUMA node."""
import glob
import os
import subprocess
import sys
import time


def cpus_of(node):
    txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    out = []
    for part in txt.split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def child(node):
    os.sched_setaffinity(0, cpus_of(node))
    import torch

    N = 1 << 20
    d = torch.empty((N, 5), dtype=torch.float64, device="cuda")
    h = torch.empty((N, 5), dtype=torch.float64).pin_memory()
    h.fill_(1.0)
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(30):
            fn()
        torch.cuda.synchronize()
        t = (time.perf_counter() - t0) / 30
        print(f"  node {node}: {name} 41.9 MB {t*1e3:.3f} ms = {41.943/t/1e3:.1f} GB/s")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(int(sys.argv[1]))
    else:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
        for f in glob.glob("/sys/bus/pci/devices/*/numa_node"):
            cls = open(os.path.dirname(f) + "/class").read().strip()
            if cls.startswith("0x0302") or cls.startswith("0x0300"):
                print(f, open(f).read().strip())
        print("default affinity:", len(os.sched_getaffinity(0)), "cpus")
        nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
        for n in nodes:
            subprocess.run([sys.executable, __file__, str(n)])
