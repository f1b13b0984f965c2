import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, bench
f = bench.make_env("as", "float64", 1 << 20, 0, 0)
a = f.pinned_actions(); a[:] = 0.7
f.reset()
for _ in range(6):
    o, r, d, i = f.step(a)
ts=[]
for _ in range(6):
    t0=time.perf_counter(); o, r, d, i = f.step(a); ts.append(time.perf_counter()-t0)
print("facade step ms:", [round(1e3*t,3) for t in ts], file=sys.stderr)
