#!/bin/bash
# usage: tools/r2_probe.sh N   -- host-path probes on N GPUs (writes gpurun_out/r2_probe_N/*)
N=${1:-2}
OUT=gpurun_out/r2_probe_$N
mkdir -p $OUT
{ lscpu | head -30; echo; nvidia-smi topo -m; echo; nvidia-smi -q | grep -A4 "GPU Link Info" | head -40; echo; cat /sys/devices/system/node/node*/cpulist; free -g; } > $OUT/host.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR tools/pcie_probe_multi.py --out $OUT/pcie_probe.json > $OUT/pcie_probe.stdout 2> $OUT/pcie_probe.stderr
for path in dma zerocopy; do
  for chunks in 4 1; do
    [ $path = zerocopy ] && [ $chunks = 1 ] && continue
    MBT_HOST_PATH=$path MBT_PIPE_CHUNKS=$chunks $TR bench.py --gpus $N --steps 20 --warmup 5 --no-episode-stats --no-cpu-baseline \
      > $OUT/bench_${path}_c${chunks}.json 2> $OUT/bench_${path}_c${chunks}.stderr
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["e2e"])
    except Exception as e: print(f, "ERR", e)
PY
