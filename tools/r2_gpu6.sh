#!/bin/bash
OUT=gpurun_out/r2f
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
python tools/variant_step_times.py > $OUT/variant_step_times.txt 2>&1; tail -4 $OUT/variant_step_times.txt
for b in 1 2 8; do echo "fill blocks per SM = $b"; MBT_FILL_BLOCKS_PER_SM=$b python tools/variant_step_times.py 2>&1 | tail -2; done
timeout 600 python bench.py --no-cpu-baseline --no-extras > $OUT/bench_f64.json 2> $OUT/bench_f64.stderr; echo "bench rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"]*1e3,2), "us  frac", round(d["roofline"]["frac"],3), "rollout", round(d["episode_stats"]["fused_rollout_ms"],3))
PY
