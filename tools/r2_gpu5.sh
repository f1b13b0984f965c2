#!/bin/bash
OUT=gpurun_out/r2e
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
python tools/variant_step_times.py > $OUT/variant_step_times.txt 2>&1; cat $OUT/variant_step_times.txt
timeout 600 python bench.py > $OUT/bench_f64.json 2> $OUT/bench_f64.stderr; echo "bench rc=$?"
for w in cjmm hawkes oe; do timeout 600 python bench.py --workload $w --no-cpu-baseline --no-extras > $OUT/bench_${w}_f64.json 2>/dev/null; done
timeout 600 python bench.py --precision f32 --no-cpu-baseline --no-extras > $OUT/bench_as_f32.json 2>/dev/null
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"]*1e3,2), "us  frac", round(d["roofline"]["frac"],3), "spread", round(d["window"]["spread"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "ms frac", round(d["e2e"]["frac"],3), "rollout", round(d["episode_stats"]["fused_rollout_ms"],3))
PY
