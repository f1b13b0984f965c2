#!/bin/bash
OUT=gpurun_out/r2b
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
python tools/variant_step_times.py > $OUT/variant_step_times.txt 2>&1; cat $OUT/variant_step_times.txt
MBT_JIT=0 python tools/variant_step_times.py > $OUT/variant_step_times_nojit.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_f64.json 2> $OUT/bench_f64.stderr; echo "bench rc=$?"
MBT_L2_PERSIST=1 timeout 600 python bench.py --no-cpu-baseline --no-extras --no-episode-stats > $OUT/bench_f64_l2persist.json 2> $OUT/bench_f64_l2persist.stderr; echo "bench l2 rc=$?"
python - <<PY
import json
for f in ("$OUT/bench_f64.json","$OUT/bench_f64_l2persist.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["window"]["spread"], d["e2e"]["ms_per_step"], d["e2e"]["frac"])
PY
