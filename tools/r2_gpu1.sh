#!/bin/bash
# one-GPU session: GPU test suite, bench line, (optional) ncu captures.  usage: tools/r2_gpu1.sh [tests|bench|ncu|all] [tag]
WHAT=${1:-all}; TAG=${2:-a}
OUT=gpurun_out/r2$TAG
mkdir -p $OUT
if [ $WHAT = tests ] || [ $WHAT = all ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
if [ $WHAT = bench ] || [ $WHAT = all ]; then
  timeout 600 python bench.py > $OUT/bench_f64.json 2> $OUT/bench_f64.stderr; echo "bench rc=$?"
  timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench_f64_k20.json 2> $OUT/bench_f64_k20.stderr; echo "bench k20 rc=$?"
  tail -c 600 $OUT/bench_f64.stderr
fi
if [ $WHAT = ncu ] || [ $WHAT = all ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_f64.csv \
      python bench.py --steps 20 --warmup 3 --reps 1 --no-extras --no-cpu-baseline --e2e-steps 3 > $OUT/bench_under_ncu.json 2> $OUT/bench_under_ncu.stderr
  for p in f64 f32; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mbt_(rollout|fill_batch|step)' -o $OUT/targets_$p -f \
        python tools/profile_targets.py $p > $OUT/targets_$p.log 2>&1
    ncu -i $OUT/targets_$p.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summarise.py > $OUT/targets_$p.ncu_summary.csv
  done
  ls -la $OUT
fi
