#!/bin/bash
OUT=gpurun_out/r2c
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
python tools/variant_step_times.py > $OUT/variant_step_times.txt 2>&1; cat $OUT/variant_step_times.txt
timeout 600 python bench.py > $OUT/bench_f64.json 2> $OUT/bench_f64.stderr; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_f64_k20.json 2> $OUT/bench_f64_k20.stderr; echo "bench k20 rc=$?"
MBT_L2_PERSIST=0 timeout 600 python bench.py --no-cpu-baseline --no-extras --no-episode-stats > $OUT/bench_f64_nol2.json 2> /dev/null
python - <<PY
import json
for f in ("$OUT/bench_f64.json","$OUT/bench_f64_k20.json","$OUT/bench_f64_nol2.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["roofline"]["frac"], d["window"]["spread"], d["e2e"]["ms_per_step"], d["e2e"]["frac"])
PY
# sanitizer over the kernels this round changed (batch-reduced fills, device-folded summaries, specialised kernels)
SEL="test_batch_fill or test_fused_rollout_matches or test_clip_events or test_group_of_one or (test_f64_matches_reference_fixture and (power or triangular or gbm or hawkes_pnl or as_pnl_reward))"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -k "$SEL" > $OUT/sanitizer_memcheck.txt 2>&1; tail -3 $OUT/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -k "$SEL" > $OUT/sanitizer_racecheck.txt 2>&1; tail -3 $OUT/sanitizer_racecheck.txt
# ncu: launch list of the bench command, then full captures of the target kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_f64.csv \
    python bench.py --steps 20 --warmup 3 --reps 2 --no-extras --no-cpu-baseline --e2e-steps 3 > $OUT/bench_under_ncu.json 2> $OUT/bench_under_ncu.stderr
for p in f64 f32; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mbt_(rollout|fill_batch|step|jit)' -o $OUT/targets_$p -f \
      python tools/profile_targets.py $p > $OUT/targets_$p.log 2>&1
  ncu -i $OUT/targets_$p.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summarise.py > $OUT/targets_$p.ncu_summary.csv
done
ls -la $OUT | head -30
