#!/bin/bash
# ncu captures (the .ncu-rep files stay in /tmp on the box: gpurun_out/ is limited to 64 MiB; only summaries come back)
OUT=gpurun_out/r2d
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_f64.csv \
    python bench.py --steps 20 --warmup 3 --reps 2 --no-extras --no-cpu-baseline --e2e-steps 3 > $OUT/bench_under_ncu.json 2> $OUT/bench_under_ncu.stderr
for p in f64 f32; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mbt_(rollout|fill_batch|step|jit)' -o /tmp/targets_$p -f \
      python tools/profile_targets.py $p > $OUT/targets_$p.log 2>&1
  ncu -i /tmp/targets_$p.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summarise.py > $OUT/targets_$p.ncu_summary.csv
done
# source-level hot spots of the two f64 step kernels (plain variant 1 = launch id 0; normalised variant 10 = second to last pair)
ncu -i /tmp/targets_f64.ncu-rep --page source --csv --print-source sass --kernel-name regex:'mbt_step_kernel' --launch-skip 0 --launch-count 1 > $OUT/step_v1_f64.source.csv 2>/dev/null
ls -la $OUT /tmp/targets_*.ncu-rep
