#!/bin/bash
# final one-GPU session of the round: smoke, GPU tests, bench lines, ncu evidence (reps stay in /tmp: gpurun_out/ <= 64 MiB)
OUT=gpurun_out/r2x
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
python tools/variant_step_times.py > $OUT/variant_step_times.txt 2>&1
MBT_JIT=0 python tools/variant_step_times.py > $OUT/variant_step_times_nojit.txt 2>&1
python examples/cuda_graph_episode.py > $OUT/cuda_graph_episode.txt 2>&1; tail -4 $OUT/cuda_graph_episode.txt
# ncu first (so that the bench lines below can read the fresh summaries from profiles/ -- copied there by hand afterwards)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_f64.csv \
    python bench.py --steps 20 --warmup 3 --reps 2 --no-extras --no-cpu-baseline --e2e-steps 3 > $OUT/bench_under_ncu.json 2> $OUT/bench_under_ncu.stderr
for p in f64 f32; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mbt_(rollout|fill_batch|step|jit)' -o /tmp/targets_$p -f \
      python tools/profile_targets.py $p > $OUT/targets_$p.log 2>&1
  ncu -i /tmp/targets_$p.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summarise.py > $OUT/r2_targets_$p.ncu_summary.csv
  # steady state: launches 60.. of the running device-resident step loop, caches NOT flushed between launches
  timeout 600 ncu --cache-control none --clock-control none \
      --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,launch__registers_per_thread \
      -k regex:'mbt_step_kernel' --launch-skip 60 --launch-count 8 -o /tmp/steady_$p -f \
      python bench.py --precision $p --steps 100 --warmup 3 --reps 1 --no-extras --no-cpu-baseline --no-episode-stats --e2e-steps 3 > /dev/null 2> $OUT/steady_$p.stderr
  ncu -i /tmp/steady_$p.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summarise.py > $OUT/r2_step_steady_$p.ncu_summary.csv
done
cp $OUT/r2_*.ncu_summary.csv profiles/ 2>/dev/null
timeout 600 python bench.py > $OUT/bench_f64.json 2> $OUT/bench_f64.stderr; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_f64_k20.json 2> $OUT/bench_f64_k20.stderr; echo "bench k20 rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference_arm.json 2>/dev/null
for w in cjmm hawkes oe; do timeout 600 python bench.py --workload $w --no-cpu-baseline --no-extras > $OUT/bench_${w}_f64.json 2>/dev/null; done
timeout 600 python bench.py --precision f32 --no-cpu-baseline --no-extras > $OUT/bench_as_f32.json 2>/dev/null
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if d.get("impl")=="reference": print(f, d["value"]); continue
        print(f, round(d["ms_per_step"]*1e3,2), "us frac", round(d["roofline"]["frac"],3), "traffic", d["roofline"].get("traffic"), "spread", round(d["window"]["spread"],3), "e2e", round(d["e2e"]["ms_per_step"],3), round(d["e2e"]["frac"],3), "rollout", d["episode_stats"] and round(d["episode_stats"]["fused_rollout_ms"],3), d["episode_stats"] and d["episode_stats"].get("roofline") and round(d["episode_stats"]["roofline"]["frac"],3))
    except Exception as e: print(f, "ERR", e)
PY
ls -la $OUT | head -40
