#!/bin/bash
OUT=gpurun_out/r2s
mkdir -p $OUT
python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "histogram or group_of_one or clip" 2>&1 | tail -3
SEL="test_batch_fill or test_fused_rollout_matches or test_clip_events or test_group_of_one or test_inventory_histogram or (test_f64_matches_reference_fixture and (power or triangular or gbm or hawkes_pnl or as_pnl_reward))"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -k "$SEL" > $OUT/r2_sanitizer_memcheck.txt 2>&1; tail -3 $OUT/r2_sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -k "$SEL" > $OUT/r2_sanitizer_racecheck.txt 2>&1; tail -3 $OUT/r2_sanitizer_racecheck.txt
