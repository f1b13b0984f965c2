#!/bin/bash
# two-GPU session: the 2-rank group test, bench at 2 GPUs with NCCL_DEBUG=INFO (like the driver may set it)
OUT=gpurun_out/r2g2
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_group.py -x -q > $OUT/pytest_group.log 2>&1; echo "group test rc=$?"; tail -5 $OUT/pytest_group.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=INFO timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.stderr; echo "bench rc=$?"
grep -c "nranks" $OUT/bench_2gpu.stderr; grep "nranks" $OUT/bench_2gpu.stderr | head -4
wc -l $OUT/bench_2gpu.json; tail -c 1500 $OUT/bench_2gpu.json
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $OUT/bench_2gpu_ref.json 2>/dev/null; echo "ref rc=$?"
