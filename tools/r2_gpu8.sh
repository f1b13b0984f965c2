#!/bin/bash
# eight-GPU session: host-link ceiling at 1/2/4/8 active GPUs, bench at 8 and 4 GPUs (NCCL_DEBUG=INFO like the driver may set it)
OUT=gpurun_out/r2g8b
mkdir -p $OUT
{ lscpu | head -20; nvidia-smi topo -m; free -g; } > $OUT/host.txt 2>&1
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542"
echo "probe skipped (profiles/r2_pcie_probe_8gpu.json)"

NCCL_DEBUG=INFO timeout 600 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_8gpu.json 2> $OUT/bench_8gpu.stderr; echo "bench8 rc=$?"
grep -c "nranks 8" $OUT/bench_8gpu.stderr; grep -v "NCCL INFO" $OUT/bench_8gpu.stderr | tail -5
timeout 600 $TR4 bench.py --gpus 4 --steps 20 --warmup 5 > $OUT/bench_4gpu.json 2> $OUT/bench_4gpu.stderr; echo "bench4 rc=$?"
timeout 300 $TR8 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > $OUT/bench_8gpu_ref.json 2>/dev/null; echo "ref rc=$?"
python - <<PY
import json
for f in ("$OUT/bench_8gpu.json","$OUT/bench_4gpu.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "us/step", d["ms_per_step"]*1e3, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["e2e"]["frac"], "ceil", d["e2e"]["pcie_peak_gbs"])
        print("   episode", {k:d["episode_stats"][k] for k in ("fused_rollout_ms","summary_collective_ms","episode_ms","returns_gathered")})
        print("   configs4", {k:d["configs4"][k] for k in ("fused_rollout_ms","summary_collective_ms","episode_ms","returns_gathered")})
        print("   f32io", d["e2e_float32_io"]["value"], d["e2e_float32_io"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
rm -f $OUT/bench_8gpu.stderr.tmp; head -c 200000 $OUT/bench_8gpu.stderr > $OUT/bench_8gpu.stderr.head; rm $OUT/bench_8gpu.stderr
