#!/usr/bin/env python
"""bench.py -- env-steps/s of the hot path (TradingEnvironment.step) on N B200s of one node.

    python bench.py --gpus 1 --steps 200 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (NumPy port, all host cores)

A "step" is one env-step of the whole batch.  Workload at N=1: BASELINE.json configs[1] -- Avellaneda-Stoikov
(Brownian midprice sigma=2, Poisson arrivals lambda=140, exponential fills kappa=1.5, PnL reward, T=1, n_steps=200),
num_trajectories = 2^20 per GPU (weak scaling: GPU g owns global trajectory ids [g*2^20, (g+1)*2^20)).

One JSON line on stdout (rank 0):
  value      whole-job env-steps/s, inputs resident in HBM (actions/obs/rewards are device buffers), K steps timed
             with CUDA events between barriers, max over ranks.
  e2e        the same metric through the public API with HOST buffers: every step copies the (N,A) action array from
             pinned host memory to the device and the (N,D) observations + (N,) rewards back.
  roofline   HBM roofline of the step kernel: algorithmic bytes per launch / mean launch duration (CUDA events around
             every launch inside the timed region) vs the measured copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the NumPy port of the reference step() (oracle/numpy_port.py) on this box's host cores (rank 0, N=1).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_PER_GPU = 1 << 20
L2_BYTES = 126e6
WORKLOADS = ("as", "cjmm", "hawkes", "oe")


def make_config(workload, precision, n_local, traj_offset):
    """BASELINE.json configs[1..4] as mbt_config (SURVEY.md 8d synthetic inputs)."""
    from mbt_gym_b200 import _abi

    n_steps, T = 200, 1.0
    dt = T / n_steps
    common = dict(precision=precision, num_trajectories=n_local, traj_offset=traj_offset, n_steps=n_steps,
                  terminal_time=T, step_size=dt, rew_terminal_time=T, mid_initial=100.0, mid_vol=2.0, mid_step=dt)
    s_max = 100.0 + 4 * 2.0 * np.sqrt(T)
    if workload in ("as", "cjmm", "hawkes"):
        cfg = _abi.new_config(dynamics=_abi.MBT_DYN_LIMIT, midprice=_abi.MBT_MID_BM, fill=_abi.MBT_FILL_EXPONENTIAL,
                              fill_exponent=1.5, arr_step=dt, max_inventory=200.0, max_cash=n_steps * s_max, **common)
        if workload == "hawkes":
            cfg.arrival = _abi.MBT_ARR_HAWKES
            cfg.arr_rate[0] = cfg.arr_rate[1] = 10.0
            cfg.hawkes_jump, cfg.hawkes_speed = 40.0, 60.0
        else:
            cfg.arrival = _abi.MBT_ARR_POISSON
            cfg.arr_rate[0] = cfg.arr_rate[1] = 140.0
        if workload == "cjmm":
            cfg.reward = _abi.MBT_REW_CJ_MM
            cfg.rew_phi, cfg.rew_alpha, cfg.max_inventory = 0.01, 0.001, 100.0
        else:
            cfg.reward = _abi.MBT_REW_PNL
    elif workload == "oe":
        cfg = _abi.new_config(dynamics=_abi.MBT_DYN_SPEED, midprice=_abi.MBT_MID_OU, impact=_abi.MBT_IMP_TEMP_PERM,
                              ou_level=100.0, ou_speed=1.0, imp_temp=0.01, imp_perm=0.01, imp_step=dt,
                              reward=_abi.MBT_REW_CJ_OE, rew_phi=0.01, rew_alpha=0.001, q0_const=100.0,
                              max_inventory=10_000.0, max_cash=n_steps * (100.0 + 4 * 2.0 * T), **common)
    else:
        raise ValueError(workload)
    return cfg


def algorithmic_bytes_per_env_step(A, D, esz):
    """B = w * (A + 2*S + O + 1), S = D-1 persistent state scalars read+written, O = D obs scalars."""
    return esz * (A + 2 * (D - 1) + D + 1)


def fixed_action_value(workload):
    return -1.0 if workload == "oe" else 0.7


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_baseline(workload, seconds_target=15.0):
    """NumPy port of the reference step(), one env per host core (fork), bounded sample."""
    from oracle import numpy_port as P

    cores = os.cpu_count() or 1
    per_core_rate = 5e6  # env-steps/s/core, order of magnitude, only used to size the sample
    steps = 20
    n = int(min(N_PER_GPU, max(cores * 4096, per_core_rate * cores * seconds_target / steps)))
    t, n_run = P.time_port(workload, n, steps, 2, cores)
    t1, n1 = P.time_port(workload, min(n, 1 << 18), steps, 2, 1)
    return {"value": n_run * steps / t, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"{steps} steps x {n_run} trajectories, one NumPy env per core ({cores} procs, fork); "
                      f"single process: {n1 * steps / t1:.3e} env-steps/s",
            "single_process_value": n1 * steps / t1}


def run_reference_arm(args):
    """The reference's own CPU implementation of the path (NumPy port; /root/reference is absent on the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import numpy_port as P

    cores = os.cpu_count() or 1
    n = N_PER_GPU * args.gpus
    # bounded sample: at most ~2 minutes of wall clock for steps+warmup at ~5e6 env-steps/s/core
    budget = 120.0 * 5e6 * cores
    n_run = int(min(n, max(cores * 1024, budget / max(1, args.steps + args.warmup))))
    t, n_done = P.time_port(args.workload, n_run, args.steps, args.warmup, cores)
    value = n_done * args.steps / t
    sample = f"{args.steps} steps x {n_done} trajectories of the {n}-trajectory workload, one NumPy env per core ({cores} procs)"
    line = {"impl": "reference", "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "num_trajectories": n, "n_steps": 200},
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_name(w):
    return {"as": "Avellaneda-Stoikov market making (BM midprice, Poisson arrivals, exponential fills, PnL), BASELINE configs[1]",
            "cjmm": "Cartea-Jaimungal-Penalva 2015 (CjMmCriterion), BASELINE configs[2]",
            "hawkes": "Hawkes arrivals (BM midprice, exponential fills, PnL), BASELINE configs[3]",
            "oe": "optimal execution (speed dynamics, OU midprice, temporary+permanent impact, CjOeCriterion), BASELINE configs[4]"}[w]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="as", choices=WORKLOADS)
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=None)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    from mbt_gym_b200 import _abi, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mbt_gym_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    precision = _abi.MBT_F64 if args.precision == "f64" else _abi.MBT_F32
    tdt = torch.float64 if args.precision == "f64" else torch.float32
    esz = 8 if args.precision == "f64" else 4
    cfg = make_config(args.workload, precision, N_PER_GPU, rank * N_PER_GPU)
    env = _lib.NativeEnv(cfg, device=local_rank)
    N, A, D = env.N, env.A, env.D
    env.seed(1234)
    stream = torch.cuda.Stream()  # a real (non-default) stream: the env's kernels and the timing events share it
    torch.cuda.set_stream(stream)
    env.set_stream(stream.cuda_stream)

    # rotating buffer sets so that the bytes touched between two uses of a buffer exceed L2 (126 MB)
    set_bytes = N * (A + D + 1) * esz
    n_sets = int(np.ceil(2 * L2_BYTES / set_bytes)) + 1
    acts = [torch.full((N, A), fixed_action_value(args.workload), dtype=tdt, device="cuda") for _ in range(n_sets)]
    obs = [torch.empty((N, D), dtype=tdt, device="cuda") for _ in range(n_sets)]
    rew = [torch.empty((N,), dtype=tdt, device="cuda") for _ in range(n_sets)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(k_steps, timing):
        """k_steps env-steps, auto-resetting at episode end like the SB3 VecEnv adapter does."""
        env.enable_timing(timing)
        for k in range(k_steps):
            i = k % n_sets
            done = env.step(acts[i], obs[i], rew[i], mem=_abi.MBT_MEM_DEVICE)
            if done:
                env.reset(obs[i], mem=_abi.MBT_MEM_DEVICE)

    env.reset(obs[0], mem=_abi.MBT_MEM_DEVICE)
    run_steps(args.warmup, False)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    launches0 = env.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    run_steps(args.steps, True)
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = env.launch_count() - launches0
    ktimes = env.kernel_times_ms()
    env.enable_timing(False)

    env.set_stream(None)  # back to the handle's own stream for the host-buffer path
    # ---- e2e: host buffers through the call a user makes (pinned action array in, pinned obs/rew out)
    e2e_steps = args.e2e_steps or max(10, min(args.steps, 50))
    h_act = _lib.PinnedArray((N, A), env.dtype)
    h_obs = _lib.PinnedArray((N, D), env.dtype)
    h_rew = _lib.PinnedArray((N,), env.dtype)
    h_act.array[:] = fixed_action_value(args.workload)
    env.reset(h_obs.array)
    for _ in range(3):
        if env.step(h_act.array, h_obs.array, h_rew.array):
            env.reset(h_obs.array)
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for _ in range(e2e_steps):
        if env.step(h_act.array, h_obs.array, h_rew.array):
            env.reset(h_obs.array)
        checksum += float(h_rew.array[0])  # the step's result is read on the host
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None

    t = torch.tensor([elapsed_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        total_n = N * world
        value = total_n * args.steps / (elapsed_ms * 1e-3)
        e2e_value = total_n * e2e_steps / (e2e_ms * 1e-3)
        b_step = algorithmic_bytes_per_env_step(A, D, esz)
        peak, peak_src = measured_hbm_peak()
        mean_kernel_ms = float(np.mean(ktimes)) if len(ktimes) else float("nan")
        achieved = N * b_step / (mean_kernel_ms * 1e-3) / 1e9
        line = {
            "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
            "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "num_trajectories": total_n,
                       "num_trajectories_per_gpu": N, "n_steps": 200, "action": fixed_action_value(args.workload),
                       "l2": f"{n_sets} rotating (action, obs, reward) buffer sets, {n_sets * set_bytes / 1e6:.0f} MB "
                             f"streamed between reuses > 126 MB L2; no explicit flush",
                       "parallelism": f"trajectory shards x{world}, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": N * A * esz,
                    "d2h_bytes_per_step": N * (D + 1) * esz, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "mbt_step_kernel",
                         "algorithmic_bytes_per_env_step": b_step, "mean_kernel_ms": mean_kernel_ms,
                         "kernel_launches_timed": int(len(ktimes)),
                         "kernel_share_of_step": mean_kernel_ms / (elapsed_ms / args.steps)},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload)
        print(json.dumps(line))
    env.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
