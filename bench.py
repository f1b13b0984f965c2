#!/usr/bin/env python
"""bench.py -- env-steps/s of the hot path (TradingEnvironment.step) on N B200s of one node.

    python bench.py --gpus 1 --steps 200 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (NumPy port, all host cores)

A "step" is one env-step of the whole batch.  Workload at N=1: BASELINE.json configs[1] -- Avellaneda-Stoikov
(Brownian midprice sigma=2, Poisson arrivals lambda=140, exponential fills kappa=1.5, PnL reward, T=1, n_steps=200),
num_trajectories = 2^20 per GPU (weak scaling: GPU g owns global trajectory ids [g*2^20, (g+1)*2^20)).

One JSON line on stdout (rank 0):
  value        whole-job env-steps/s, inputs resident in HBM (actions / obs / rewards are device buffers).  EXACTLY K steps
               are timed with CUDA events between barriers, max over ranks -- and that K-step window is repeated `reps`
               times (default 11): `value` / `ms_per_step` are the MEDIAN window, `window` holds min / max / spread.
  value_graph  the same steps captured once into a CUDA graph (reset + one whole episode) and replayed.
  e2e          the same metric through the public API with HOST buffers: every step copies the (N,A) action array from
               pinned host memory to the device and the (N,D) observations + (N,) rewards back.  `pcie_peak_gbs` is the
               rate bare cudaMemcpyAsync calls reach for the same bytes on the same ranks at the same time, `frac` the
               e2e path's share of it.
  roofline     HBM roofline of the step kernel: algorithmic bytes per launch / launch duration vs the measured copy
               bandwidth in MEASURED_PEAKS.json; `traffic` is read from the committed ncu summary under profiles/;
               `n_sweep` repeats the measurement at 2^20 / 2^22 / 2^24 trajectories (fixed cost vs per-trajectory cost).
  episode_stats  fused on-device rollout + the ONE collective of the path, inside the library (mbt_group_rollout: NCCL
               all-reduce of the device-resident summary on the env's stream, all-gather of returns overlapped).
  configs4     BASELINE.json configs[4]: optimal execution + OU midprice, 2^20 trajectories per GPU, fused rollout with
               the NCCL gather of all returns (8 388 608 of them on 8 GPUs).
  cpu_baseline the NumPy port of the reference step() (oracle/numpy_port.py) on this box's host cores (rank 0, N=1).
"""
import argparse
import csv
import glob
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
N_STEPS = 200
L2_BYTES = 126e6
WORKLOADS = ("as", "cjmm", "hawkes", "oe")
SM_COUNT = 148


def make_env(workload, precision, n_local, traj_offset, device, io_dtype=None):
    """BASELINE.json configs[1..4] through the PUBLIC API (SURVEY.md 8d synthetic inputs), normalisation off."""
    from mbt_gym_b200.gym.ModelDynamics import LimitOrderModelDynamics, TradinghWithSpeedModelDynamics
    from mbt_gym_b200.gym.TradingEnvironment import TradingEnvironment
    from mbt_gym_b200.rewards.RewardFunctions import CjMmCriterion, CjOeCriterion, PnL
    from mbt_gym_b200.stochastic_processes.arrival_models import HawkesArrivalModel, PoissonArrivalModel
    from mbt_gym_b200.stochastic_processes.fill_probability_models import ExponentialFillFunction
    from mbt_gym_b200.stochastic_processes.midprice_models import BrownianMotionMidpriceModel, OuMidpriceModel
    from mbt_gym_b200.stochastic_processes.price_impact_models import TemporaryAndPermanentPriceImpact

    n_steps, T, N = N_STEPS, 1.0, n_local
    dt = T / n_steps
    kw = dict(terminal_time=T, n_steps=n_steps, seed=1234, num_trajectories=N, normalise_action_space=False,
              normalise_observation_space=False, precision=precision, device=device, traj_offset=traj_offset,
              io_dtype=io_dtype)
    if workload in ("as", "cjmm", "hawkes"):
        mid = BrownianMotionMidpriceModel(volatility=2.0, initial_price=100.0, terminal_time=T, step_size=dt, num_trajectories=N)
        if workload == "hawkes":
            arr = HawkesArrivalModel(baseline_arrival_rate=np.array([[10.0, 10.0]]), step_size=dt, jump_size=40.0,
                                     mean_reversion_speed=60.0, terminal_time=T, num_trajectories=N)
        else:
            arr = PoissonArrivalModel(intensity=np.array([140.0, 140.0]), step_size=dt, num_trajectories=N)
        fill = ExponentialFillFunction(fill_exponent=1.5, step_size=dt, num_trajectories=N)
        dyn = LimitOrderModelDynamics(midprice_model=mid, arrival_model=arr, fill_probability_model=fill, num_trajectories=N)
        if workload == "cjmm":
            return TradingEnvironment(reward_function=CjMmCriterion(0.01, 0.001, 2.0, T), model_dynamics=dyn,
                                      max_inventory=100, **kw)
        return TradingEnvironment(reward_function=PnL(), model_dynamics=dyn, max_inventory=n_steps, **kw)
    if workload == "oe":
        mid = OuMidpriceModel(mean_reversion_level=100.0, mean_reversion_speed=1.0, volatility=2.0, initial_price=100.0,
                              terminal_time=T, step_size=dt, num_trajectories=N)
        imp = TemporaryAndPermanentPriceImpact(0.01, 0.01, n_steps=n_steps, terminal_time=T, num_trajectories=N)
        dyn = TradinghWithSpeedModelDynamics(midprice_model=mid, price_impact_model=imp, num_trajectories=N)
        return TradingEnvironment(reward_function=CjOeCriterion(0.01, 0.001, 2.0, T), model_dynamics=dyn,
                                  initial_inventory=100, max_inventory=10_000, **kw)
    raise ValueError(workload)


def algorithmic_bytes_per_env_step(A, D, esz):
    """B = w * (A + 2*S + O + 1), S = D-1 persistent state scalars read+written, O = D obs scalars."""
    return esz * (A + 2 * (D - 1) + D + 1)


def fixed_action_value(workload):
    return -1.0 if workload == "oe" else 0.7


def workload_name(w):
    return {"as": "Avellaneda-Stoikov market making (BM midprice, Poisson arrivals, exponential fills, PnL), BASELINE configs[1]",
            "cjmm": "Cartea-Jaimungal-Penalva 2015 (CjMmCriterion), BASELINE configs[2]",
            "hawkes": "Hawkes arrivals (BM midprice, exponential fills, PnL), BASELINE configs[3]",
            "oe": "optimal execution (speed dynamics, OU midprice, temporary+permanent impact, CjOeCriterion), BASELINE configs[4]"}[w]


def config_dict(workload, n_total, n_per_gpu):
    """The `config` object of the JSON line: the SAME keys and values in both arms (ours and --impl reference)."""
    return {"workload": workload_name(workload), "num_trajectories": int(n_total), "num_trajectories_per_gpu": int(n_per_gpu),
            "n_steps": N_STEPS, "action": fixed_action_value(workload),
            "l2": "inputs larger than L2: rotating (action, obs, reward) buffer sets, > 126 MB streamed between reuses",
            "parallelism": "trajectory shards, one per GPU; no data-path collective"}


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
        # "under load" = the upper half of the samples (the GPU idles at ~150 MHz between the legs of this script)
        busy = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm), "sm_mhz_all_samples_median": float(np.median(sm))}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", 1965.0


def ncu_profile_value(kernel_substrings, metrics, pattern="profiles/r2*.ncu_summary.csv"):
    """Per-launch values of `metrics` for the LAST captured launch whose kernel name contains all `kernel_substrings`, read
    from the committed ncu summaries (tools/ncu_summarise.py format: one row per metric, one column per launch).
    Returns (dict metric -> float in base units, file name) or (None, None)."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "": 1.0,
             "inst": 1.0, "%": 1.0}
    for path in sorted(glob.glob(os.path.join(ROOT, pattern)), reverse=True):
        try:
            with open(path) as f:
                rows = {r[0]: r for r in csv.reader(f) if r}
            names = rows["Kernel Name"][2:]
            cols = [i for i, n in enumerate(names) if all(s in n for s in kernel_substrings)]
            if not cols or any(m not in rows for m in metrics):
                continue
            c = cols[-1] + 2
            out = {m: float(rows[m][c].replace(",", "")) * scale.get(rows[m][1], 1.0) for m in metrics}
            return out, os.path.relpath(path, ROOT)
        except Exception:
            continue
    return None, None


def size_matched_copy_us(nbytes, stream, reps=60):
    """Duration of a plain device copy moving `nbytes` (half read, half written), back to back like the step loop, rotating
    over buffers that exceed L2: what a memcpy achieves at THIS size on THIS box."""
    import torch

    half = int(nbytes // 2)
    nbuf = max(2, int(2 * L2_BYTES // half) + 1)
    src = [torch.empty(half, dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
    dst = [torch.empty(half, dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
    for i in range(nbuf):
        dst[i].copy_(src[i])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for k in range(reps):
        dst[k % nbuf].copy_(src[k % nbuf])
    b.record(stream)
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / reps


def cpu_baseline(workload, seconds_target=12.0):
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
    n = args.n_per_gpu * args.gpus
    # bounded sample: at most ~2 minutes of wall clock for steps+warmup at ~5e6 env-steps/s/core
    budget = 120.0 * 5e6 * cores
    n_run = int(min(n, max(cores * 1024, budget / max(1, args.steps + args.warmup))))
    t, n_done = P.time_port(args.workload, n_run, args.steps, args.warmup, cores)
    value = n_done * args.steps / t
    sample = f"{args.steps} steps x {n_done} trajectories of the {n}-trajectory workload, one NumPy env per core ({cores} procs)"
    line = {"impl": "reference", "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.workload, n, args.n_per_gpu),
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ GPU legs
class Ctx:
    """What every leg needs: ranks, the barrier, reductions over ranks."""

    def __init__(self, torch, dist, world, rank, local_rank):
        self.torch, self.dist, self.world, self.rank, self.local_rank = torch, dist, world, rank, local_rank

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]


def device_buffers(torch, N, A, D, tdt, esz, action_value):
    """Rotating buffer sets so that the bytes touched between two uses of a buffer exceed L2 (126 MB)."""
    set_bytes = N * (A + D + 1) * esz
    n_sets = int(np.ceil(2 * L2_BYTES / set_bytes)) + 1
    acts = [torch.full((N, A), action_value, dtype=tdt, device="cuda") for _ in range(n_sets)]
    obs = [torch.empty((N, D), dtype=tdt, device="cuda") for _ in range(n_sets)]
    rew = [torch.empty((N,), dtype=tdt, device="cuda") for _ in range(n_sets)]
    return acts, obs, rew, n_sets, set_bytes


def timed_windows(ctx, env, bufs, stream, steps, warmup, reps, mem_device):
    """`reps` windows of EXACTLY `steps` env-steps (auto-resetting at episode end like the SB3 VecEnv adapter), each between
    barriers and timed with two CUDA events on the env's stream; returns per-window ms (max over ranks)."""
    torch = ctx.torch
    acts, obs, rew, n_sets, _ = bufs
    k_global = [0]

    def run(k_steps):
        for _ in range(k_steps):
            i = k_global[0] % n_sets
            k_global[0] += 1
            if env.step(acts[i], obs[i], rew[i], mem=mem_device):
                env.reset(obs[i], mem=mem_device)

    env.reset(obs[0], mem=mem_device)
    run(warmup)
    ctx.barrier()
    local_ms = []
    for _ in range(reps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        ev0.record(stream)
        run(steps)
        ev1.record(stream)
        ctx.barrier()
        local_ms.append(ev0.elapsed_time(ev1))
    return ctx.max_over_ranks(local_ms)


def graph_leg(ctx, make, bufs, stream, N, mem_device, replays=5):
    """reset + one whole episode (N_STEPS steps) captured ONCE into a CUDA graph on the env's stream, then replayed: the
    launch path of a device-resident learner that captures its loop (examples/cuda_graph_episode.py).  Runs on its own
    handle (a handle that was captured keeps its draw counters on the device from then on)."""
    torch = ctx.torch
    acts, obs, rew, n_sets, _ = bufs
    facade = make()
    env = facade._ensure_native()
    env.set_stream(stream.cuda_stream)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        env.reset(obs[0], mem=mem_device)
        for k in range(N_STEPS):
            i = k % n_sets
            env.step(acts[i], obs[i], rew[i], mem=mem_device)
        env.fold_counters()
    g.replay()
    ctx.barrier()
    ms = []
    for _ in range(replays):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        ev0.record(stream)
        g.replay()
        ev1.record(stream)
        ctx.barrier()
        ms.append(ev0.elapsed_time(ev1))
    ms = sorted(ctx.max_over_ranks(ms))
    med = ms[len(ms) // 2]
    del g
    facade.close()
    return {"value": N * ctx.world * N_STEPS / (med * 1e-3), "unit": "env-steps/s", "us_per_step": 1e3 * med / N_STEPS,
            "steps_per_graph": N_STEPS, "replays_timed": replays,
            "note": "one CUDA graph = mbt_reset + 200 mbt_step launches (+ the counter fold); median replay, reset included"}


def pcie_ceiling(ctx, h2d_bytes, d2h_bytes, reps=20):
    """Bare cudaMemcpyAsync of one step's bytes (H2D on one stream, D2H on another, both waited for, like mbt_step does), on
    all ranks at the same time: the host-link ceiling of THIS box at THIS concurrency.  No kernels, no library code."""
    torch = ctx.torch
    h_src = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    h_dst = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    d_dst = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    d_src = torch.ones(d2h_bytes, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def step():
        with torch.cuda.stream(s_in):
            d_dst.copy_(h_src, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_dst.copy_(d_src, non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()

    for _ in range(3):
        step()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = time.perf_counter() - t0
    ctx.barrier()
    (dt,) = ctx.max_over_ranks([dt])
    return {"ms_per_step": 1e3 * dt / reps, "gbs": ctx.world * (h2d_bytes + d2h_bytes) / (dt / reps) / 1e9}


def e2e_leg(ctx, facade, workload, steps):
    """The call a user makes: TradingEnvironment.step(numpy action) -> numpy obs, rewards, dones, infos.  Every step: H2D
    of the (N,A) action array from pinned host memory, D2H of (N,D) observations + (N,) rewards, results read on the host."""
    h_act = facade.pinned_actions()
    h_act[:] = fixed_action_value(workload)
    facade.reset()
    checksum = 0.0
    for _ in range(5):  # warm-up with the loop's own variable names: the output pool reaches its steady state (2 blocks)
        o, r, d, _i = facade.step(h_act)
        if d[0]:
            facade.reset()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        o, r, d, _i = facade.step(h_act)
        checksum += float(r[0]) + float(o[-1, 0])  # the step's result is read on the host
        if d[0]:
            facade.reset()
    dt = time.perf_counter() - t0
    ctx.barrier()
    (dt,) = ctx.max_over_ranks([dt])
    return dt, checksum


def episode_leg(ctx, env, stream, N, A, tdt, workload, mem_device, episodes=8):
    """Per-episode statistics: fused on-device rollout, summary all-reduced and returns all-gathered over NCCL by the
    library (mbt_group_rollout).  Two returns buffers alternate, so an episode's gather overlaps the next one's rollout."""
    from mbt_gym_b200 import _abi, sharding

    torch = ctx.torch
    world = ctx.world
    pol = _abi.mbt_policy()
    pol.kind = _abi.MBT_POL_FIXED
    for j in range(A):
        pol.fixed[j] = fixed_action_value(workload)
    ret = [torch.empty((N,), dtype=tdt, device="cuda") for _ in range(2)]
    all_ret = [torch.empty((N * world,), dtype=tdt, device="cuda") for _ in range(2)] if world > 1 else ret

    def one(i):
        env.reset(mem=mem_device)
        if world > 1:
            return env.group_rollout(pol, ret[i % 2], all_ret[i % 2])
        return env.rollout(pol, ret[i % 2], None, mem=mem_device)

    one(0)  # warm-up episode (first use of a collective also sets up its NCCL channels)
    if world > 1:
        env.group_wait()
    ctx.barrier()
    # (a) rollout kernel alone: events around the launch, no collective
    env.reset(mem=mem_device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record(stream)
    summ_local = env.rollout(pol, ret[0], None, mem=mem_device)
    e1.record(stream)
    ctx.barrier()
    rollout_ms = e0.elapsed_time(e1)
    # (b) the summary collective alone: all-reduce of the device-resident summary + its 56-byte D2H, wall clock on the host
    coll_ms = 0.0
    if world > 1:
        env.group_summary(summ_local)
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            merged_struct = env.group_summary(summ_local)
        coll_ms = 1e3 * (time.perf_counter() - t0) / 10
    # (c) whole episodes back to back: reset + rollout + all-reduce (+ overlapped gather), wall clock per episode (each call
    # returns when the GLOBAL summary is on the host; the gather of episode i runs behind episode i+1's rollout)
    ctx.barrier()
    per_ep = []
    t0 = time.perf_counter()
    for i in range(episodes):
        t1 = time.perf_counter()
        summ = one(i)
        per_ep.append(time.perf_counter() - t1)
    if world > 1:
        env.group_wait()
    torch.cuda.synchronize()
    ep_ms = 1e3 * (time.perf_counter() - t0) / episodes
    ep_med_ms = 1e3 * float(np.median(per_ep))
    ctx.barrier()
    rollout_ms, coll_ms, ep_ms, ep_med_ms = ctx.max_over_ranks([rollout_ms, coll_ms, ep_ms, ep_med_ms])
    merged = sharding.summary_struct_to_dict(summ)
    table = sharding.results_table(merged, A)
    gathered = all_ret[(episodes - 1) % 2]
    return {"fused_rollout_ms": rollout_ms, "fused_rollout_env_steps_per_sec": N * world * summ.steps / (rollout_ms * 1e-3),
            "summary_collective_ms": coll_ms, "episode_ms": ep_ms, "episode_ms_median": ep_med_ms, "episodes_timed": episodes,
            "episode_env_steps_per_sec": N * world * summ.steps / (ep_ms * 1e-3),
            "collective": ("library (mbt_group_rollout): ncclAllReduce of the 7-double device summary on the env's stream; "
                           "ncclAllGather of returns on the group's stream, overlapped with the next episode") if world > 1 else "none (1 GPU)",
            "returns_gathered": int(gathered.numel()), "trajectories_summarised": int(merged["count"]),
            "mean_episode_return": table["Mean PnL"], "std_episode_return": table["Std PnL"],
            "mean_terminal_inventory": table["Mean terminal inventory"],
            "policy": f"fixed action {fixed_action_value(workload)}"}, float(gathered.double().mean())


def rollout_roofline(episode, precision, sm_max_mhz):
    """Fused rollout against its ceiling, the warp-instruction issue rate ('report it against the RNG/ALU
    ceiling'): executed warp-instructions per launch from the committed ncu summary / duration measured here, vs
    148 SMs x 4 schedulers x 1 instruction per clock."""
    T = "double" if precision == "f64" else "float"
    prof, src = ncu_profile_value(("mbt_rollout_kernel<" + T, "0, 0>(RolloutArgs"), ["smsp__inst_executed.sum"])
    if prof is None:
        return None
    inst = prof["smsp__inst_executed.sum"]
    peak = SM_COUNT * 4 * sm_max_mhz * 1e6
    achieved = inst / (episode["fused_rollout_ms"] * 1e-3)
    return {"bound": "issue", "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "G warp-instructions/s", "frac": achieved / peak,
            "warp_instructions_per_launch": inst, "warp_instructions_per_warp_step": inst / (N_PER_GPU / 32 * N_STEPS),
            "source": src, "note": "no HBM traffic per step in this mode (state in registers): the ceiling is instruction issue; "
                                   "peak = 148 SMs x 4 schedulers x sm_max_mhz; float64 instructions issue at half rate, so the "
                                   "reachable fraction in f64 is below 1"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="as", choices=WORKLOADS)
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--reps", type=int, default=11, help="how many times the K-step window is timed (median reported)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--n-per-gpu", type=int, default=N_PER_GPU)
    ap.add_argument("--no-episode-stats", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip value_graph, n_sweep, configs4, float32-io legs")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and os.environ.get("NCCL_DEBUG") and "NCCL_DEBUG_FILE" not in os.environ:
        # NCCL's log (rank / nranks lines at NCCL_DEBUG=INFO) is evidence, not noise: keep it, but on stderr, so that stdout
        # stays the one JSON line
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"

    import torch
    import torch.distributed as dist

    from mbt_gym_b200 import _abi, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mbt_gym_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    ctx = Ctx(torch, dist, world, rank, local_rank)
    DEV = _abi.MBT_MEM_DEVICE

    tdt = torch.float64 if args.precision == "f64" else torch.float32
    esz = 8 if args.precision == "f64" else 4
    prec_name = "float64" if args.precision == "f64" else "float32"
    n_local = args.n_per_gpu
    facade = make_env(args.workload, prec_name, n_local, rank * n_local, local_rank)
    env = facade._ensure_native()  # the handle behind the public API: same kernels, explicit device buffers
    N, A, D = env.N, env.A, env.D
    if world > 1:
        sharding.create_group(env)  # library-level NCCL group (ncclCommInitRank inside libmbt_b200)
    stream = torch.cuda.Stream()  # a real (non-default) stream: the env's kernels and the timing events share it
    torch.cuda.set_stream(stream)
    env.set_stream(stream.cuda_stream)
    bufs = device_buffers(torch, N, A, D, tdt, esz, fixed_action_value(args.workload))
    n_sets, set_bytes = bufs[3], bufs[4]

    # ---- value: `reps` windows of exactly K steps
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    launches0 = env.launch_count()
    windows = timed_windows(ctx, env, bufs, stream, args.steps, args.warmup, args.reps, DEV)
    launches = (env.launch_count() - launches0 - 0) // max(1, args.reps)  # per window (resets included)
    srt = sorted(windows)
    elapsed_ms = srt[len(srt) // 2]
    # separate instrumented pass (NOT the timed region): every launch bracketed by its own two events
    ctx.barrier()
    env.enable_timing(1)
    for k in range(min(args.steps, 100)):
        if env.step(bufs[0][k % n_sets], bufs[1][k % n_sets], bufs[2][k % n_sets], mem=DEV):
            env.reset(mem=DEV)
    ctx.barrier()
    ktimes_bracketed = env.kernel_times_ms()
    env.enable_timing(0)
    b_step = algorithmic_bytes_per_env_step(A, D, esz)
    copy_us = size_matched_copy_us(N * b_step, stream) if rank == 0 else None

    # ---- value_graph
    graph = None
    if not args.no_extras:
        try:
            graph = graph_leg(ctx, lambda: make_env(args.workload, prec_name, n_local, rank * n_local, local_rank), bufs, stream, N, DEV)
        except Exception as ex:  # noqa: BLE001  (a leg that fails must not cost the whole line)
            graph = {"error": repr(ex)[:300]}
            ctx.barrier()

    # ---- per-episode statistics (fused rollout + the library's collective)
    episode = None
    if not args.no_episode_stats:
        episode, _ = episode_leg(ctx, env, stream, N, A, tdt, args.workload, DEV)
        peak_gbs, peak_src, sm_max_mhz = measured_peaks()
        episode["roofline"] = rollout_roofline(episode, args.precision, sm_max_mhz)
    env.set_stream(None)  # back to the handle's own stream for the host-buffer path

    # ---- e2e
    e2e_steps = args.e2e_steps or max(10, min(args.steps, 50))
    e2e_s, checksum = e2e_leg(ctx, facade, args.workload, e2e_steps)
    ceiling = pcie_ceiling(ctx, N * A * esz, N * (D + 1) * esz)
    clocks = sampler.stop() if sampler else None

    # ---- secondary: the same host path with float32 caller buffers over the same float64 arithmetic (io_dtype), every N
    e2e_f32io = None
    if args.precision == "f64" and not args.no_extras:
        f32io = make_env(args.workload, "float64", n_local, rank * n_local, local_rank, io_dtype=np.float32)
        dt32, _ = e2e_leg(ctx, f32io, args.workload, e2e_steps)
        ceil32 = pcie_ceiling(ctx, N * A * 4, N * (D + 1) * 4)
        e2e_f32io = {"value": N * world * e2e_steps / dt32, "unit": "env-steps/s", "ms_per_step": 1e3 * dt32 / e2e_steps,
                     "h2d_bytes_per_step": N * A * 4, "d2h_bytes_per_step": N * (D + 1) * 4,
                     "pcie_peak_gbs": ceil32["gbs"], "frac": ceil32["ms_per_step"] / (1e3 * dt32 / e2e_steps),
                     "note": "io_dtype=float32: float64 state and arithmetic, float32 action/observation/reward arrays"}
        f32io.close()

    # ---- configs[4]: OE + OU midprice, 2^20 trajectories per GPU, fused rollout + NCCL gather of all returns
    configs4 = None
    if not args.no_extras and not args.no_episode_stats and args.workload != "oe":
        oe = make_env("oe", prec_name, n_local, rank * n_local, local_rank)
        oe_env = oe._ensure_native()
        if world > 1:
            sharding.create_group(oe_env)
        oe_env.set_stream(stream.cuda_stream)
        configs4, mean_all = episode_leg(ctx, oe_env, stream, N, oe_env.A, tdt, "oe", DEV)
        configs4["workload"] = workload_name("oe")
        configs4["num_trajectories"] = N * world
        configs4["mean_of_gathered_returns"] = mean_all
        oe_env.set_stream(None)
        oe.close()

    facade.close()
    del bufs
    torch.cuda.empty_cache()
    # ---- n_sweep: fixed launch cost vs per-trajectory cost of the step kernel (1 GPU only); every other handle is closed by
    # now (a live handle's L2 window / persisting set-aside would cost the larger sizes a few per cent)
    n_sweep = None
    peak_gbs, peak_src, sm_max_mhz = measured_peaks()
    if world == 1 and not args.no_extras:
        n_sweep = []
        for logn in (20, 22, 24):
            n = 1 << logn
            f = make_env(args.workload, prec_name, n, 0, local_rank)
            e = f._ensure_native()
            e.set_stream(stream.cuda_stream)
            b = device_buffers(torch, n, A, D, tdt, esz, fixed_action_value(args.workload))
            w = sorted(timed_windows(ctx, e, b, stream, 20, 5, 5, DEV))
            us = 1e3 * w[len(w) // 2] / 20
            n_sweep.append({"num_trajectories": n, "us_per_step": us, "achieved_gbs": n * b_step / (us * 1e-6) / 1e9,
                            "frac": n * b_step / (us * 1e-6) / 1e9 / peak_gbs})
            e.set_stream(None)
            f.close()
            del b
            torch.cuda.empty_cache()
        if len(n_sweep) == 3:  # least-squares fit  t(N) = fixed + per_traj * N
            xs = np.array([r["num_trajectories"] for r in n_sweep], float)
            ys = np.array([r["us_per_step"] for r in n_sweep], float)
            slope, icpt = np.polyfit(xs, ys, 1)
            n_sweep.append({"fit_fixed_us": float(icpt), "fit_ps_per_trajectory": float(slope * 1e6),
                            "fit_asymptotic_frac": float(b_step / (slope * 1e-6) / 1e9 / peak_gbs)})

    e2e_ms = e2e_s * 1e3
    if rank == 0:
        total_n = N * world
        value = total_n * args.steps / (elapsed_ms * 1e-3)
        e2e_value = total_n * e2e_steps / (e2e_ms * 1e-3)
        mean_kernel_ms = elapsed_ms / args.steps
        achieved = N * b_step / (mean_kernel_ms * 1e-3) / 1e9
        T = "double" if args.precision == "f64" else "float"
        dram = ["dram__bytes_read.sum", "dram__bytes_write.sum"]
        prof = prof_src = cold = cold_src = None
        if args.workload == "as" and N == N_PER_GPU:
            # steady state first (ncu --cache-control none on the running step loop), the cold single-launch capture second
            prof, prof_src = ncu_profile_value(("mbt_step_kernel<" + T,), dram, f"profiles/r2_step_steady_{args.precision}.ncu_summary.csv")
            cold, cold_src = ncu_profile_value(("mbt_step_kernel<" + T, "0, 1, 1, 0, 0, 1, 0, 0, 0, 0>"), dram, f"profiles/r2_targets_{args.precision}.ncu_summary.csv")
            if prof is None:
                prof, prof_src = cold, cold_src
        traffic = (prof["dram__bytes_read.sum"] + prof["dram__bytes_write.sum"]) if prof else None
        traffic_cold = (cold["dram__bytes_read.sum"] + cold["dram__bytes_write.sum"]) if cold else None
        line = {
            "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
            "data": "synthetic",
            "config": config_dict(args.workload, total_n, N),
            "window": {"reps": args.reps, "statistic": "median", "ms_per_step_min": srt[0] / args.steps,
                       "ms_per_step_max": srt[-1] / args.steps, "spread": (srt[-1] - srt[0]) / elapsed_ms,
                       "buffer_sets": n_sets, "bytes_streamed_between_reuses": n_sets * set_bytes},
            "value_graph": graph,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": N * A * esz,
                    "d2h_bytes_per_step": N * (D + 1) * esz, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "achieved_gbs": world * N * (A + D + 1) * esz / (e2e_ms / e2e_steps * 1e-3) / 1e9,
                    "pcie_peak_gbs": ceiling["gbs"], "pcie_peak_ms_per_step": ceiling["ms_per_step"],
                    "frac": ceiling["ms_per_step"] / (e2e_ms / e2e_steps),
                    "frac_note": "bare-copy time / e2e step time: the bare copies wait for both directions every step, the library "
                                 "pipelines chunks, so a value slightly above 1 is possible when the host link is saturated (N > 1)",
                    "pcie_peak_how": f"bare cudaMemcpyAsync of the same bytes (H2D + D2H on two streams, both waited for per step) on "
                                     f"all {world} ranks at once, measured in this run; a k-GPU table is in profiles/r2_pcie_probe_*.json"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                         "traffic": traffic, "traffic_source": prof_src,
                         "traffic_cold_launch": traffic_cold, "traffic_cold_source": cold_src,
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu captures, read "
                                         "at run time.  `traffic`: launches 60+ of the running step loop under `ncu --cache-control none` "
                                         "(steady state: the state columns are L2-resident through the access-policy window, DRAM moves "
                                         "actions in and observations / rewards out).  `traffic_cold_launch`: one launch under `ncu --set "
                                         "full` with ncu's cache flush -- reads = actions + state columns; written lines still in L2 when "
                                         "the kernel's window ends are not counted",
                         "dram_frac": (traffic / (mean_kernel_ms * 1e-3) / 1e9 / peak_gbs) if traffic else None,
                         "frac_note": "achieved = ALGORITHMIC bytes (SURVEY 8d: 104 B per env-step in f64) / median window time per "
                                      "step.  At N = 2^20 a step is ~5 us of fixed launch/ramp/drain cost plus ~15 ps per trajectory "
                                      "(n_sweep fit): co-limited by launch latency and instruction issue, at the speed of a "
                                      "size-matched device copy; the asymptotic per-trajectory rate is the HBM-bound one",
                         "peak_source": peak_src, "kernel": "mbt_step_kernel",
                         "size_matched_copy_us": copy_us,
                         "frac_of_size_matched_copy": (copy_us * 1e-3 / mean_kernel_ms) if copy_us else None,
                         "algorithmic_bytes_per_env_step": b_step, "mean_kernel_ms": mean_kernel_ms,
                         "duration_method": "CUDA events bracketing each K-step window / K (upper bound: includes launch gaps and "
                                            "1 reset per 200 steps); median over the windows",
                         "kernel_launches_timed": int(args.steps * args.reps),
                         "two_event_bracket_kernel_us": float(1e3 * np.median(ktimes_bracketed)) if len(ktimes_bracketed) else None,
                         "n_sweep": n_sweep},
            "clocks": clocks,
            "episode_stats": episode,
            "configs4": configs4,
            "e2e_float32_io": e2e_f32io,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        sys.stdout.write(json.dumps(line) + "\n")
        sys.stdout.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
