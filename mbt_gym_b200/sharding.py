"""Multi-GPU: trajectory sharding and the ONE collective of the path.

Trajectories never interact, so a batch of N trajectories is cut into contiguous shards, one process per GPU, each
shard a `TradingEnvironment(num_trajectories=n_local, traj_offset=first_global_id)`.  The Philox counters use GLOBAL
trajectory ids, so the simulated trajectories are identical for 1, 2, 4 or 8 GPUs.  There is no per-step
communication; the only exchange is per episode: an all-reduce (sum) of the `mbt_summary` moments and, optionally, an
all-gather of per-trajectory episode returns (BASELINE configs[4] "NCCL gather of returns").

One exception, inherited from the reference: the Triangular and Power fill functions reduce the quoted depths over the
WHOLE batch (`np.max(depths, 0)`, fill_probability_models.py:82,113), which couples the trajectories of a step.  With a
library group attached (`create_group`) `step()` all-reduces (max) the shard maxima over NCCL before the step, so results
are again independent of the shard layout; WITHOUT a group every shard reduces over its own trajectories only, i.e. it
behaves like one worker of the reference's MultiprocessTradingEnv.

On GPUs the collectives run INSIDE the library (`mbt_group_*`, include/mbt_b200.h): NCCL on device buffers, on the
handle's stream, the summary never leaving the device before it is reduced.  `create_group` only ships the NCCL unique id
over `torch.distributed`.  The `allreduce_summary` / `allgather_returns` helpers below do the same exchange through
`torch.distributed` itself; they are what the CPU (gloo) tests exercise and a cross-check for the library path.

Replaces the reference's process fan-out (mbt_gym/gym/MultiprocessTradingEnv.py:72-116: SubprocVecEnv over pipes,
results concatenated by `flatten_multi`).
"""
import numpy as np

SUMMARY_FIELDS = ("count", "steps", "sum_return", "sum_return_sq", "sum_q", "sum_q_sq", "sum_action", "sum_reward_sq",
                  "clipped")
# `steps` is identical on every rank (uniform clock): it is carried, not summed
_SUMMED = tuple(f for f in SUMMARY_FIELDS if f != "steps")


def shard_bounds(num_trajectories, world_size, rank):
    """Contiguous shard [lo, hi) of rank `rank`; sizes differ by at most one; covers [0, N) exactly."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(int(num_trajectories), int(world_size))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def summary_to_array(summary):
    """mbt_summary (ctypes struct or dict) -> float64 vector in SUMMARY_FIELDS order."""
    get = (lambda f: summary[f]) if isinstance(summary, dict) else (lambda f: getattr(summary, f))
    return np.array([float(get(f)) for f in SUMMARY_FIELDS], dtype=np.float64)


def array_to_summary(vec):
    out = {f: float(v) for f, v in zip(SUMMARY_FIELDS, vec)}
    for f in ("count", "steps", "clipped"):
        out[f] = int(round(out[f]))
    return out


def merge_summaries(summaries):
    """Sum of shard summaries = the summary of the whole batch (moments are additive)."""
    vecs = [summary_to_array(s) for s in summaries]
    steps = {int(v[SUMMARY_FIELDS.index("steps")]) for v in vecs}
    if len(steps) != 1:
        raise ValueError(f"shards disagree on the episode length: {sorted(steps)}")
    tot = np.sum(vecs, axis=0)
    tot[SUMMARY_FIELDS.index("steps")] = steps.pop()
    return array_to_summary(tot)


def results_table(summary, action_dim):
    """The reference's results table (mbt_gym/gym/helpers/plotting.py:96-108) from the additive moments."""
    s = summary if isinstance(summary, dict) else array_to_summary(summary_to_array(summary))
    n = s["count"]
    mean_r, mean_q = s["sum_return"] / n, s["sum_q"] / n
    return {
        "Mean spread": 2 * s["sum_action"] / (n * s["steps"] * action_dim),
        "Mean PnL": mean_r,
        "Std PnL": float(np.sqrt(max(s["sum_return_sq"] / n - mean_r ** 2, 0.0))),
        "Mean terminal inventory": mean_q,
        "Std terminal inventory": float(np.sqrt(max(s["sum_q_sq"] / n - mean_q ** 2, 0.0))),
    }


def create_group(native_env, group=None):
    """Join this rank's device handle (`env._ensure_native()` / `_lib.NativeEnv`) to a library-level NCCL group spanning
    the ranks of the `torch.distributed` group: rank 0 creates the NCCL unique id, `broadcast_object_list` ships it."""
    import torch.distributed as dist

    from . import _lib

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [_lib.NativeEnv.group_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    native_env.group_create(box[0], rank, world)
    return native_env.group_info()


def summary_struct_to_dict(summary):
    return array_to_summary(summary_to_array(summary))


def allreduce_summary(summary, group=None, device=None):
    """All-reduce (sum) of the summary moments over `torch.distributed` (NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    vec = summary_to_array(summary)
    steps_idx = SUMMARY_FIELDS.index("steps")
    t = torch.from_numpy(vec.copy())
    if device is not None:
        t = t.to(device)
    steps = t[steps_idx].clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    # every rank ran the same number of steps: check instead of trusting
    steps_max = steps.clone()
    dist.all_reduce(steps_max, op=dist.ReduceOp.MAX, group=group)
    if float(steps_max) != float(steps):
        raise RuntimeError("ranks disagree on the episode length")
    t[steps_idx] = steps
    return array_to_summary(t.cpu().numpy())


def allgather_returns(returns, group=None):
    """All-gather per-trajectory episode returns (torch tensor) in global-id order; shard lengths may differ by one
    (`shard_bounds` when N % world != 0), so the lengths are exchanged first and the short shards padded."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    returns = returns.contiguous().reshape(-1)
    n = torch.tensor([returns.numel()], dtype=torch.int64, device=returns.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x) for x in sizes]
    width = max(sizes)
    if min(sizes) == width:
        out = torch.empty((world * width,), dtype=returns.dtype, device=returns.device)
        dist.all_gather_into_tensor(out, returns, group=group)
        return out
    padded = torch.zeros((width,), dtype=returns.dtype, device=returns.device)
    padded[: returns.numel()] = returns
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:k] for p, k in zip(parts, sizes)])
