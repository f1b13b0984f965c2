"""GPU test (-m gpu) of the library-level NCCL group across TWO ranks (needs two visible GPUs; skipped on one):
mbt_group_rollout's summary and gathered returns equal a single handle over all trajectories, with equal and ragged
shards; and the batch-reduced fill models give shard-independent results once a group is attached."""
import os
import socket

import numpy as np
import pytest

from mbt_gym_b200 import _abi, sharding
from tests.helpers import build_facade_env, golden_specs

pytestmark = pytest.mark.gpu
SPECS = golden_specs()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _policy(A, value):
    pol = _abi.mbt_policy()
    pol.kind = _abi.MBT_POL_FIXED
    for j in range(A):
        pol.fixed[j] = value
    return pol


def _worker(rank, world, port, n_total, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    lo, hi = sharding.shard_bounds(n_total, world, rank)
    out = {}
    # (1) OE rollout: summary + gather
    spec = dict(SPECS["oe_ou_cjoe"], N=hi - lo, n_steps=20)
    env = build_facade_env(spec, device=rank, traj_offset=lo)._ensure_native()
    info = sharding.create_group(env)
    assert info == dict(rank=rank, world=world, total_trajectories=n_total)
    env.reset(mem=_abi.MBT_MEM_DEVICE)
    loc = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
    allr = torch.zeros(n_total, dtype=torch.float64, device="cuda")
    summ = env.group_rollout(_policy(1, -1.0), loc, allr)
    env.group_wait()
    out["summary"] = sharding.summary_struct_to_dict(summ)
    out["returns"] = allr.cpu().numpy()
    out["hist"] = env.inventory_histogram(60, 100, group_sum=True)  # terminal inventories of ALL ranks (NCCL sum)
    # the same exchange through torch.distributed (cross-check of the library path)
    env.reset(mem=_abi.MBT_MEM_DEVICE)
    local = env.rollout(_policy(1, -1.0), loc, None, mem=_abi.MBT_MEM_DEVICE)
    out["summary_torch"] = sharding.allreduce_summary(local, device=torch.device("cuda", rank))
    out["summary_lib_again"] = sharding.summary_struct_to_dict(env.group_summary(local))
    env.group_destroy()
    env.close()
    # (2) power fills: np.max(depths, 0) over ALL ranks' trajectories
    spec = dict(SPECS["power_fill"], N=hi - lo, n_steps=6, normalise_action=False, normalise_obs=False)
    fenv = build_facade_env(spec, device=rank, traj_offset=lo)
    sharding.create_group(fenv._ensure_native())
    acts = np.random.default_rng(1).uniform(0.0, 2.5, size=(6, n_total, 2))
    obs = [fenv.reset()[:, :].copy()]
    for k in range(6):
        o, r, d, _ = fenv.step(acts[k, lo:hi])
        obs.append(o.copy())
    out["power_obs"] = np.stack(obs)
    fenv._native.group_destroy()
    fenv.close()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [4096, 4097])
def test_two_rank_group_equals_one_handle(n_total):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single handle over all trajectories
    one = build_facade_env(dict(SPECS["oe_ou_cjoe"], N=n_total, n_steps=20))._ensure_native()
    one.reset()
    ret = np.empty(n_total)
    want = sharding.summary_struct_to_dict(one.rollout(_policy(1, -1.0), ret))
    want_hist = one.inventory_histogram(60, 100)
    assert want_hist.sum() == n_total
    one.close()
    for r in range(world):
        got = res[r]["summary"]
        assert got["count"] == n_total and got["steps"] == want["steps"] and got["clipped"] == want["clipped"]
        for f in ("sum_return", "sum_return_sq", "sum_q", "sum_q_sq", "sum_action", "sum_reward_sq"):
            np.testing.assert_allclose(got[f], want[f], rtol=1e-12, err_msg=f)  # summation order differs across shards
            # (a second episode, other draws:) the library's all-reduce against torch.distributed's on the same local summary
            np.testing.assert_allclose(res[r]["summary_lib_again"][f], res[r]["summary_torch"][f], rtol=1e-14, err_msg=f)
        assert np.array_equal(res[r]["returns"], ret), "gathered returns: global-id order, bit-identical to one handle"
        assert np.array_equal(res[r]["hist"], want_hist), "inventory histogram summed over the group"
    spec = dict(SPECS["power_fill"], N=n_total, n_steps=6, normalise_action=False, normalise_obs=False)
    fenv = build_facade_env(spec)
    acts = np.random.default_rng(1).uniform(0.0, 2.5, size=(6, n_total, 2))
    obs = [fenv.reset().copy()]
    for k in range(6):
        obs.append(fenv.step(acts[k])[0].copy())
    fenv.close()
    obs = np.stack(obs)
    lo0, hi0 = sharding.shard_bounds(n_total, world, 0)
    assert np.array_equal(res[0]["power_obs"], obs[:, lo0:hi0]) and np.array_equal(res[1]["power_obs"], obs[:, hi0:])
