"""Multi-process host logic on CPU (gloo, world_size 2): shard bounds, summary all-reduce, returns all-gather.
The summaries come from the CPU oracle (test infrastructure) stepping two shards with global trajectory ids."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mbt_gym_b200 import _abi, sharding
from oracle import oracle as O
from tests.helpers import Golden


def test_shard_bounds_cover_exactly():
    for n in (1, 7, 8, 1000, 1 << 20, (1 << 23) + 5):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _oracle_episode_summary(cfg, seed, action):
    """Fixed-action episode on the oracle -> the same additive moments mbt_rollout reports."""
    orc = O.OracleEnv(cfg)
    orc.seed(seed)
    orc.reset()
    R = np.zeros(orc.N)
    r2 = 0.0
    steps, done = 0, False
    while not done:
        _o, r, done = orc.step(np.full((orc.N, orc.A), action))
        R += r
        r2 += float((r ** 2).sum())
        steps += 1
    q = orc.state[:, 1]
    return dict(count=orc.N, steps=steps, sum_return=R.sum(), sum_return_sq=(R ** 2).sum(), sum_q=q.sum(),
                sum_q_sq=(q ** 2).sum(), sum_action=action * orc.N * orc.A * steps, sum_reward_sq=r2, clipped=0), R


def _worker(rank, world, port, n_total, out_queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = Golden("as_pnl_tight")
    lo, hi = sharding.shard_bounds(n_total, world, rank)
    cfg = g.config(_abi.MBT_F64, num_trajectories=hi - lo, traj_offset=lo)
    local, R = _oracle_episode_summary(cfg, 77, 0.6)
    merged = sharding.allreduce_summary(local)
    gathered = sharding.allgather_returns(torch.from_numpy(R))  # shard sizes differ by one when n_total is odd
    if rank == 0:
        out_queue.put((merged, gathered.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n_total", [96, 97])
def test_two_rank_summary_equals_single_process(n_total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = Golden("as_pnl_tight")
    whole, R = _oracle_episode_summary(g.config(_abi.MBT_F64, num_trajectories=n_total, traj_offset=0), 77, 0.6)
    for f in sharding.SUMMARY_FIELDS:
        np.testing.assert_allclose(merged[f], whole[f], rtol=1e-12, err_msg=f)
    assert np.array_equal(gathered, R), "global trajectory ids: shards reproduce the single-process trajectories"
    # merge_summaries (no process group) agrees too
    parts = []
    for r in range(3):
        lo, hi = sharding.shard_bounds(n_total, 3, r)
        parts.append(_oracle_episode_summary(g.config(_abi.MBT_F64, num_trajectories=hi - lo, traj_offset=lo), 77, 0.6)[0])
    m3 = sharding.merge_summaries(parts)
    for f in sharding.SUMMARY_FIELDS:
        np.testing.assert_allclose(m3[f], whole[f], rtol=1e-12, err_msg=f)
    table = sharding.results_table(m3, 2)
    assert table["Mean spread"] == pytest.approx(1.2) and table["Mean PnL"] == pytest.approx(R.mean())
