"""Rollout helpers (reference: mbt_gym/gym/helpers/generate_trajectory.py:8-38, helpers/plotting.py:94-108)."""
import numpy as np

from ..index_names import INVENTORY_INDEX


def generate_trajectory(env, agent, seed=None, include_log_probs=False):
    """reset, then (agent.get_action, env.step) until done; returns observations (N, D, T+1), actions (N, A, T),
    rewards (N, 1, T) [and log-probs] exactly like the reference's helper."""
    if seed is not None:
        env.seed(seed)
    n, d, a_dim, t_max = env.num_trajectories, env.observation_space.shape[0], env.action_space.shape[0], env.n_steps
    observations = np.zeros((n, d, t_max + 1))
    actions = np.zeros((n, a_dim, t_max))
    rewards = np.zeros((n, 1, t_max))
    log_probs = None
    obs = env.reset()
    observations[:, :, 0] = obs
    for count in range(t_max):
        if include_log_probs:
            action, log_prob = agent.get_action(obs, include_log_probs=True)
            if log_probs is None:
                import torch

                log_probs = torch.zeros((n, a_dim, t_max))
            log_probs[:, :, count] = log_prob
        else:
            action = agent.get_action(obs)
        obs, reward, done, _ = env.step(action)
        actions[:, :, count] = action
        observations[:, :, count + 1] = obs
        rewards[:, :, count] = np.asarray(reward).reshape(-1, 1)
        if (n > 1 and done[0]) or (n == 1 and done):
            break
    return (observations, actions, rewards, log_probs) if include_log_probs else (observations, actions, rewards)


def generate_trajectory_fused(env, agent, seed=None):
    """`generate_trajectory` for agents with an on-device form (`agent.to_policy`): the whole episode runs in ONE kernel
    and the recording comes back in the reference's shapes -- observations (N, D, T+1), actions (N, A, T),
    rewards (N, 1, T) -- as transposed views of the time-major device recording."""
    if seed is not None:
        env.seed(seed)
    env.reset()
    native = env._native
    _summary, obs, act, rew = native.rollout_record(agent.to_policy(env), env.n_steps)
    return obs.transpose(1, 2, 0), act.transpose(1, 2, 0), rew.T[:, None, :]


RESULT_COLUMNS = ["Mean spread", "Mean PnL", "Std PnL", "Mean terminal inventory", "Std terminal inventory"]


def results_from_trajectory(observations, actions, rewards):
    """The numbers of the reference's results table (plotting.py:96-108) from recorded trajectories."""
    total = rewards.sum(axis=-1).reshape(-1)
    q_t = observations[:, INVENTORY_INDEX, -1]
    return dict(zip(RESULT_COLUMNS, [2 * np.mean(actions.mean(axis=(-1, -2))), np.mean(total), np.std(total),
                                     np.mean(q_t), np.std(q_t)]))


def generate_results_table(env, agent):
    """Step-by-step rollout with a host agent -> results dict (and per-trajectory total rewards)."""
    assert env.num_trajectories > 1, "To generate a results table, env must roll out > 1 trajectory."
    observations, actions, rewards = generate_trajectory(env, agent)
    return results_from_trajectory(observations, actions, rewards), rewards.sum(axis=-1).reshape(-1)


def generate_results_table_fused(env, agent):
    """Same table from the fused on-device rollout (`agent.to_policy`): no per-step host traffic."""
    env.reset()
    summary, returns, q_t = env.rollout_summary(agent.to_policy(env), return_trajectory_stats=True)
    n, a_dim = summary.count, env.action_space.shape[0]
    mean_r, mean_q = summary.sum_return / n, summary.sum_q / n
    out = dict(zip(RESULT_COLUMNS, [2 * summary.sum_action / (n * summary.steps * a_dim), mean_r,
                                    np.sqrt(max(summary.sum_return_sq / n - mean_r ** 2, 0.0)), mean_q,
                                    np.sqrt(max(summary.sum_q_sq / n - mean_q ** 2, 0.0))]))
    return out, returns
