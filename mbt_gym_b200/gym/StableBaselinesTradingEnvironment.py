"""stable-baselines3 `VecEnv` adapter (reference: mbt_gym/gym/StableBaselinesTradingEnvironment.py:11-66).

SB3 is optional here: when it is importable the class derives from its `VecEnv`; otherwise it is a duck-typed object
with the same methods, which is what the conformance tests exercise.  Difference from the reference: the
`terminal_observation` entries are not produced by an O(N) Python loop per episode (:31-35) -- `infos` is a lazy
sequence whose items materialise `{"terminal_observation": obs[i]}` on access.

Device mode: when `step_async` receives a CUDA torch tensor, observations and rewards stay CUDA tensors, the auto-reset
uses `reset_device()`, and the monitor accumulates the episode returns on the device -- nothing crosses PCIe per step
(`reset(device=True)` starts an episode that way).
"""
import numpy as np

try:  # pragma: no cover - depends on the installation
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase  # type: ignore
except Exception:  # noqa: BLE001
    _VecEnvBase = object


class _TerminalInfos:
    """Sequence of N dicts built on demand: {"terminal_observation": terminal_obs[i]} and, with `monitor=True`,
    {"episode": {"r": return, "l": length, "t": seconds}} -- the entries SB3's VecMonitor would add."""

    def __init__(self, terminal_obs, episode=None):
        self._obs = terminal_obs
        self._episode = episode  # (returns (N,), length, elapsed seconds) or None

    def __len__(self):
        return self._obs.shape[0] if self._obs is not None else self._episode[0].shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        info = {}
        if self._obs is not None:
            info["terminal_observation"] = self._obs[i, :]
        if self._episode is not None:
            ret, length, elapsed = self._episode
            info["episode"] = {"r": float(ret[i]), "l": int(length), "t": elapsed}
        return info

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def copy(self):
        return list(self)


class StableBaselinesTradingEnvironment(_VecEnvBase):
    def __init__(self, trading_env, store_terminal_observation_info=True, monitor=False):
        """monitor=True replaces SB3's VecMonitor (whose bookkeeping is a Python loop over all N environments at every
        episode end): episode returns are accumulated with one vectorised add per step and surface as lazy
        `info["episode"]` entries; `episode_returns` / `last_episode_statistics` expose them in bulk."""
        self.env = trading_env
        self.store_terminal_observation_info = store_terminal_observation_info
        self.monitor = monitor
        self.episode_returns = None
        self.episode_length = 0
        self.last_episode_statistics = None
        self._t_start = None
        self.actions = self.env.action_space.sample()
        if _VecEnvBase is not object:
            super().__init__(self.env.num_trajectories, self.env.observation_space, self.env.action_space)
        else:
            self.num_envs = self.env.num_trajectories
            self.observation_space, self.action_space = self.env.observation_space, self.env.action_space

    def reset(self, device=False):
        """First observation of a new episode: a NumPy array, or with `device=True` a CUDA tensor (no host round-trip)."""
        self._begin_episode(on_device=device)
        return self.env.reset_device() if device else self.env.reset()

    def _begin_episode(self, on_device=False):
        if self.monitor:
            import time

            if on_device:
                import torch

                self.episode_returns = torch.zeros((self.env.num_trajectories,), dtype=torch.float64,
                                                   device=torch.device("cuda", self.env.device))
            else:
                self.episode_returns = np.zeros((self.env.num_trajectories,), dtype=np.float64)
            self.episode_length = 0
            self._t_start = time.time()

    def step_async(self, actions):
        self.actions = actions

    def step_wait(self):
        obs, rewards, dones, infos = self.env.step(self.actions)
        on_device = bool(getattr(rewards, "is_cuda", False))  # CUDA tensors in -> CUDA tensors out
        if self.monitor:
            if self.episode_returns is None or bool(getattr(self.episode_returns, "is_cuda", False)) != on_device:
                self._begin_episode(on_device)
            self.episode_returns += rewards if on_device else np.asarray(rewards)
            self.episode_length += 1
        if dones.min():
            episode = None
            if self.monitor:
                import time

                r = self.episode_returns
                episode = (r, self.episode_length, round(time.time() - self._t_start, 6))
                r_host = r.cpu().numpy() if on_device else r  # one (N,) copy per episode, for the summary only
                self.last_episode_statistics = {"mean_return": float(r_host.mean()), "std_return": float(r_host.std()),
                                                "length": self.episode_length, "num_episodes": int(r_host.shape[0])}
            if self.store_terminal_observation_info or episode is not None:
                terminal = None
                if self.store_terminal_observation_info:
                    terminal = obs.clone() if on_device else np.array(obs, copy=True)
                infos = _TerminalInfos(terminal, episode)
            self._begin_episode(on_device)
            # SB3 convention: auto-reset, return the first observation of the new episode
            obs = self.env.reset_device() if on_device else self.env.reset()
        return obs, rewards, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        pass

    def get_attr(self, attr_name, indices=None):
        pass

    def set_attr(self, attr_name, value, indices=None):
        pass

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        pass

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in range(self.env.num_trajectories)]

    def seed(self, seed=None):
        return self.env.seed(seed)

    def get_images(self):
        pass

    @property
    def num_trajectories(self):
        return self.env.num_trajectories

    @property
    def n_steps(self):
        return self.env.n_steps
