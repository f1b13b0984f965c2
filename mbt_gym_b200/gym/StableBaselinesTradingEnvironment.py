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


class _TerminalInfos(list):
    """The `infos` of an episode-ending step: a real `list` of N dicts -- {"terminal_observation": terminal_obs[i]} and,
    with `monitor=True`, {"episode": {"r": return, "l": length, "t": seconds}} (what SB3's VecMonitor would add) -- whose
    items are built on first use instead of by an O(N) Python loop at every episode end (reference :31-35).
    `len()` and single-item reads cost O(1); a built item is kept, so `infos[i]["key"] = v` and `infos[i] = d` stick;
    anything that looks at the whole list (iteration, slices, `list(infos)`, `==`, pickling ...) materialises all items
    once and from then on it IS an ordinary list."""

    def __init__(self, terminal_obs, episode=None):
        super().__init__()
        self._obs = terminal_obs
        self._episode = episode  # (returns (N,), length, elapsed seconds) or None
        self._n = int(terminal_obs.shape[0] if terminal_obs is not None else episode[0].shape[0])
        self._made = {}
        self._filled = False

    def _make(self, i):
        info = {}
        if self._obs is not None:
            info["terminal_observation"] = self._obs[i, :]
        if self._episode is not None:
            ret, length, elapsed = self._episode
            info["episode"] = {"r": float(ret[i]), "l": int(length), "t": elapsed}
        return info

    def _fill(self):
        if not self._filled:
            made = self._made
            list.extend(self, (made[i] if i in made else self._make(i) for i in range(self._n)))
            self._filled, self._made = True, {}
        return self

    def __len__(self):
        return list.__len__(self) if self._filled else self._n

    def __bool__(self):
        return len(self) > 0

    def __getitem__(self, i):
        if self._filled or isinstance(i, slice):
            return list.__getitem__(self._fill(), i)
        j = i + self._n if i < 0 else i
        if not 0 <= j < self._n:
            raise IndexError("list index out of range")
        if j not in self._made:
            self._made[j] = self._make(j)
        return self._made[j]

    def __setitem__(self, i, value):
        if self._filled or isinstance(i, slice):
            return list.__setitem__(self._fill(), i, value)
        j = i + self._n if i < 0 else i
        if not 0 <= j < self._n:
            raise IndexError("list assignment index out of range")
        self._made[j] = value

    def __iter__(self):
        return list.__iter__(self._fill())

    def __reversed__(self):
        return list.__reversed__(self._fill())

    def __contains__(self, x):
        return list.__contains__(self._fill(), x)

    def __eq__(self, other):
        return list.__eq__(self._fill(), other)

    __hash__ = None

    def __repr__(self):
        return list.__repr__(self) if self._filled else f"<{self._n} lazy episode-end infos>"

    def __reduce__(self):
        return (list, (list(self),))

    def copy(self):
        return list(self)


def _mutator(name):
    def method(self, *args, **kwargs):
        return getattr(list, name)(self._fill(), *args, **kwargs)

    method.__name__ = name
    return method


for _name in ("append", "extend", "insert", "pop", "remove", "clear", "sort", "reverse", "index", "count", "__delitem__",
              "__add__", "__iadd__", "__mul__", "__imul__", "__rmul__"):
    setattr(_TerminalInfos, _name, _mutator(_name))


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
