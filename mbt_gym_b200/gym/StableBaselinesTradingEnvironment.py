"""stable-baselines3 `VecEnv` adapter (reference: mbt_gym/gym/StableBaselinesTradingEnvironment.py:11-66).

SB3 is optional here: when it is importable the class derives from its `VecEnv`; otherwise it is a duck-typed object
with the same methods, which is what the conformance tests exercise.  Difference from the reference: the
`terminal_observation` entries are not produced by an O(N) Python loop per episode (:31-35) -- `infos` is a lazy
sequence whose items materialise `{"terminal_observation": obs[i]}` on access.
"""
import numpy as np

try:  # pragma: no cover - depends on the installation
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase  # type: ignore
except Exception:  # noqa: BLE001
    _VecEnvBase = object


class _TerminalInfos:
    """Sequence of N dicts, each {"terminal_observation": terminal_obs[i]}, built on demand."""

    def __init__(self, terminal_obs):
        self._obs = terminal_obs

    def __len__(self):
        return self._obs.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        return {"terminal_observation": self._obs[i, :]}

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def copy(self):
        return list(self)


class StableBaselinesTradingEnvironment(_VecEnvBase):
    def __init__(self, trading_env, store_terminal_observation_info=True):
        self.env = trading_env
        self.store_terminal_observation_info = store_terminal_observation_info
        self.actions = self.env.action_space.sample()
        if _VecEnvBase is not object:
            super().__init__(self.env.num_trajectories, self.env.observation_space, self.env.action_space)
        else:
            self.num_envs = self.env.num_trajectories
            self.observation_space, self.action_space = self.env.observation_space, self.env.action_space

    def reset(self):
        return self.env.reset()

    def step_async(self, actions):
        self.actions = actions

    def step_wait(self):
        obs, rewards, dones, infos = self.env.step(self.actions)
        if dones.min():
            if self.store_terminal_observation_info:
                infos = _TerminalInfos(np.array(obs, copy=True))
            obs = self.env.reset()  # SB3 convention: auto-reset, return the first observation of the new episode
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
