"""`TradingEnvironment` -- the reference's vectorised gym.Env surface over the B200 step kernel.

Host-side mirror of mbt_gym/gym/TradingEnvironment.py: same constructor keywords, attributes, spaces,
`reset() -> obs (N,D)`, `step(action (N,A)) -> (obs, rewards, dones, infos)`, `seed`, `state`, normalisation helpers.
What differs is where the work happens: every `reset()` / `step()` is ONE call through the C ABI
(include/mbt_b200.h) into a CUDA kernel; this module contains no simulation arithmetic and there is no CPU fallback.

Extra keyword-only arguments (all optional, defaults keep the reference behaviour):
    precision     "float64" (default, the reference's dtype) or "float32"
    io_dtype      dtype of the action / observation / reward arrays exchanged with the caller: default = `precision`;
                  `np.float32` with float64 arithmetic keeps the dynamics reference-exact and halves the host traffic
    device        CUDA device index
    traj_offset   global id of trajectory 0 (multi-GPU sharding: RNG counters use global ids)
    copy_outputs  False (default): `step()`/`reset()` return arrays backed by a POOL of page-locked buffers that are
                  handed out again only when nothing references them any more (`_lib.PinnedPool`) -- like the
                  reference's fresh `.copy()` arrays they are never overwritten while the caller holds them (or any
                  view / `torch.from_numpy` of them), and their memory outlives `close()`; the device-to-host copy is
                  a direct DMA.  True: ordinary `np.empty` arrays every call (staged through pinned memory).
Actions may also be a CUDA torch tensor: then observations and rewards come back as CUDA tensors on the same device
and nothing crosses PCIe (the zero-copy path for on-device policies).
"""
import os
from collections import OrderedDict
from copy import copy

import numpy as np

from .. import _abi, _lib, _track
from ..rewards.RewardFunctions import PnL, RewardFunction
from ..spaces import Box
from ..stochastic_processes.arrival_models import PoissonArrivalModel
from ..stochastic_processes.fill_probability_models import ExponentialFillFunction
from ..stochastic_processes.midprice_models import BrownianMotionMidpriceModel
from .index_names import CASH_INDEX, INVENTORY_INDEX, TIME_INDEX  # noqa: F401  (re-exported like the reference)
from .ModelDynamics import LimitOrderModelDynamics, ModelDynamics

try:  # optional: make isinstance(env, gym.Env) true when a gym flavour is installed
    from gymnasium import Env as _EnvBase  # type: ignore
except Exception:  # noqa: BLE001
    try:
        from gym import Env as _EnvBase  # type: ignore
    except Exception:  # noqa: BLE001
        _EnvBase = object

_PROCESS_ORDER = ("midprice_model", "arrival_model", "fill_probability_model", "price_impact_model")


class _EmptyInfos(list):
    """The reference returns one pre-built list of N empty dicts every step (TradingEnvironment.py:320-321)."""


class TradingEnvironment(_track.Tracked, _EnvBase):
    metadata = {"render.modes": ["human"]}

    def __init__(self, terminal_time=1.0, n_steps=20 * 10, reward_function=None, model_dynamics=None, initial_cash=0.0,
                 initial_inventory=0, max_inventory=10_000, max_cash=None, max_stock_price=None, start_time=0.0,
                 info_calculator=None, seed=None, num_trajectories=1, normalise_action_space=True,
                 normalise_observation_space=True, normalise_rewards=False, *, precision="float64", device=0,
                 traj_offset=0, copy_outputs=False, obs_columns=None, io_dtype=None):
        super().__init__()
        self._native = None
        self._native_cfg_bytes = None
        self._started = False
        self._episode_open = False  # between reset() and the step that returned done
        self._seen_version = -1     # _track.version the device handle was last validated against
        self._pool = None
        self._act_scratch = None  # page-locked (N, A) buffer for actions that need a dtype / layout conversion
        self._infos = None
        self.precision = {"float64": _abi.MBT_F64, "f64": _abi.MBT_F64, "float32": _abi.MBT_F32, "f32": _abi.MBT_F32}[str(precision)]
        self.dtype = np.dtype(np.float64 if self.precision == _abi.MBT_F64 else np.float32)
        # dtype of the arrays step()/reset() exchange with the caller: the arithmetic dtype, or float32 over float64
        # arithmetic (SB3's buffers are float32; the PCIe-bound host path then moves half the bytes)
        self.io_dtype = self.dtype if io_dtype is None else np.dtype(io_dtype)
        if self.io_dtype not in (self.dtype, np.dtype(np.float32)):
            raise ValueError("io_dtype must be the environment's dtype or float32")
        self.device, self.traj_offset, self.copy_outputs = int(device), int(traj_offset), bool(copy_outputs)
        self._obs_columns = None

        self.terminal_time = terminal_time
        self.n_steps = n_steps
        self._step_size = self.terminal_time / self.n_steps
        self.reward_function = reward_function or PnL()
        self.model_dynamics = model_dynamics or LimitOrderModelDynamics(
            midprice_model=BrownianMotionMidpriceModel(step_size=self._step_size, num_trajectories=num_trajectories, seed=seed),
            arrival_model=PoissonArrivalModel(intensity=np.array([100, 100]), step_size=self._step_size,
                                              num_trajectories=num_trajectories, seed=seed),
            fill_probability_model=ExponentialFillFunction(step_size=self._step_size, num_trajectories=num_trajectories, seed=seed),
            num_trajectories=num_trajectories, seed=seed)
        assert isinstance(self.model_dynamics, ModelDynamics), "model_dynamics must be a mbt_gym_b200 ModelDynamics"
        assert isinstance(self.reward_function, RewardFunction), "reward_function must be a mbt_gym_b200 RewardFunction"
        self.model_dynamics._env = self
        self.stochastic_processes = self._get_stochastic_processes()
        self.stochastic_process_indices = self._get_stochastic_process_indices()
        self.num_trajectories = num_trajectories
        self.initial_cash = initial_cash
        self.initial_inventory = initial_inventory
        self.max_inventory = max_inventory
        self._key = int.from_bytes(os.urandom(8), "little")  # "unseeded" = fresh entropy, like default_rng(None)
        if seed:  # the reference ignores seed=0 the same way (TradingEnvironment.py:70)
            self.seed(seed)
        self.rng = np.random.default_rng(seed)
        self.start_time = start_time
        self.max_stock_price = max_stock_price or self.model_dynamics.midprice_model.max_value[0, 0]
        self.max_cash = max_cash or self._get_max_cash()
        self.info_calculator = info_calculator
        self.observation_space = self._get_observation_space()
        self.action_space = self.model_dynamics.get_action_space()
        self.normalise_action_space_ = normalise_action_space
        self.normalise_observation_space_ = normalise_observation_space
        self.normalise_rewards_ = normalise_rewards
        self.reward_scaling = 1.0
        if self.normalise_observation_space_:
            self.original_observation_space = copy(self.observation_space)
            self.observation_space = self._unit_box(self.observation_space)
        if self.normalise_action_space_:
            self.original_action_space = copy(self.action_space)
            self.action_space = self._unit_box(self.action_space)
        self._full_observation_space = self.observation_space
        if self.normalise_rewards_:
            assert isinstance(self.model_dynamics.arrival_model, PoissonArrivalModel) and isinstance(
                self.model_dynamics.fill_probability_model, ExponentialFillFunction
            ), "Arrival model must be Poisson and fill probability model must be exponential to scale rewards"
            self.reward_scaling = 1 / self._get_inventory_neutral_rewards()
        if obs_columns is not None:
            self.select_observation_columns(obs_columns)

    def select_observation_columns(self, columns):
        """Emit only these observation columns (in increasing order) -- `ReduceStateSizeWrapper` fused into the kernel's
        observation store (mbt_gym/gym/wrappers.py:10-43): no host-side fancy-index copy and proportionally fewer D2H
        bytes.  `None` restores all columns.  Takes effect at the next `reset()` (the device handle is rebuilt from the
        Python attributes there).  `env.state` and the agents' `to_policy` forms are unaffected."""
        if columns is None:
            self._obs_columns = None
        else:
            cols = sorted({int(c) for c in columns})
            d = self._full_observation_space.shape[0]
            if not cols or cols[0] < 0 or cols[-1] >= d:
                raise ValueError(f"observation columns must be within 0..{d - 1}")
            self._obs_columns = cols
        full = self._full_observation_space
        idx = self._obs_columns if self._obs_columns is not None else list(range(full.shape[0]))
        self.observation_space = Box(low=full.low[idx], high=full.high[idx], dtype=full.low.dtype)

    # ------------------------------------------------------------------ reference surface: hot path
    def reset(self):
        """Start an episode; returns the (N, D) observation  (TradingEnvironment.py:96-101)."""
        native = self._ensure_native(at_reset=True)
        args = _abi.mbt_reset_args()
        args.start_time = float(self._get_start_time())
        self._fill_initial_inventory(args)
        out = self._out_buffers(rewards=False)[0]
        native.reset(out, args)
        self._after_reset(args)
        return out

    def _after_reset(self, args):
        self._started = True
        self._episode_open = True
        if getattr(self.reward_function, "terminal_time", None) is not None:
            # what reward_function.reset(initial_state) records (RewardFunctions.py:72-74,111-113), without a D2H copy
            self.reward_function.episode_length = self.reward_function.terminal_time - args.start_time
            self.reward_function.initial_inventory = (
                None if args.q0_mode == _abi.MBT_Q0_UNIFORM_INT else
                self._q0_values if args.q0_mode == _abi.MBT_Q0_PER_TRAJ else args.q0_const)
        self._seen_version = _track.version[0]  # the two assignments above are ours, not the caller's

    def reset_device(self, out=None):
        """`reset()` that stays on the device: the first observation as a CUDA torch tensor (written into `out` when
        given), enqueued on torch's current stream -- no host round-trip, so it can be captured into a CUDA graph
        together with `step(cuda_tensor)` calls (see `fold_counters`)."""
        import torch

        native = self._ensure_native(at_reset=True)
        args = _abi.mbt_reset_args()
        args.start_time = float(self._get_start_time())
        self._fill_initial_inventory(args)
        tdt = torch.float64 if self.io_dtype == np.float64 else torch.float32
        dev = torch.device("cuda", self.device)
        if out is None:
            out = torch.empty((self.num_trajectories, native.Dout), dtype=tdt, device=dev)
        elif out.dtype != tdt or not out.is_contiguous() or tuple(out.shape) != (self.num_trajectories, native.Dout):
            raise ValueError(f"out must be a contiguous {tdt} tensor of shape ({self.num_trajectories}, {native.Dout})")
        native.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        native.reset(out, args, mem=_abi.MBT_MEM_DEVICE)
        self._after_reset(args)
        return out

    def prepare_capture(self):
        """Call BEFORE `torch.cuda.graph(...)` when the environment was already used (warm-up `reset_device()` / steps):
        the random-number counters consumed so far move to the device, so the launches baked into the graph count from
        zero and later eager calls and graph replays interleave without reusing or skipping a draw.
        C ABI: `mbt_prepare_capture` (include/mbt_b200.h)."""
        self._ensure_native(at_reset=True).prepare_capture()

    def fold_counters(self):
        """Call LAST inside a `torch.cuda.graph` capture that spans whole episodes (`reset_device`, then policy and
        `step(cuda_tensor)` until done): the random-number counters move to the device, so every `graph.replay()` is a
        NEW episode -- bit-identical to stepping the same episodes eagerly -- instead of a repeat of the captured one.
        Bind the environment to the capture stream first (a warm-up `reset_device()` under `torch.cuda.stream(s)`), call
        `prepare_capture()`, then capture with `torch.cuda.graph(g, stream=s)`.  C ABI: `mbt_fold_counters`."""
        self._native.fold_counters()

    def step(self, action):
        """One env-step for all trajectories  (TradingEnvironment.py:103-110).
        action (N, A) -> observations (N, D), rewards (N,), dones (N,) bool, infos."""
        native = self._native
        if native is None or not self._started:
            raise RuntimeError("step() called before reset()")
        if _track.version[0] != self._seen_version:
            # some attribute of an environment / model object was assigned since the handle was validated: re-flatten; an
            # edit of THIS environment is applied in place, like the reference applies it at its next step
            native = self._ensure_native()
        if hasattr(action, "is_cuda") and action.is_cuda:
            return self._step_device(native, action)
        if type(action) is np.ndarray and action.dtype == self.io_dtype and action.flags.c_contiguous:
            a = action
        else:
            a = np.asarray(action)
            if a.shape == (self.num_trajectories, native.A) and a.dtype.kind in "fiu":
                # a conversion is needed anyway (e.g. SB3's float32 actions into a float64 environment): convert straight
                # into page-locked memory, so the upload is a direct DMA instead of a second staging copy
                scratch = self._act_scratch
                if scratch is None or scratch.shape != a.shape or scratch.dtype != self.io_dtype:
                    scratch = self._act_scratch = _lib.PinnedArray(a.shape, self.io_dtype, self.device).array
                np.copyto(scratch, a, casting="unsafe")
                a = scratch
            else:
                a = np.ascontiguousarray(action, dtype=self.io_dtype)
        if a.shape != (self.num_trajectories, native.A):
            if a.size == self.num_trajectories * native.A and self.num_trajectories == 1:
                a = a.reshape(1, native.A)
            else:
                raise ValueError(f"action must have shape ({self.num_trajectories}, {native.A}); got {np.shape(action)}")
        obs, rew = self._out_buffers()
        done = native.step(a, obs, rew)
        if done:
            self._episode_open = False
        infos = self._calculate_infos()
        return obs, rew, self._dones(done), infos

    def _step_device(self, native, action):
        import torch

        tdt = torch.float64 if self.io_dtype == np.float64 else torch.float32
        if action.dtype != tdt or not action.is_contiguous():
            action = action.to(tdt).contiguous()
        if tuple(action.shape) != (self.num_trajectories, native.A):
            raise ValueError(f"action must have shape ({self.num_trajectories}, {native.A}); got {tuple(action.shape)}")
        if action.device.index != self.device:
            raise ValueError(f"action is on cuda:{action.device.index}, the environment on cuda:{self.device}")
        native.set_stream(torch.cuda.current_stream(action.device).cuda_stream)
        obs = torch.empty((self.num_trajectories, native.Dout), dtype=tdt, device=action.device)
        rew = torch.empty((self.num_trajectories,), dtype=tdt, device=action.device)
        done = native.step(action, obs, rew, mem=_abi.MBT_MEM_DEVICE)
        if done:
            self._episode_open = False
        return obs, rew, self._dones(done), self._calculate_infos()

    # ------------------------------------------------------------------ reference surface: helpers
    def normalise_observation(self, obs, inverse=False):
        if not self.normalise_observation_space_:
            return obs
        if inverse:
            return (obs + 1) * self._gradient_obs_norm + self._intercept_obs_norm
        return (obs - self._intercept_obs_norm) / self._gradient_obs_norm - 1

    def normalise_action(self, action, inverse=False):
        if not self.normalise_action_space_:
            return action
        if inverse:
            return (action + 1) * self._gradient_action_norm + self._intercept_action_norm
        return (action - self._intercept_action_norm) / self._gradient_action_norm - 1

    def normalise_rewards(self, rewards):
        return self.reward_scaling * rewards if self.normalise_rewards_ else rewards

    def seed(self, seed=None):
        """Re-key the counter-based RNG (TradingEnvironment.py:345-348).  One Philox key replaces the reference's
        per-process PCG64 generators; `seed + i + 1` is still recorded on the processes for introspection."""
        self.rng = np.random.default_rng(seed)
        self._key = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        for i, process in enumerate(self.stochastic_processes.values()):
            process.seed(None if seed is None else seed + i + 1)
        if self._native is not None:
            self._native.seed(self._key)

    @property
    def state(self):
        """Raw (un-normalised) (N, D) state matrix, copied from the device  (TradingEnvironment.py:142-144)."""
        return self._get_state()

    @property
    def initial_state(self):
        """(N, D) initial state the NEXT reset would produce for a constant initial inventory (host-side view)."""
        s = np.repeat(np.array([[self.initial_cash, 0, 0.0]]), self.num_trajectories, axis=0)
        s[:, TIME_INDEX] = self._get_start_time()
        s[:, INVENTORY_INDEX] = self._get_initial_inventories()
        for process in self.stochastic_processes.values():
            s = np.append(s, process.initial_vector_state, axis=1)
        return s

    @property
    def is_at_max_inventory(self):
        return self.state[:, INVENTORY_INDEX] >= self.max_inventory

    @property
    def is_at_min_inventory(self):
        return self.state[:, INVENTORY_INDEX] <= -self.max_inventory

    @property
    def step_size(self):
        return self._step_size

    @step_size.setter
    def step_size(self, step_size):
        self._step_size = step_size
        for process in self.stochastic_processes.values():
            if process.step_size != step_size:
                process.step_size = step_size
        if hasattr(self.reward_function, "step_size"):
            self.reward_function.step_size = step_size

    @property
    def num_trajectories(self):
        return self._num_trajectories

    @num_trajectories.setter
    def num_trajectories(self, num_trajectories):
        self._num_trajectories = num_trajectories
        for process in self.stochastic_processes.values():
            if process.num_trajectories != num_trajectories:
                process.num_trajectories = num_trajectories
        self.model_dynamics.num_trajectories = num_trajectories
        self._infos = None

    def save_checkpoint(self):
        """The whole environment (device state + clock + RNG counters) as a uint8 array; see `load_checkpoint`."""
        if self._native is None or not self._started:
            raise RuntimeError("nothing to checkpoint before reset()")
        return self._native.checkpoint()

    def load_checkpoint(self, blob):
        """Resume exactly where `save_checkpoint` was taken (same trajectories, same future random draws)."""
        native = self._ensure_native(at_reset=True)
        native.restore(blob)
        self._key = native.get_seed()  # the checkpoint carries its own key; a later seed() replaces it
        self._started = True
        clk = native.clock()
        self._episode_open = clk["time"] < self.terminal_time - self.step_size / 2

    def pinned_actions(self):
        """A page-locked (N, A) array in the environment's dtype: fill it and pass it to `step()` and the action copy
        is one direct DMA (any other host array is first staged through the handle's own pinned buffer)."""
        native = self._ensure_native()
        return _lib.PinnedArray((self.num_trajectories, native.A), self.io_dtype, self.device).array

    def close(self):
        if self._native is not None:
            self._native.close()
            self._native = None
        self._pool = None  # arrays the caller still holds keep their own page-locked blocks alive

    # ------------------------------------------------------------------ fused rollout (agents on the device)
    def rollout_summary(self, policy, return_trajectory_stats=False):
        """Run the rest of the episode on the device with an on-device policy (`agent.to_policy(env)`), state in
        registers; returns the `mbt_summary` moments (and per-trajectory returns / terminal inventories)."""
        native = self._ensure_native()
        if not self._started:
            raise RuntimeError("rollout_summary() called before reset()")
        if return_trajectory_stats:
            ret = np.empty((self.num_trajectories,), self.dtype)
            q = np.empty((self.num_trajectories,), self.dtype)
            return native.rollout(policy, ret, q), ret, q
        return native.rollout(policy)

    def inventory_histogram(self, lo=None, hi=None, group_sum=False):
        """Distribution of the current (after an episode: terminal) inventories, binned on the device: returns
        `(inventories, counts)` for the integer bins lo..hi (default: -max_inventory..max_inventory, at most 4096 bins)
        plus `(below, above)` -- what the reference's results plot histograms on the host (helpers/plotting.py:94-110).
        `group_sum=True` sums the counts of all ranks of the handle's library group (sharding.create_group)."""
        native = self._ensure_native()
        if not self._started:
            raise RuntimeError("inventory_histogram() called before reset()")
        q = int(min(self.max_inventory, 2047))
        lo = -q if lo is None else int(lo)
        hi = q if hi is None else int(hi)
        c = native.inventory_histogram(lo, hi, group_sum)
        return np.arange(lo, hi + 1), c[1:-1], (int(c[0]), int(c[-1]))

    # ------------------------------------------------------------------ internals
    @property
    def _intercept_obs_norm(self):
        return self.original_observation_space.low

    @property
    def _gradient_obs_norm(self):
        return (self.original_observation_space.high - self.original_observation_space.low) / 2

    @property
    def _intercept_action_norm(self):
        return self.original_action_space.low

    @property
    def _gradient_action_norm(self):
        return (self.original_action_space.high - self.original_action_space.low) / 2

    def _get_stochastic_processes(self):
        procs = OrderedDict()
        for name in _PROCESS_ORDER:  # fixed order midprice -> arrival -> fill -> impact (TradingEnvironment.py:303-309)
            p = getattr(self.model_dynamics, name)
            if p is not None:
                procs[name] = p
        return procs

    def _get_stochastic_process_indices(self):
        indices, count = OrderedDict(), 3
        for name, p in self.stochastic_processes.items():
            d = int(p.initial_vector_state.shape[1])
            indices[name] = (count, count + d)
            p._env, p._columns = self, (count, count + d)
            count += d
        return indices

    def _get_max_cash(self):
        return self.n_steps * self.max_stock_price

    def _get_observation_space(self):
        low = np.array([-self.max_cash, -self.max_inventory, 0])
        high = np.array([self.max_cash, self.max_inventory, self.terminal_time])
        for p in self.stochastic_processes.values():
            low, high = np.append(low, p.min_value), np.append(high, p.max_value)
        return Box(low=np.float32(low), high=np.float32(high))

    @staticmethod
    def _unit_box(space):
        return Box(low=-np.ones_like(space.low, dtype=np.float32), high=np.ones_like(space.high, dtype=np.float32))

    def _get_start_time(self):
        if isinstance(self.start_time, (float, int)):
            t = self.start_time
        elif callable(self.start_time):
            t = self.start_time()
        else:
            raise NotImplementedError
        assert (t >= 0.0) and (t < self.terminal_time), "Start time is not within (0, env.terminal_time)."
        return np.round(t / self.step_size) * self.step_size

    def _get_initial_inventories(self):
        q0 = self.initial_inventory
        if isinstance(q0, tuple) and len(q0) == 2:
            return self.rng.integers(*q0, size=self.num_trajectories)
        if isinstance(q0, int):
            return q0 * np.ones((self.num_trajectories,))
        if callable(q0):
            v = q0()
            return int(np.round(v)) if self.model_dynamics.round_initial_inventory else v
        raise Exception("Initial inventory must be a tuple of length 2 or an int.")

    def _fill_initial_inventory(self, args):
        q0 = self.initial_inventory
        if isinstance(q0, tuple) and len(q0) == 2:  # rng.integers(lo, hi) per trajectory, drawn on the device
            args.q0_mode, args.q0_lo, args.q0_hi = _abi.MBT_Q0_UNIFORM_INT, int(q0[0]), int(q0[1])
        elif isinstance(q0, (int, np.integer)):
            args.q0_mode, args.q0_const = _abi.MBT_Q0_CONST, float(q0)
        elif callable(q0):
            v = q0()
            if self.model_dynamics.round_initial_inventory:
                v = int(np.round(v))  # (fails for arrays of more than one element, like the reference's :277-278)
            if np.ndim(v) == 0 or np.size(v) == 1:
                args.q0_mode, args.q0_const = _abi.MBT_Q0_CONST, float(np.asarray(v, float).reshape(-1)[0])
            else:
                # the reference assigns whatever the callable returns to the inventory column (:137,275-279): an
                # (N,) array gives every trajectory its own initial inventory
                vals = np.ascontiguousarray(np.broadcast_to(np.asarray(v, np.float64), (self.num_trajectories,)))
                self._q0_values = vals  # kept alive for the call (host pointer in the struct) and for reward_function
                args.q0_mode, args.q0_values = _abi.MBT_Q0_PER_TRAJ, vals.ctypes.data
        else:
            raise Exception("Initial inventory must be a tuple of length 2 or an int.")

    def _build_config(self):
        cfg = _abi.new_config(
            precision=self.precision, num_trajectories=int(self.num_trajectories), traj_offset=self.traj_offset,
            n_steps=int(self.n_steps), terminal_time=float(self.terminal_time), step_size=float(self.step_size),
            initial_cash=float(self.initial_cash), max_inventory=float(self.max_inventory), max_cash=float(self.max_cash),
            rew_terminal_time=float(self.terminal_time))
        st = self.start_time if isinstance(self.start_time, (float, int)) else 0.0
        cfg.start_time = float(np.round(st / self.step_size) * self.step_size)
        q0 = self.initial_inventory
        if isinstance(q0, tuple) and len(q0) == 2:
            cfg.q0_mode, cfg.q0_lo, cfg.q0_hi = _abi.MBT_Q0_UNIFORM_INT, int(q0[0]), int(q0[1])
        elif isinstance(q0, (int, np.integer)):
            cfg.q0_const = float(q0)
        self.model_dynamics._flatten(cfg)
        for p in self.stochastic_processes.values():
            p._flatten(cfg)
        self.reward_function._flatten(cfg)
        cfg.normalise_action = int(bool(self.normalise_action_space_))
        cfg.normalise_obs = int(bool(self.normalise_observation_space_))
        cfg.normalise_rewards = int(bool(self.normalise_rewards_))
        cfg.reward_scaling = float(self.reward_scaling)
        cfg.obs_select = sum(1 << c for c in self._obs_columns) if self._obs_columns else 0
        cfg.io_precision = _abi.MBT_IO_F32 if (self.io_dtype == np.float32 and self.dtype == np.float64) else _abi.MBT_IO_SAME
        if self.normalise_action_space_:
            lo, gr = self._intercept_action_norm, self._gradient_action_norm
            for i in range(lo.shape[0]):
                cfg.act_low[i], cfg.act_grad[i] = float(lo[i]), float(gr[i])
        if self.normalise_observation_space_:
            lo, gr = self._intercept_obs_norm, self._gradient_obs_norm
            for i in range(lo.shape[0]):
                cfg.obs_low[i], cfg.obs_grad[i] = float(lo[i]), float(gr[i])
        return cfg

    def _ensure_native(self, at_reset=False):
        """The device handle for the CURRENT Python attributes.  An edited attribute (setters, model parameters) is
        applied to the live handle in place (`mbt_reconfigure`: state, clock and random streams continue, as they do in
        the reference); an edit that changes the shape of the device state (num_trajectories, dtype, another set of model
        state columns) needs a new handle, which is only possible between episodes -- it inherits the key and the draw
        counters of the old one, so the random streams continue instead of restarting."""
        cfg = self._build_config()
        raw = bytes(cfg)
        if self._native is None:
            self._native = _lib.NativeEnv(cfg, device=self.device)
            if self._key is not None:
                self._native.seed(self._key)
            self._started = self._episode_open = False
        elif raw != self._native_cfg_bytes:
            try:
                self._native.reconfigure(cfg)
            except _lib.MbtError as err:
                if err.code != _abi.MBT_E_STATE:
                    raise
                if self._episode_open and not at_reset:
                    raise RuntimeError(
                        "an attribute that changes the shape of the device state (num_trajectories, precision, the set of "
                        "model state columns) was edited in the middle of an episode; finish the episode or call reset()"
                    ) from err
                old = self._native
                clk, key = old.clock(), old.get_seed()
                old.close()
                self._native = _lib.NativeEnv(cfg, device=self.device)
                self._native.seed(key)
                self._native.set_counters(clk["n_step"], clk["n_episode"])  # the Philox streams go on, they do not restart
                self._started = self._episode_open = False
        self._native_cfg_bytes = raw
        self._seen_version = _track.version[0]
        return self._native

    def _out_buffers(self, rewards=True):
        n, d = self.num_trajectories, self._native.Dout
        obs = rew = None
        if not self.copy_outputs:
            if self._pool is None:
                self._pool = _lib.PinnedPool(self.device)
            obs = self._pool.get((n, d), self.io_dtype)
            rew = self._pool.get((n,), self.io_dtype) if rewards else None
        if obs is None:
            obs = np.empty((n, d), self.io_dtype)
        if rew is None and rewards:
            rew = np.empty((n,), self.io_dtype)
        return obs, rew

    def _dones(self, done):
        # a fresh array every step, like the reference's (:218-220)
        return np.ones((self.num_trajectories,), bool) if done else np.zeros((self.num_trajectories,), bool)

    def _calculate_infos(self):
        if self.info_calculator is not None:
            raise NotImplementedError("info calculators are not part of the fused step (the reference's are dead code, "
                                      "see SURVEY.md section 2 row 15)")
        if self._infos is None:
            self._infos = _EmptyInfos({} for _ in range(self.num_trajectories)) if self.num_trajectories > 1 else {}
        return self._infos

    def _get_state(self):
        native = self._ensure_native()
        if not self._started:
            return self.initial_state.astype(self.dtype)
        return native.get_state()

    def _set_state(self, value):
        native = self._ensure_native()
        native.set_state(np.asarray(value, dtype=self.dtype))
        self._started = True

    def _get_inventory_neutral_rewards(self, num_total_trajectories=100_000):
        """Mean episode reward of the fixed action 1/kappa over 100 000 trajectories (TradingEnvironment.py:329-343),
        computed by the fused rollout kernel on a temporary handle."""
        fixed_action = 1 / self.model_dynamics.fill_probability_model.fill_exponent
        cfg = self._build_config()
        cfg.num_trajectories = num_total_trajectories
        cfg.start_time = 0.0
        cfg.normalise_rewards = 0
        tmp = _lib.NativeEnv(cfg, device=self.device)
        tmp.seed(self._key)
        tmp.reset()
        pol = _abi.mbt_policy()
        pol.kind = _abi.MBT_POL_FIXED
        for j in range(tmp.A):
            pol.fixed[j] = fixed_action
        s = tmp.rollout(pol)
        tmp.close()
        return s.sum_return / s.count / s.steps * self.n_steps
