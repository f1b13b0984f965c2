"""TEST INFRASTRUCTURE: ctypes loader for the CPU oracle (oracle/mbt_oracle.c).

May be imported ONLY by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from mbt_gym_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmbt_oracle.so")
_lib = None


def build(force=False):
    """Compile the oracle with gcc (seconds).  Building the checker is not using it."""
    if force or not os.path.exists(_SO) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in ("mbt_oracle.c", "mbt_oracle_impl.h", "../include/mbt_math.h", "../include/mbt_philox.h",
                  "../include/mbt_b200.h")
    ):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u8p, i64, f64 = C.c_void_p, C.POINTER(C.c_uint8), C.c_int64, C.c_double
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.POINTER(_abi.mbt_config)]
        L.orc_destroy.argtypes = [vp]
        L.orc_seed.argtypes = [vp, C.c_uint64]
        L.orc_reset.argtypes = [vp, C.POINTER(_abi.mbt_reset_args), vp]
        L.orc_step.argtypes = [vp, vp, vp, vp, u8p]
        L.orc_step_draws.argtypes = [vp, vp, vp, vp, vp, vp, u8p]
        L.orc_draws.argtypes = [C.c_int, C.c_uint64, i64, i64, i64, vp, vp]
        L.orc_draws2.argtypes = [C.c_int, C.c_uint64, i64, i64, i64, vp]
        L.orc_q0_draws.argtypes = [C.c_uint64, i64, i64, i64, i64, i64, vp]
        L.orc_get_state.argtypes = [vp, vp]
        L.orc_set_state.argtypes = [vp, vp]
        L.orc_get_clock.argtypes = [vp, C.POINTER(f64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
        L.orc_reward_eval.argtypes = [C.POINTER(_abi.mbt_config), i64, vp, vp, vp, C.c_int, f64, f64, vp]
        L.orc_config_dims.argtypes = [C.POINTER(_abi.mbt_config)] + [C.POINTER(C.c_int32)] * 3
        L.orc_config_obs_out_dim.argtypes = [C.POINTER(_abi.mbt_config)]
        L.orc_philox.argtypes = [vp, vp, vp]
        for name in ("exp_f32", "log_f32", "exp_f64", "log_f64", "normal_f32", "normal_f64"):
            getattr(L, "orc_vec_" + name).argtypes = [i64, vp, vp]
        L.orc_vec_pow_f64.argtypes = [i64, vp, f64, vp]
        L.orc_count_fill_filter_mismatch.argtypes = [i64, vp, vp, C.POINTER(f64)]
        L.orc_count_fill_filter_mismatch.restype = i64
        L.orc_count_div_rcp_mismatch_f64.argtypes = [i64, vp, vp]
        L.orc_count_div_rcp_mismatch_f64.restype = i64
        L.orc_count_div_rcp_mismatch_f32.argtypes = [i64, vp, vp]
        L.orc_count_div_rcp_mismatch_f32.restype = i64
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def dims(cfg):
    a, d, s = C.c_int32(), C.c_int32(), C.c_int32()
    if lib().orc_config_dims(C.byref(cfg), C.byref(a), C.byref(d), C.byref(s)) != 0:
        raise ValueError("unsupported config")
    return a.value, d.value, s.value


class OracleEnv:
    """The oracle behind the same call shapes as the C ABI (reset / step / get_state)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.N = int(cfg.num_trajectories)
        self.A, self.D, _ = dims(cfg)
        self.Dout = lib().orc_config_obs_out_dim(C.byref(cfg))
        self.dtype = np.float64 if cfg.precision == _abi.MBT_F64 else np.float32
        self._h = lib().orc_create(C.byref(cfg))
        if not self._h:
            raise ValueError("orc_create failed")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def seed(self, seed):
        lib().orc_seed(self._h, C.c_uint64(int(seed)))

    def reset(self, args=None):
        obs = np.empty((self.N, self.Dout), self.dtype)
        lib().orc_reset(self._h, None if args is None else C.byref(args), _ptr(obs))
        return obs

    def _check_actions(self, actions):
        a = np.ascontiguousarray(actions, dtype=self.dtype)
        assert a.shape == (self.N, self.A), (a.shape, (self.N, self.A))
        return a

    def step(self, actions):
        a = self._check_actions(actions)
        obs = np.empty((self.N, self.Dout), self.dtype)
        rew = np.empty((self.N,), self.dtype)
        done = C.c_uint8(0)
        lib().orc_step(self._h, _ptr(a), _ptr(obs), _ptr(rew), C.byref(done))
        return obs, rew, bool(done.value)

    def step_draws(self, actions, u, z):
        a = self._check_actions(actions)
        u = np.ascontiguousarray(u, dtype=self.dtype)
        z = np.ascontiguousarray(z, dtype=self.dtype).reshape(-1)
        assert u.shape == (self.N, 4) and z.shape == (self.N,)
        obs = np.empty((self.N, self.Dout), self.dtype)
        rew = np.empty((self.N,), self.dtype)
        done = C.c_uint8(0)
        lib().orc_step_draws(self._h, _ptr(a), _ptr(u), _ptr(z), _ptr(obs), _ptr(rew), C.byref(done))
        return obs, rew, bool(done.value)

    @property
    def state(self):
        s = np.empty((self.N, self.D), self.dtype)
        lib().orc_get_state(self._h, _ptr(s))
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, dtype=self.dtype)
        assert s.shape == (self.N, self.D)
        lib().orc_set_state(self._h, _ptr(s))

    def clock(self):
        t, k, n, e, c = C.c_double(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        lib().orc_get_clock(self._h, C.byref(t), C.byref(k), C.byref(n), C.byref(e), C.byref(c))
        return dict(time=t.value, k=k.value, n_step=n.value, n_episode=e.value, clipped=c.value)


def draws(precision, seed, traj_offset, N, n_step):
    """(u (N,4), z (N,)) that env-step number `n_step` consumes under the Philox draw contract."""
    dt = np.float64 if precision == _abi.MBT_F64 else np.float32
    u = np.empty((N, 4), dt)
    z = np.empty((N,), dt)
    lib().orc_draws(precision, C.c_uint64(int(seed)), traj_offset, N, n_step, _ptr(u), _ptr(z))
    return u, z


def draws2(precision, seed, traj_offset, N, n_step):
    """z2 (N,): the second normal of env-step `n_step` (Heston variance) under the Philox draw contract."""
    z2 = np.empty((N,), np.float64 if precision == _abi.MBT_F64 else np.float32)
    lib().orc_draws2(precision, C.c_uint64(int(seed)), traj_offset, N, n_step, _ptr(z2))
    return z2


def q0_draws(seed, traj_offset, N, n_episode, lo, hi):
    out = np.empty((N,), np.int64)
    lib().orc_q0_draws(C.c_uint64(int(seed)), traj_offset, N, n_episode, lo, hi, _ptr(out))
    return out


def reward_eval(cfg, cur, act, nxt, is_terminal=False, q0=0.0, episode_length=1.0):
    dt = np.float64 if cfg.precision == _abi.MBT_F64 else np.float32
    cur = np.ascontiguousarray(cur, dt)
    act = np.ascontiguousarray(act, dt)
    nxt = np.ascontiguousarray(nxt, dt)
    out = np.empty((cur.shape[0],), dt)
    lib().orc_reward_eval(C.byref(cfg), cur.shape[0], _ptr(cur), _ptr(act), _ptr(nxt), int(bool(is_terminal)),
                          float(q0), float(episode_length), _ptr(out))
    return out


def div_rcp_mismatches(a, b):
    """How many pairs the kernels' division-through-reciprocal (mbt_div_rcp_*) gets different from a / b (bitwise)."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape
    fn = lib().orc_count_div_rcp_mismatch_f64 if a.dtype == np.float64 else lib().orc_count_div_rcp_mismatch_f32
    return int(fn(a.size, _ptr(a), _ptr(b)))


def fill_filter_check(k, x):
    """(mismatches of the float-filtered fill decision against its float64 definition, worst relative error of the float
    estimate of exp(x) * 2^24 over -16 <= x <= 0)."""
    k, x = np.ascontiguousarray(k, np.uint32), np.ascontiguousarray(x, np.float64)
    worst = C.c_double()
    bad = lib().orc_count_fill_filter_mismatch(k.size, _ptr(k), _ptr(x), C.byref(worst))
    return int(bad), worst.value


def philox(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    out = np.empty(4, np.uint32)
    lib().orc_philox(_ptr(c), _ptr(k), _ptr(out))
    return out


def vec(name, x, *extra):
    x = np.ascontiguousarray(x)
    out_dt = {"exp_f32": np.float32, "log_f32": np.float32, "exp_f64": np.float64, "log_f64": np.float64,
              "normal_f32": np.float32, "normal_f64": np.float64, "pow_f64": np.float64}[name]
    out = np.empty(x.shape, out_dt)
    getattr(lib(), "orc_vec_" + name)(x.size, _ptr(x), *extra, _ptr(out))
    return out
