"""ctypes binding of libmbt_b200.so (include/mbt_b200.h) and a thin handle class.

The product has NO CPU fallback: if the library is missing or no CUDA device is usable, the calls here raise.
"""
import ctypes as C
import os
import sys

import numpy as np

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MBT_LIB_PATH") or os.path.join(_HERE, "libmbt_b200.so")  # MBT_LIB_PATH: tuning builds only

# every symbol include/mbt_b200.h declares (checked by tests/test_abi_symbols.py)
ABI_SYMBOLS = [
    "mbt_abi_version", "mbt_last_error", "mbt_config_dims", "mbt_config_obs_out_dim", "mbt_create", "mbt_destroy", "mbt_set_stream", "mbt_sync",
    "mbt_seed", "mbt_reset", "mbt_step", "mbt_get_state", "mbt_set_state", "mbt_get_clock", "mbt_get_clip_count",
    "mbt_reward_eval", "mbt_rollout", "mbt_rollout_record", "mbt_get_launch_count", "mbt_enable_timing", "mbt_get_kernel_times",
    "mbt_host_alloc", "mbt_host_alloc_near", "mbt_host_free", "mbt_checkpoint_size", "mbt_checkpoint_save",
    "mbt_checkpoint_load", "mbt_fold_counters", "mbt_prepare_capture", "mbt_get_seed", "mbt_set_counters", "mbt_reconfigure",
    "mbt_group_unique_id", "mbt_group_create", "mbt_group_destroy", "mbt_group_info", "mbt_group_rollout", "mbt_group_summary",
    "mbt_group_wait", "mbt_get_kernel_info", "mbt_jit_precompile", "mbt_inventory_histogram",
]

_lib = None


class MbtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmbt_b200 error {code}: {msg}")
        self.code = code


def load():
    """Load libmbt_b200.so (built in-tree by `python -m mbt_gym_b200._build` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found. mbt_gym_b200 has no CPU implementation: build the CUDA library first "
            "(python -m mbt_gym_b200._build, needs nvcc)."
        )
    L = C.CDLL(LIB_PATH)
    vp, i64p = C.c_void_p, C.POINTER(C.c_int64)
    cfgp = C.POINTER(_abi.mbt_config)
    L.mbt_abi_version.restype = C.c_int
    L.mbt_last_error.restype = C.c_char_p
    L.mbt_config_dims.argtypes = [cfgp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.mbt_config_obs_out_dim.argtypes = [cfgp, C.POINTER(C.c_int32)]
    L.mbt_create.argtypes = [cfgp, C.c_int, C.POINTER(vp)]
    L.mbt_destroy.argtypes = [vp]
    L.mbt_set_stream.argtypes = [vp, vp]
    L.mbt_sync.argtypes = [vp]
    L.mbt_seed.argtypes = [vp, C.c_uint64]
    L.mbt_reset.argtypes = [vp, C.POINTER(_abi.mbt_reset_args), vp, C.c_int]
    L.mbt_step.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_uint8), C.c_int]
    L.mbt_get_state.argtypes = [vp, vp, C.c_int]
    L.mbt_set_state.argtypes = [vp, vp, C.c_int]
    L.mbt_get_clock.argtypes = [vp, C.POINTER(C.c_double), i64p, i64p, i64p]
    L.mbt_get_clip_count.argtypes = [vp, i64p]
    L.mbt_reward_eval.argtypes = [vp, C.c_int64, vp, vp, vp, C.c_int, vp, C.c_int]
    L.mbt_rollout.argtypes = [vp, C.POINTER(_abi.mbt_policy), C.POINTER(_abi.mbt_summary), vp, vp, C.c_int]
    L.mbt_rollout_record.argtypes = [vp, C.POINTER(_abi.mbt_policy), C.POINTER(_abi.mbt_summary), C.POINTER(_abi.mbt_record), C.c_int]
    L.mbt_get_launch_count.argtypes = [vp, i64p]
    L.mbt_enable_timing.argtypes = [vp, C.c_int]
    L.mbt_get_kernel_times.argtypes = [vp, C.POINTER(C.c_float), C.c_int64, i64p]
    L.mbt_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.mbt_host_alloc_near.argtypes = [C.c_size_t, C.c_int, C.POINTER(vp)]
    L.mbt_host_free.argtypes = [vp]
    L.mbt_checkpoint_size.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.mbt_checkpoint_save.argtypes = [vp, vp, C.c_size_t]
    L.mbt_checkpoint_load.argtypes = [vp, vp, C.c_size_t]
    L.mbt_fold_counters.argtypes = [vp]
    L.mbt_prepare_capture.argtypes = [vp]
    L.mbt_get_seed.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.mbt_set_counters.argtypes = [vp, C.c_int64, C.c_int64]
    L.mbt_reconfigure.argtypes = [vp, cfgp]
    L.mbt_group_unique_id.argtypes = [vp]
    L.mbt_group_create.argtypes = [vp, vp, C.c_int, C.c_int]
    L.mbt_group_destroy.argtypes = [vp]
    L.mbt_group_info.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), i64p]
    L.mbt_group_rollout.argtypes = [vp, C.POINTER(_abi.mbt_policy), C.POINTER(_abi.mbt_summary), vp, vp]
    L.mbt_group_summary.argtypes = [vp, C.POINTER(_abi.mbt_summary), C.POINTER(_abi.mbt_summary)]
    L.mbt_group_wait.argtypes = [vp, C.c_int]
    L.mbt_get_kernel_info.argtypes = [vp, C.POINTER(_abi.mbt_kernel_info)]
    L.mbt_jit_precompile.argtypes = [cfgp, C.c_int32, C.c_int32, C.c_int32]
    L.mbt_inventory_histogram.argtypes = [vp, C.c_int64, C.c_int64, i64p, C.c_int]
    if L.mbt_abi_version() != _abi.MBT_ABI_VERSION:
        raise ImportError(f"libmbt_b200.so ABI {L.mbt_abi_version()} != binding ABI {_abi.MBT_ABI_VERSION}")
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise MbtError(rc, load().mbt_last_error().decode("utf-8", "replace"))


def jit_precompile(cfg, kind=0, policy_kind=0, record=False):
    """Compile the run-time specialised kernel of `cfg` into the on-disk cache (no CUDA device needed)."""
    _check(load().mbt_jit_precompile(C.byref(cfg), int(kind), int(policy_kind), int(bool(record))))


def config_dims(cfg):
    a, d, s = C.c_int32(), C.c_int32(), C.c_int32()
    _check(load().mbt_config_dims(C.byref(cfg), C.byref(a), C.byref(d), C.byref(s)))
    return a.value, d.value, s.value


class _PinnedBlock:
    """Page-locked host memory (mbt_host_alloc_near) whose lifetime is tied to the arrays that view it: numpy arrays made
    by `array()` -- and every view, slice or `torch.from_numpy` of them -- keep a reference to the block, and the memory
    is returned to CUDA only when the last of them is gone.  Nothing a caller still holds is ever freed or reused."""

    def __init__(self, nbytes, device=0):
        self.nbytes = max(int(nbytes), 1)
        p = C.c_void_p()
        _check(load().mbt_host_alloc_near(self.nbytes, int(device), C.byref(p)))
        self._ptr = p

    def array(self, shape, dtype):
        dtype = np.dtype(dtype)
        shape = tuple(int(x) for x in np.atleast_1d(shape))
        assert int(np.prod(shape)) * dtype.itemsize <= self.nbytes
        self.__array_interface__ = {"shape": shape, "typestr": dtype.str, "data": (self._ptr.value, False), "version": 3}
        try:
            return np.asarray(self)
        finally:
            del self.__array_interface__

    def __del__(self):
        p = getattr(self, "_ptr", None)
        if p is not None and p.value:
            try:
                load().mbt_host_free(p)
            except Exception:
                pass
            self._ptr = None


def _refcount(obj):
    return sys.getrefcount(obj)


class PinnedPool:
    """Output buffers for the host path: pinned blocks that are handed out again only when NO array viewing them is alive
    (CPython reference counts), so `step()` / `reset()` results behave like the reference's fresh `.copy()` arrays -- never
    overwritten behind the caller's back -- while the steady state (a caller that consumes each step's arrays before the
    next few steps, as SB3's rollout buffer and `generate_trajectory` do) reuses the same few page-locked buffers and the
    device-to-host copy stays a direct DMA.  A caller that keeps everything makes the pool grow up to `max_bytes`; beyond
    that `get` returns None and the caller allocates ordinary (pageable) arrays."""

    def __init__(self, device=0, max_bytes=1 << 31):
        self.device, self.max_bytes, self.total = int(device), int(max_bytes), 0
        self._blocks = []
        holder = [object()]
        self._idle_refs = _refcount(holder[0])  # list slot + call arguments: what an unused block shows in `get`

    def get(self, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        blocks = self._blocks
        for i in range(len(blocks)):
            if blocks[i].nbytes == max(nbytes, 1) and _refcount(blocks[i]) <= self._idle_refs:
                return blocks[i].array(shape, dtype)
        if self.total + nbytes > self.max_bytes:
            return None
        blocks.append(_PinnedBlock(nbytes, self.device))
        self.total += nbytes
        return blocks[-1].array(shape, dtype)

    def clear(self):
        """Forget the blocks (those still viewed by caller arrays stay alive until the arrays go)."""
        self._blocks = []
        self.total = 0


class PinnedArray:
    """A numpy array backed by page-locked memory from mbt_host_alloc (freed when the last view of it is gone)."""

    def __init__(self, shape, dtype, device=0):
        self.shape = tuple(int(x) for x in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.array = _PinnedBlock(nbytes, device).array(self.shape, self.dtype)


def _addr(x):
    """Address of a host numpy array, a raw int device pointer, or anything with data_ptr() (torch tensor)."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "__cuda_array_interface__"):
        return x.__cuda_array_interface__["data"][0]
    raise TypeError(f"cannot take the address of {type(x)}")


class NativeEnv:
    """One mbt_env handle.  Host numpy buffers (mem=HOST) or device pointers / torch tensors (mem=DEVICE)."""

    def __init__(self, cfg, device=0):
        self._h = C.c_void_p()
        self.cfg = cfg
        self.device = device
        self.N = int(cfg.num_trajectories)
        self.A, self.D, self.S = config_dims(cfg)
        dout = C.c_int32()
        _check(load().mbt_config_obs_out_dim(C.byref(cfg), C.byref(dout)))
        self.Dout = dout.value  # emitted observation width (D unless cfg.obs_select picks columns)
        self.dtype = np.dtype(np.float64 if cfg.precision == _abi.MBT_F64 else np.float32)
        # element type of the action / observation / reward buffers of reset() and step()
        self.io_dtype = np.dtype(np.float32) if cfg.io_precision == _abi.MBT_IO_F32 else self.dtype
        _check(load().mbt_create(C.byref(cfg), device, C.byref(self._h)))

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            load().mbt_destroy(h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- control
    def seed(self, seed):
        _check(load().mbt_seed(self._h, C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF)))

    def sync(self):
        _check(load().mbt_sync(self._h))

    def set_stream(self, stream_ptr):
        """cudaStream_t as an int (0 = legacy default stream, torch's default); None = the handle's own stream."""
        _check(load().mbt_set_stream(self._h, C.c_void_p(-1 if stream_ptr is None else int(stream_ptr))))

    def fold_counters(self):
        """Last call inside a CUDA-graph captured episode: moves the step / episode counters to the device so every
        replay draws fresh random numbers (include/mbt_b200.h, mbt_fold_counters)."""
        _check(load().mbt_fold_counters(self._h))

    def prepare_capture(self):
        """Before a CUDA-graph capture on a handle that was already used: counters consumed so far move to the device."""
        _check(load().mbt_prepare_capture(self._h))

    def get_seed(self):
        v = C.c_uint64()
        _check(load().mbt_get_seed(self._h, C.byref(v)))
        return v.value

    def set_counters(self, n_step, n_episode):
        _check(load().mbt_set_counters(self._h, int(n_step), int(n_episode)))

    def reconfigure(self, cfg):
        """Swap the configuration in place (state, clock and counters kept); MbtError(MBT_E_STATE) if the new config
        changes the shape of the device state."""
        _check(load().mbt_reconfigure(self._h, C.byref(cfg)))
        self.cfg = cfg
        dout = C.c_int32()
        _check(load().mbt_config_obs_out_dim(C.byref(cfg), C.byref(dout)))
        self.Dout = dout.value

    # -- group of handles over NCCL (one process per GPU)
    @staticmethod
    def group_unique_id():
        buf = (C.c_char * _abi.MBT_GROUP_ID_BYTES)()
        _check(load().mbt_group_unique_id(buf))
        return bytes(buf)

    def group_create(self, unique_id, rank, world):
        assert len(unique_id) == _abi.MBT_GROUP_ID_BYTES
        buf = (C.c_char * _abi.MBT_GROUP_ID_BYTES).from_buffer_copy(unique_id)
        _check(load().mbt_group_create(self._h, buf, int(rank), int(world)))

    def group_destroy(self):
        _check(load().mbt_group_destroy(self._h))

    def group_info(self):
        r, w, t = C.c_int32(), C.c_int32(), C.c_int64()
        _check(load().mbt_group_info(self._h, C.byref(r), C.byref(w), C.byref(t)))
        return dict(rank=r.value, world=w.value, total_trajectories=t.value)

    def group_rollout(self, policy, returns_local=None, returns_all=None):
        """Fused rollout of the local shard + NCCL all-reduce of the summary (+ all-gather of returns, overlapped);
        returns the summary of ALL trajectories of the group.  Buffers are device tensors / pointers."""
        summary = _abi.mbt_summary()
        _check(load().mbt_group_rollout(self._h, C.byref(policy), C.byref(summary), _addr(returns_local), _addr(returns_all)))
        return summary

    def group_summary(self, local):
        out = _abi.mbt_summary()
        _check(load().mbt_group_summary(self._h, C.byref(local), C.byref(out)))
        return out

    def group_wait(self, host_sync=True):
        _check(load().mbt_group_wait(self._h, int(bool(host_sync))))

    # -- hot path
    def reset(self, obs_out=None, args=None, mem=_abi.MBT_MEM_HOST):
        _check(load().mbt_reset(self._h, None if args is None else C.byref(args), _addr(obs_out), mem))
        return obs_out

    def step(self, actions, obs_out=None, rew_out=None, mem=_abi.MBT_MEM_HOST):
        done = C.c_uint8(0)
        _check(load().mbt_step(self._h, _addr(actions), _addr(obs_out), _addr(rew_out), C.byref(done), mem))
        return bool(done.value)

    def rollout(self, policy, returns_out=None, terminal_q_out=None, mem=_abi.MBT_MEM_HOST):
        summary = _abi.mbt_summary()
        _check(load().mbt_rollout(self._h, C.byref(policy), C.byref(summary), _addr(returns_out), _addr(terminal_q_out), mem))
        return summary

    def rollout_record(self, policy, steps_capacity, obs=True, actions=True, rewards=True):
        """Fused rollout that records the trajectory into host arrays (time-major); returns (summary, obs, act, rew)."""
        o = np.empty((steps_capacity + 1, self.N, self.Dout), self.dtype) if obs else None
        a = np.empty((steps_capacity, self.N, self.A), self.dtype) if actions else None
        r = np.empty((steps_capacity, self.N), self.dtype) if rewards else None
        rec = _abi.mbt_record(_addr(o), _addr(a), _addr(r), int(steps_capacity))
        summary = _abi.mbt_summary()
        _check(load().mbt_rollout_record(self._h, C.byref(policy), C.byref(summary), C.byref(rec), _abi.MBT_MEM_HOST))
        n = int(summary.steps)
        return summary, (None if o is None else o[: n + 1]), (None if a is None else a[:n]), (None if r is None else r[:n])

    # -- state
    def get_state(self, out=None, mem=_abi.MBT_MEM_HOST):
        if out is None:
            out = np.empty((self.N, self.D), self.dtype)
        _check(load().mbt_get_state(self._h, _addr(out), mem))
        return out

    def set_state(self, state, mem=_abi.MBT_MEM_HOST):
        if mem == _abi.MBT_MEM_HOST:
            state = np.ascontiguousarray(state, self.dtype)
            assert state.shape == (self.N, self.D)
        _check(load().mbt_set_state(self._h, _addr(state), mem))

    def clock(self):
        t, k, n, e = C.c_double(), C.c_int64(), C.c_int64(), C.c_int64()
        _check(load().mbt_get_clock(self._h, C.byref(t), C.byref(k), C.byref(n), C.byref(e)))
        return dict(time=t.value, k=k.value, n_step=n.value, n_episode=e.value)

    def clip_count(self):
        c = C.c_int64()
        _check(load().mbt_get_clip_count(self._h, C.byref(c)))
        return c.value

    def reward_eval(self, cur, act, nxt, is_terminal=False):
        cur = np.ascontiguousarray(cur, self.dtype)
        act = np.ascontiguousarray(act, self.dtype)
        nxt = np.ascontiguousarray(nxt, self.dtype)
        out = np.empty((cur.shape[0],), self.dtype)
        _check(load().mbt_reward_eval(self._h, cur.shape[0], _addr(cur), _addr(act), _addr(nxt), int(bool(is_terminal)),
                                      _addr(out), _abi.MBT_MEM_HOST))
        return out

    def inventory_histogram(self, lo, hi, group_sum=False):
        """int64 counts of the current inventory column: [below lo, lo, lo+1, ..., hi, above hi or NaN]."""
        out = np.zeros(int(hi) - int(lo) + 3, np.int64)
        _check(load().mbt_inventory_histogram(self._h, int(lo), int(hi), out.ctypes.data_as(C.POINTER(C.c_int64)), int(bool(group_sum))))
        return out

    # -- checkpoint / resume
    def checkpoint(self):
        n = C.c_size_t()
        _check(load().mbt_checkpoint_size(self._h, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        _check(load().mbt_checkpoint_save(self._h, buf.ctypes.data, n.value))
        return buf

    def restore(self, buf):
        buf = np.ascontiguousarray(buf, np.uint8)
        _check(load().mbt_checkpoint_load(self._h, buf.ctypes.data, buf.size))

    # -- statistics
    def kernel_info(self):
        """Which step / rollout kernels this handle launches (run-time specialised or ahead-of-time), registers, cache."""
        info = _abi.mbt_kernel_info()
        _check(load().mbt_get_kernel_info(self._h, C.byref(info)))
        return dict(aot_variant=info.aot_variant, jit_mode=info.jit_mode, step_is_jit=bool(info.step_is_jit),
                    step_registers=info.step_registers, step_local_bytes=info.step_local_bytes,
                    jit_from_disk_cache=bool(info.jit_from_disk_cache), jit_compile_ms=info.jit_compile_ms,
                    jit_hash=f"{info.jit_hash:016x}", rollout_is_jit=list(info.rollout_is_jit),
                    message=info.message.decode("utf-8", "replace"))

    def launch_count(self):
        c = C.c_int64()
        _check(load().mbt_get_launch_count(self._h, C.byref(c)))
        return c.value

    def enable_timing(self, mode=1):
        """0 off; 1 two events around each kernel; 2 one event per kernel (interval to the next launch)."""
        _check(load().mbt_enable_timing(self._h, int(mode)))

    def kernel_times_ms(self):
        n = C.c_int64()
        _check(load().mbt_get_kernel_times(self._h, None, 0, C.byref(n)))
        buf = (C.c_float * max(n.value, 1))()
        _check(load().mbt_get_kernel_times(self._h, buf, n.value, C.byref(n)))
        return np.array(buf[: n.value], dtype=np.float32)
