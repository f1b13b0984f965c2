"""ctypes mirror of include/mbt_b200.h (PODs and constants only -- no library is loaded here).

Kept in one place so the product binding (`_lib.py`) and the test oracle's loader
(`oracle/oracle.py`) describe the SAME `mbt_config` bytes.
"""
import ctypes as C

MBT_ABI_VERSION = 3

MBT_OK = 0
MBT_E_INVALID_ARG = -1
MBT_E_CUDA = -2
MBT_E_STATE = -3
MBT_E_UNSUPPORTED = -4
MBT_E_NOMEM = -5
MBT_E_NCCL = -6

MBT_MEM_HOST = 0
MBT_MEM_DEVICE = 1

MBT_F64 = 0
MBT_F32 = 1

MBT_IO_SAME = 0
MBT_IO_F32 = 1

MBT_DYN_LIMIT = 0
MBT_DYN_SPEED = 1
MBT_DYN_AT_TOUCH = 2
MBT_DYN_LIMIT_AND_MARKET = 3

MBT_MID_CONSTANT = 0
MBT_MID_BM = 1
MBT_MID_GBM = 2
MBT_MID_OU = 3
MBT_MID_BM_JUMP = 4
MBT_MID_OU_JUMP = 5
MBT_MID_HESTON = 6

MBT_ARR_NONE = 0
MBT_ARR_POISSON = 1
MBT_ARR_POISSON_NONLINEAR = 2
MBT_ARR_HAWKES = 3

MBT_FILL_NONE = 0
MBT_FILL_EXPONENTIAL = 1
MBT_FILL_TRIANGULAR = 2
MBT_FILL_POWER = 3
MBT_FILL_EXOGENOUS_MM = 4
FILLS_WITH_BATCH_REDUCTION = (MBT_FILL_TRIANGULAR, MBT_FILL_POWER)

MBT_IMP_NONE = 0
MBT_IMP_TEMP_PERM = 1
MBT_IMP_TEMP_POWER = 2
MBT_IMP_TEMP_TRANSIENT = 3
MBT_IMP_TRANSIENT = 4
IMPACTS_WITH_STATE = (MBT_IMP_TEMP_PERM, MBT_IMP_TEMP_TRANSIENT, MBT_IMP_TRANSIENT)

MBT_REW_PNL = 0
MBT_REW_RUNNING_INVENTORY_PENALTY = 1
MBT_REW_CJ_MM = 2
MBT_REW_CJ_OE = 3
MBT_REW_EXP_UTILITY = 4

MBT_Q0_CONST = 0
MBT_Q0_UNIFORM_INT = 1
MBT_Q0_PER_TRAJ = 2

MBT_POL_FIXED = 0
MBT_POL_AVELLANEDA_STOIKOV = 1
MBT_POL_CJ_MM_TABLE = 2
MBT_POL_SCHEDULE = 3

MBT_MAX_ACTION_DIM = 4
MBT_MAX_OBS_DIM = 8


class mbt_config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("precision", C.c_int32),
        ("num_trajectories", C.c_int64),
        ("traj_offset", C.c_int64),
        ("n_steps", C.c_int32),
        ("dynamics", C.c_int32),
        ("midprice", C.c_int32),
        ("arrival", C.c_int32),
        ("fill", C.c_int32),
        ("impact", C.c_int32),
        ("reward", C.c_int32),
        ("terminal_time", C.c_double),
        ("step_size", C.c_double),
        ("start_time", C.c_double),
        ("initial_cash", C.c_double),
        ("q0_mode", C.c_int32),
        ("_pad0", C.c_int32),
        ("q0_const", C.c_double),
        ("q0_lo", C.c_int64),
        ("q0_hi", C.c_int64),
        ("max_inventory", C.c_double),
        ("max_cash", C.c_double),
        ("mid_initial", C.c_double),
        ("mid_drift", C.c_double),
        ("mid_vol", C.c_double),
        ("mid_step", C.c_double),
        ("ou_level", C.c_double),
        ("ou_speed", C.c_double),
        ("mid_jump", C.c_double),
        ("heston_speed", C.c_double),
        ("heston_level", C.c_double),
        ("heston_corr", C.c_double),
        ("heston_volvol", C.c_double),
        ("heston_var0", C.c_double),
        ("arr_rate", C.c_double * 2),
        ("arr_step", C.c_double),
        ("hawkes_jump", C.c_double),
        ("hawkes_speed", C.c_double),
        ("fill_exponent", C.c_double),
        ("fill_max_depth", C.c_double),
        ("fill_multiplier", C.c_double),
        ("fill_base", C.c_double),
        ("fill_depth0", C.c_double * 2),
        ("imp_temp", C.c_double),
        ("imp_perm", C.c_double),
        ("imp_exponent", C.c_double),
        ("imp_step", C.c_double),
        ("imp_transient", C.c_double),
        ("imp_resilience", C.c_double),
        ("imp_kernel", C.c_double),
        ("imp_initial", C.c_double),
        ("half_spread", C.c_double),
        ("rew_phi", C.c_double),
        ("rew_alpha", C.c_double),
        ("rew_exponent", C.c_double),
        ("rew_terminal_time", C.c_double),
        ("rew_risk_aversion", C.c_double),
        ("normalise_action", C.c_int32),
        ("normalise_obs", C.c_int32),
        ("normalise_rewards", C.c_int32),
        ("_pad1", C.c_int32),
        ("act_low", C.c_double * MBT_MAX_ACTION_DIM),
        ("act_grad", C.c_double * MBT_MAX_ACTION_DIM),
        ("obs_low", C.c_double * MBT_MAX_OBS_DIM),
        ("obs_grad", C.c_double * MBT_MAX_OBS_DIM),
        ("reward_scaling", C.c_double),
        ("obs_select", C.c_uint32),
        ("io_precision", C.c_uint32),
    ]


class mbt_reset_args(C.Structure):
    _fields_ = [
        ("start_time", C.c_double),
        ("q0_mode", C.c_int32),
        ("_pad", C.c_int32),
        ("q0_const", C.c_double),
        ("q0_lo", C.c_int64),
        ("q0_hi", C.c_int64),
        ("q0_values", C.c_void_p),
    ]


MBT_GROUP_ID_BYTES = 128


class mbt_policy(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("table_rows", C.c_int32),
        ("table_cols", C.c_int32),
        ("inv_offset", C.c_int32),
        ("fixed", C.c_double * MBT_MAX_ACTION_DIM),
        ("as_gamma", C.c_double),
        ("as_sigma_sq", C.c_double),
        ("as_fill_comp", C.c_double),
        ("as_terminal_time", C.c_double),
        ("table", C.c_void_p),
    ]


class mbt_record(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("actions", C.c_void_p), ("rewards", C.c_void_p), ("steps_capacity", C.c_int64)]


class mbt_summary(C.Structure):
    _fields_ = [
        ("count", C.c_int64),
        ("steps", C.c_int64),
        ("sum_return", C.c_double),
        ("sum_return_sq", C.c_double),
        ("sum_q", C.c_double),
        ("sum_q_sq", C.c_double),
        ("sum_action", C.c_double),
        ("sum_reward_sq", C.c_double),
        ("clipped", C.c_int64),
    ]


class mbt_kernel_info(C.Structure):
    _fields_ = [
        ("aot_variant", C.c_int32),
        ("jit_mode", C.c_int32),
        ("step_is_jit", C.c_int32),
        ("step_registers", C.c_int32),
        ("step_local_bytes", C.c_int32),
        ("jit_from_disk_cache", C.c_int32),
        ("jit_compile_ms", C.c_double),
        ("jit_hash", C.c_uint64),
        ("rollout_is_jit", C.c_int32 * 5),
        ("_pad", C.c_int32),
        ("message", C.c_char * 256),
    ]


def new_config(**kw):
    """Zero-initialised mbt_config with struct_size set and the given fields assigned."""
    cfg = mbt_config()
    cfg.struct_size = C.sizeof(mbt_config)
    cfg.reward_scaling = 1.0
    cfg.rew_exponent = 2.0
    cfg.imp_exponent = 1.0
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError(f"mbt_config has no field {k!r}")
        if isinstance(v, (list, tuple)) or hasattr(v, "__len__"):
            arr = getattr(cfg, k)
            for i, x in enumerate(v):
                arr[i] = x
        else:
            setattr(cfg, k, v)
    return cfg
