/*
 * mbt_host_params.h -- host-side (no CUDA) helpers: config validation, dims, and the per-launch
 * StepParams<T> built from mbt_config.  Derived constants are formed in float64 exactly the way the
 * reference's Python floats form them, then cast once to the arithmetic type.
 */
#ifndef MBT_HOST_PARAMS_H
#define MBT_HOST_PARAMS_H

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "mbt_step_core.cuh"

/* action dim A (ModelDynamics.get_action_space), obs dim D = 3 + sum(process dims)
 * (TradingEnvironment.py:232-241,311-318), persistent state columns S = D - 1 (time is a uniform scalar). */
static inline int mbt_dims(const mbt_config *c, int32_t *A, int32_t *D, int32_t *S) {
    int a, d = 4;
    switch (c->dynamics) {
    case MBT_DYN_LIMIT: a = 2; break;
    case MBT_DYN_AT_TOUCH: a = 2; break;
    case MBT_DYN_LIMIT_AND_MARKET: a = 4; break;
    case MBT_DYN_SPEED: a = 1; break;
    default: return MBT_E_UNSUPPORTED;
    }
    if (c->midprice == MBT_MID_HESTON) d += 1; /* (price, variance)  midprice_models.py:346 */
    if (c->arrival == MBT_ARR_HAWKES) d += 2;
    if (c->fill == MBT_FILL_EXOGENOUS_MM && (c->dynamics == MBT_DYN_LIMIT || c->dynamics == MBT_DYN_LIMIT_AND_MARKET))
        d += 2; /* the two exogenous best-depth columns  fill_probability_models.py:144-152 */
    if (c->impact == MBT_IMP_TEMP_PERM || c->impact == MBT_IMP_TEMP_TRANSIENT || c->impact == MBT_IMP_TRANSIENT) d += 1;
    if (A) *A = a;
    if (D) *D = d;
    if (S) *S = d - 1;
    return MBT_OK;
}

/* width of emitted observation rows: D, or the number of selected columns (ReduceStateSizeWrapper fused) */
static inline int mbt_obs_out_dim(const mbt_config *c, int D) {
    if (!c->obs_select) return D;
    int n = 0;
    for (int d = 0; d < D; ++d) n += (c->obs_select >> d) & 1u;
    return n;
}

static inline int mbt_validate_config(const mbt_config *c, std::string &err) {
    char buf[256];
    if (!c) { err = "config is NULL"; return MBT_E_INVALID_ARG; }
    if (c->struct_size != (int32_t)sizeof(mbt_config)) {
        snprintf(buf, sizeof buf, "mbt_config.struct_size=%d, library expects %zu (ABI mismatch)", c->struct_size,
                 sizeof(mbt_config));
        err = buf;
        return MBT_E_INVALID_ARG;
    }
    if (c->precision != MBT_F64 && c->precision != MBT_F32) { err = "precision must be MBT_F64 or MBT_F32"; return MBT_E_INVALID_ARG; }
    if (c->io_precision != MBT_IO_SAME && c->io_precision != MBT_IO_F32) { err = "io_precision must be MBT_IO_SAME or MBT_IO_F32"; return MBT_E_INVALID_ARG; }
    if (c->num_trajectories <= 0) { err = "num_trajectories must be > 0"; return MBT_E_INVALID_ARG; }
    if (c->traj_offset < 0) { err = "traj_offset must be >= 0"; return MBT_E_INVALID_ARG; }
    if (c->n_steps <= 0 || !(c->step_size > 0) || !(c->terminal_time > 0)) {
        err = "n_steps, step_size and terminal_time must be > 0";
        return MBT_E_INVALID_ARG;
    }
    if (mbt_dims(c, nullptr, nullptr, nullptr) != MBT_OK) { err = "unknown dynamics kind"; return MBT_E_UNSUPPORTED; }
    if (c->midprice < MBT_MID_CONSTANT || c->midprice > MBT_MID_HESTON) { err = "unknown midprice model"; return MBT_E_UNSUPPORTED; }
    if (c->midprice == MBT_MID_HESTON && c->normalise_obs) {
        /* the reference publishes bounds for ONE of the model's two state columns (midprice_models.py:343-346), so its
         * normalise_observation fails to broadcast (N, D) against (D-1,) bounds */
        err = "the Heston midprice model has no bounds for its variance column: observation normalisation is undefined "
              "(the reference fails too)";
        return MBT_E_UNSUPPORTED;
    }
    if (c->midprice == MBT_MID_HESTON && !(c->heston_corr >= -1.0 && c->heston_corr <= 1.0)) {
        err = "heston_corr (weiner_correlation) must be within [-1, 1]";
        return MBT_E_INVALID_ARG;
    }
    if (c->reward < MBT_REW_PNL || c->reward > MBT_REW_EXP_UTILITY) { err = "unknown reward function"; return MBT_E_UNSUPPORTED; }
    if (c->dynamics == MBT_DYN_SPEED) {
        /* ModelDynamics.py:273-275: required_processes = ["price_impact_model"] */
        if (c->impact < MBT_IMP_TEMP_PERM || c->impact > MBT_IMP_TRANSIENT) {
            err = "speed dynamics needs a price impact model";
            return MBT_E_UNSUPPORTED;
        }
        if (c->midprice == MBT_MID_BM_JUMP || c->midprice == MBT_MID_OU_JUMP) {
            err = "jump midprice models need limit-order fills (the reference fails with speed dynamics too)";
            return MBT_E_UNSUPPORTED;
        }
        if (c->arrival == MBT_ARR_HAWKES) { err = "speed dynamics with a Hawkes arrival model is not supported"; return MBT_E_UNSUPPORTED; }
    } else {
        /* ModelDynamics.py:123-125,163-165,234-236: arrival (+ fill) models required */
        if (c->arrival < MBT_ARR_POISSON || c->arrival > MBT_ARR_HAWKES) { err = "limit-order dynamics need an arrival model"; return MBT_E_UNSUPPORTED; }
        if (c->dynamics != MBT_DYN_AT_TOUCH && (c->fill < MBT_FILL_EXPONENTIAL || c->fill > MBT_FILL_EXOGENOUS_MM)) {
            err = "limit-order dynamics need a fill probability model (exponential, triangular, power or exogenous-MM)";
            return MBT_E_UNSUPPORTED;
        }
        {
            int32_t D = 0;
            mbt_dims(c, nullptr, &D, nullptr);
            if (D > MBT_MAX_OBS_DIM) {
                err = "Heston midprice + Hawkes arrivals + exogenous-MM fills need 9 observation columns; at most 8 are supported";
                return MBT_E_UNSUPPORTED;
            }
        }
        if (c->impact != MBT_IMP_NONE) { err = "price impact models only combine with speed dynamics"; return MBT_E_UNSUPPORTED; }
        if (c->reward == MBT_REW_CJ_OE) { err = "CjOeCriterion needs a 1-d action (reference fails too, RewardFunctions.py:66)"; return MBT_E_UNSUPPORTED; }
    }
    {
        int32_t D = 0;
        mbt_dims(c, nullptr, &D, nullptr);
        if (c->obs_select && (mbt_obs_out_dim(c, D) == 0 || (c->obs_select >> D) != 0)) {
            err = "obs_select must pick at least one of the D observation columns and no others";
            return MBT_E_INVALID_ARG;
        }
    }
    if (c->q0_mode == MBT_Q0_UNIFORM_INT && !(c->q0_hi > c->q0_lo)) { err = "initial inventory range needs hi > lo"; return MBT_E_INVALID_ARG; }
    if (c->q0_mode != MBT_Q0_UNIFORM_INT && c->q0_mode != MBT_Q0_CONST) { err = "unknown q0_mode"; return MBT_E_INVALID_ARG; }
    return MBT_OK;
}

/* the time column of an observation: the uniform clock, normalised here with the very expression the kernels apply to the
 * other columns (norm_obs), in the arithmetic type T */
template <typename T>
static inline T mbt_time_obs(const mbt_config &c, double t) {
    if (!c.normalise_obs) return (T)t;
    const T low = (T)c.obs_low[2], grad = (T)c.obs_grad[2];
    return mbt_div_rcp_t((T)t - low, grad, mbt_rcp_for_div_t(grad)) - (T)1;
}

/* The uniform clock of the step that moves time from t_cur to t_next. */
template <typename T>
static inline StepClock<T> mbt_make_clock(const mbt_config &c, double t_cur, double t_next, double t0) {
    StepClock<T> ck;
    ck.t_next = (T)t_next;
    ck.t_obs = mbt_time_obs<T>(c, t_next);
    ck.dt_r = (T)(t_next - t_cur);
    ck.done = t_next >= c.terminal_time - c.step_size / 2; /* TradingEnvironment.py:218-220 */
    clock_derive<T>(ck, (T)c.rew_phi, (T)c.rew_alpha, (T)(c.rew_terminal_time - t0));
    return ck;
}

/* Uniform per-episode parameters (t0 = start time of the running episode). */
template <typename T>
static inline StepParams<T> mbt_make_params(const mbt_config &c, double t0, int q0_per_traj, double q0_uniform) {
    StepParams<T> p;
    memset(&p, 0, sizeof p);
    int32_t A, D, S;
    mbt_dims(&c, &A, &D, &S);
    p.dyn = c.dynamics; p.mid = c.midprice; p.arr = c.arrival; p.imp = c.impact; p.rew = c.reward;
    /* only dynamics that draw fills from the fill model look at its kind (AtTheTouch / speed dynamics ignore it) */
    p.fill = (c.dynamics == MBT_DYN_LIMIT || c.dynamics == MBT_DYN_LIMIT_AND_MARKET) ? c.fill : MBT_FILL_NONE;
    p.action_dim = A; p.obs_dim = D;
    p.obs_select = (int)c.obs_select; p.obs_out_dim = mbt_obs_out_dim(&c, D);
    p.normalise_action = c.normalise_action; p.normalise_obs = c.normalise_obs; p.normalise_rewards = c.normalise_rewards;
    p.q0_per_traj = q0_per_traj;
    p.ep_len = (T)(c.rew_terminal_time - t0);
    p.q0_uniform = (T)q0_uniform;
    p.qmax = (T)c.max_inventory; p.cmax = (T)c.max_cash;
    if (c.arrival == MBT_ARR_POISSON) {
        p.p_arr[0] = (T)(c.arr_rate[0] * c.arr_step);
        p.p_arr[1] = (T)(c.arr_rate[1] * c.arr_step);
    } else if (c.arrival == MBT_ARR_POISSON_NONLINEAR) {
        p.p_arr[0] = (T)(1.0 - mbt_exp_f64(-c.arr_rate[0] * c.arr_step));
        p.p_arr[1] = (T)(1.0 - mbt_exp_f64(-c.arr_rate[1] * c.arr_step));
    }
    for (int j = 0; j < 2; ++j) { /* k*2^-24 < p  <=>  k < ceil(p*2^24), k integer in [0, 2^24) */
        double scaled = std::ceil((double)p.p_arr[j] * 16777216.0);
        p.arr_thr[j] = scaled <= 0.0 ? 0u : (scaled >= 16777216.0 ? 16777216u : (uint32_t)scaled);
        if (!((double)p.p_arr[j] == (double)p.p_arr[j])) p.arr_thr[j] = 0u; /* NaN: never */
    }
    p.arr_step_2p24 = (T)c.arr_step * (T)16777216.0;
    p.arr_step = (T)c.arr_step; p.arr_rate[0] = (T)c.arr_rate[0]; p.arr_rate[1] = (T)c.arr_rate[1];
    p.hawkes_speed = (T)c.hawkes_speed; p.hawkes_jump = (T)c.hawkes_jump;
    p.neg_kappa = -(T)c.fill_exponent;
    p.fill_max_depth = (T)c.fill_max_depth; p.fill_mult = (T)c.fill_multiplier; p.fill_pexp = (T)c.fill_exponent;
    p.fill_base = (T)c.fill_base; p.fill_depth0[0] = (T)c.fill_depth0[0]; p.fill_depth0[1] = (T)c.fill_depth0[1];
    /* exp() overflows to +inf above ln(max finite value) of the arithmetic type */
    p.exp_overflow = sizeof(T) == 8 ? (T)709.782712893384 : (T)88.72284;
    p.drift_dt = (T)(c.mid_drift * c.mid_step);
    p.vol_sqdt = (T)(c.mid_vol * std::sqrt(c.mid_step));
    p.sqdt = (T)std::sqrt(c.mid_step);
    p.mid_drift = (T)c.mid_drift; p.mid_vol = (T)c.mid_vol; p.mid_step = (T)c.mid_step;
    p.heston_speed = (T)c.heston_speed; p.heston_level = (T)c.heston_level; p.heston_rho = (T)c.heston_corr;
    p.heston_rho_c = (T)std::sqrt(1.0 - c.heston_corr * c.heston_corr); p.heston_xi = (T)c.heston_volvol;
    p.ou_neg_speed = -(T)c.ou_speed; p.ou_speed = (T)c.ou_speed; p.ou_level = (T)c.ou_level; p.mid_jump = (T)c.mid_jump;
    p.imp_temp = (T)c.imp_temp; p.imp_perm = (T)c.imp_perm; p.imp_exp = (T)c.imp_exponent; p.imp_step = (T)c.imp_step;
    p.imp_transient = (T)c.imp_transient; p.imp_resilience = (T)c.imp_resilience; p.imp_kernel = (T)c.imp_kernel;
    p.half_spread = (T)c.half_spread;
    p.phi = (T)c.rew_phi; p.alpha = (T)c.rew_alpha; p.pexp = (T)c.rew_exponent; p.risk_aversion = (T)c.rew_risk_aversion;
    p.reward_scaling = (T)c.reward_scaling;
    for (int i = 0; i < MBT_MAX_ACTION_DIM; ++i) { p.act_low[i] = (T)c.act_low[i]; p.act_grad[i] = (T)c.act_grad[i]; }
    for (int i = 0; i < MBT_MAX_OBS_DIM; ++i) {
        p.obs_low[i] = (T)c.obs_low[i];
        p.obs_grad[i] = (T)c.obs_grad[i];
        p.obs_rcp[i] = c.normalise_obs ? mbt_rcp_for_div_t(p.obs_grad[i]) : (T)0; /* reciprocal of the value the kernel divides by */
    }
    return p;
}

#endif /* MBT_HOST_PARAMS_H */
