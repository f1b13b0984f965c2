/*
 * mbt_oracle_impl.h -- TEST INFRASTRUCTURE.  Body of the CPU oracle, included twice by
 * mbt_oracle.c with REAL = double (SFX = f64) and REAL = float (SFX = f32).
 *
 * It restates, array-by-array and in the reference's own order of evaluation, what one call of
 * TradingEnvironment.step() / reset() computes (citations: /root/reference/mbt_gym/...).  The state
 * is kept exactly like the reference keeps it: one (N, D) row-major matrix with columns
 * CASH=0 INVENTORY=1 TIME=2 ASSET_PRICE=3 (gym/index_names.py:1-4) followed by the arrival-model and
 * price-impact-model columns (gym/TradingEnvironment.py:303-318).
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

typedef struct FN(orc_env) {
    mbt_config cfg;
    int64_t N;
    int A, D;
    REAL *state;    /* (N, D): model_dynamics.state                         ModelDynamics.py:39   */
    REAL *cur;      /* (N, D): current_state = state.copy()                  TradingEnvironment.py:105 */
    REAL *q0;       /* (N,):   reward_function.initial_inventory             RewardFunctions.py:72,111 */
    double t;       /* the uniform clock state[:, TIME] (every row equal)    TradingEnvironment.py:216 */
    double t0;      /* start time of the running episode                                           */
    uint64_t seed;
    int64_t n_step;    /* env-steps since seed()  -> Philox draw index, stream STEP  */
    int64_t n_episode; /* resets since seed()     -> Philox draw index, stream RESET */
    int64_t k;         /* steps in this episode */
    int started;
    int64_t clipped;
} FN(orc_env);

static REAL FN(orc_pow)(REAL x, REAL p) {
#if ORC_IS_F64
    return mbt_pow_f64(x, p);
#else
    return mbt_pow_f32(x, p);
#endif
}
static REAL FN(orc_exp)(REAL x) {
#if ORC_IS_F64
    return mbt_exp_f64(x);
#else
    return mbt_exp_f32(x);
#endif
}
static REAL FN(orc_normal)(uint32_t bits) {
    /* draw contract (include/mbt_philox.h): the float quantile in both precisions, widened for float64 */
    return (REAL)mbt_normal_from_bits_f32(bits);
}
static REAL FN(orc_nanmax)(REAL a, REAL b) { /* np.max / np.maximum: NaN wins */
    if (a != a) return a;
    if (b != b) return b;
    return a > b ? a : b;
}
static REAL FN(orc_clip)(REAL x, REAL lo, REAL hi) { /* np.clip = minimum(maximum(x, lo), hi) */
    REAL y = x < lo ? lo : x;
    return y > hi ? hi : y;
}

/* reward_function.calculate(current_state, action, next_state, is_terminal_step)
 * cs / s: one row of current / next state; a: de-normalised action row; dt_r = next[TIME]-current[TIME];
 * q0, L: what reward_function.reset captured (initial inventory, episode length). */
static REAL FN(orc_reward)(const mbt_config *c, const REAL *cs, const REAL *s, const REAL *a, int done, REAL dt_r,
                           REAL q0, REAL L) {
    REAL pnl = (s[0] + s[1] * s[3]) - (cs[0] + cs[1] * cs[3]); /* RewardFunctions.py:26-33 */
    REAL pe = (REAL)c->rew_exponent;
    switch (c->reward) {
    case MBT_REW_RUNNING_INVENTORY_PENALTY: /* RewardFunctions.py:128-138 */
        return (pnl - (dt_r * (REAL)c->rew_phi) * FN(orc_pow)(s[1], pe)) -
               ((REAL)c->rew_alpha * (REAL)done) * FN(orc_pow)(s[1], pe);
    case MBT_REW_CJ_MM: /* RewardFunctions.py:96-109 */
        return (pnl - (dt_r * (REAL)c->rew_phi) * FN(orc_pow)(s[1], pe)) -
               (REAL)c->rew_alpha *
                   ((FN(orc_pow)(s[1], pe) - FN(orc_pow)(cs[1], pe)) + (dt_r / L) * FN(orc_pow)(q0, pe));
    case MBT_REW_CJ_OE: /* RewardFunctions.py:55-70 */
        return (pnl - (dt_r * (REAL)c->rew_phi) * FN(orc_pow)(s[1], pe)) -
               (dt_r * (REAL)c->rew_alpha) *
                   ((pe * a[0]) * FN(orc_pow)(cs[1], pe - (REAL)1) + FN(orc_pow)(q0, pe) * L);
    case MBT_REW_EXP_UTILITY: /* RewardFunctions.py:156-163 */
        return done ? -FN(orc_exp)(-(REAL)c->rew_risk_aversion * (s[0] + s[1] * s[3])) : (REAL)0;
    default: /* MBT_REW_PNL */
        return pnl;
    }
}

static FN(orc_env) * FN(orc_create)(const mbt_config *cfg) {
    int32_t A, D, S;
    if (orc_dims(cfg, &A, &D, &S) != 0) return NULL;
    FN(orc_env) *e = (FN(orc_env) *)calloc(1, sizeof *e);
    e->cfg = *cfg;
    e->N = cfg->num_trajectories;
    e->A = A;
    e->D = D;
    e->state = (REAL *)calloc((size_t)e->N * D, sizeof(REAL));
    e->cur = (REAL *)calloc((size_t)e->N * D, sizeof(REAL));
    e->q0 = (REAL *)calloc((size_t)e->N, sizeof(REAL));
    e->seed = 0;
    return e;
}

static void FN(orc_destroy)(FN(orc_env) * e) {
    if (!e) return;
    free(e->state);
    free(e->cur);
    free(e->q0);
    free(e);
}

/* observation = normalise_observation(state.copy())     TradingEnvironment.py:101,110,112-118 */
static void FN(orc_write_obs)(const FN(orc_env) * e, REAL *obs) {
    const mbt_config *c = &e->cfg;
    const int Dout = orc_obs_out_dim(c, e->D);
    for (int64_t i = 0; i < e->N; ++i) {
        int j = 0;
        for (int d = 0; d < e->D; ++d) {
            REAL x = e->state[i * e->D + d];
            if (c->normalise_obs) x = (x - (REAL)c->obs_low[d]) / (REAL)c->obs_grad[d] - (REAL)1;
            /* ReduceStateSizeWrapper: obs[:, list_of_state_indices]      gym/wrappers.py:30-38 */
            if (!c->obs_select || ((c->obs_select >> d) & 1u)) obs[i * Dout + j++] = x;
        }
    }
}

/* TradingEnvironment.reset: processes reset to their initial vector state, state = initial_state,
 * reward_function.reset(state)                         TradingEnvironment.py:96-101,131-140 */
static void FN(orc_reset)(FN(orc_env) * e, const mbt_reset_args *args, REAL *obs_out) {
    const mbt_config *c = &e->cfg;
    double t0 = args ? args->start_time : c->start_time;
    int q0_mode = args ? args->q0_mode : c->q0_mode;
    double q0_const = args ? args->q0_const : c->q0_const;
    int64_t lo = args ? args->q0_lo : c->q0_lo, hi = args ? args->q0_hi : c->q0_hi;
    for (int64_t i = 0; i < e->N; ++i) {
        REAL *s = e->state + i * e->D;
        s[0] = (REAL)c->initial_cash;
        if (q0_mode == MBT_Q0_UNIFORM_INT) { /* rng.integers(lo, hi)  TradingEnvironment.py:271-272 */
            mbt_u32x4 r = mbt_draw(e->seed, (uint64_t)(c->traj_offset + i), (uint64_t)e->n_episode, MBT_STREAM_RESET);
            uint64_t span = (uint64_t)(hi - lo);
            s[1] = (REAL)(lo + (int64_t)(((uint64_t)r.x * span) >> 32));
        } else if (q0_mode == MBT_Q0_PER_TRAJ) {
            s[1] = (REAL)args->q0_values[i]; /* a callable that returned an array: assigned as is  TradingEnvironment.py:275-279,137 */
        } else {
            s[1] = (REAL)q0_const; /* TradingEnvironment.py:273-279 */
        }
        s[2] = (REAL)t0;
        s[3] = (REAL)c->mid_initial; /* midprice initial_state */
        int col = 4;
        if (c->midprice == MBT_MID_HESTON) s[col++] = (REAL)c->heston_var0; /* initial_state = [[price, variance]]  midprice_models.py:346 */
        if (c->arrival == MBT_ARR_HAWKES) { /* initial_state = baseline_arrival_rate  arrival_models.py:103 */
            s[col++] = (REAL)c->arr_rate[0];
            s[col++] = (REAL)c->arr_rate[1];
        }
        if (c->fill == MBT_FILL_EXOGENOUS_MM && (c->dynamics == MBT_DYN_LIMIT || c->dynamics == MBT_DYN_LIMIT_AND_MARKET)) {
            /* initial_state = the two depth processes' initial states; the reference never updates these columns
             * (fill_probability_models.py:144-152,168-170) */
            s[col++] = (REAL)c->fill_depth0[0];
            s[col++] = (REAL)c->fill_depth0[1];
        }
        if (c->impact == MBT_IMP_TEMP_PERM) s[col++] = (REAL)0; /* price_impact_models.py:83 */
        if (c->impact == MBT_IMP_TEMP_TRANSIENT || c->impact == MBT_IMP_TRANSIENT)
            s[col++] = (REAL)c->imp_initial; /* initial_transient_impact   price_impact_models.py:124,167 */
        e->q0[i] = s[1]; /* reward_function.reset: initial_inventory  RewardFunctions.py:72,111 */
    }
    e->t = t0;
    e->t0 = t0;
    e->k = 0;
    e->n_episode += 1;
    e->started = 1;
    if (obs_out) FN(orc_write_obs)(e, obs_out);
}

/*
 * One env-step given the step's random numbers:
 *   u (N,4): [arrival bid, arrival ask, fill bid, fill ask] uniforms;  z (N,): midprice normal.
 * This is the function the reference is compared with draw-for-draw (oracle/ref_shim.py).
 */
static void FN(orc_step_core)(FN(orc_env) * e, const REAL *actions_in, const REAL *u, const REAL *z, const REAL *z2,
                               REAL *obs_out, REAL *rew_out, uint8_t *done_out) {
    const mbt_config *c = &e->cfg;
    const int D = e->D, A = e->A;
    const REAL qmax = (REAL)c->max_inventory, cmax = (REAL)c->max_cash;
    const double t_cur = e->t;
    const double t_next = t_cur + c->step_size; /* state[:, TIME] += step_size   TradingEnvironment.py:216 */
    const int done = t_next >= c->terminal_time - c->step_size / 2; /* :218-220 */
    memcpy(e->cur, e->state, sizeof(REAL) * (size_t)e->N * D); /* current_state = state.copy()  :105 */

    /* host-side scalars exactly as the reference's Python floats form them */
    const REAL drift_dt = (REAL)(c->mid_drift * c->mid_step);      /* midprice_models.py:63 */
    const REAL vol_sqdt = (REAL)(c->mid_vol * sqrt(c->mid_step));  /* midprice_models.py:64,143 */
    const REAL sqdt = (REAL)sqrt(c->mid_step);
    /* state columns after [cash, inventory, time, price]: Heston's variance first, then the arrival / impact model's */
    const int mc = 4 + (c->midprice == MBT_MID_HESTON ? 1 : 0);
    const REAL rho = (REAL)c->heston_corr, rho_c = (REAL)sqrt(1.0 - c->heston_corr * c->heston_corr);
    REAL p_arr[2] = {0, 0};
    if (c->arrival == MBT_ARR_POISSON) { /* intensity * step_size  arrival_models.py:56 */
        p_arr[0] = (REAL)(c->arr_rate[0] * c->arr_step);
        p_arr[1] = (REAL)(c->arr_rate[1] * c->arr_step);
    } else if (c->arrival == MBT_ARR_POISSON_NONLINEAR) { /* 1 - exp(-intensity*step)  arrival_models.py:83 */
        p_arr[0] = (REAL)(1.0 - mbt_exp_f64(-c->arr_rate[0] * c->arr_step));
        p_arr[1] = (REAL)(1.0 - mbt_exp_f64(-c->arr_rate[1] * c->arr_step));
    }

    /* Triangular / Power fill functions: `np.max(depths, 0)` is a reduction over the trajectory axis, so the step's
     * fill probabilities are batch-wide scalars             fill_probability_models.py:82,113 */
    REAL p_fill_batch[2] = {0, 0};
    if ((c->fill == MBT_FILL_TRIANGULAR || c->fill == MBT_FILL_POWER) &&
        (c->dynamics == MBT_DYN_LIMIT || c->dynamics == MBT_DYN_LIMIT_AND_MARKET)) {
        REAL m[2] = {0, 0};
        for (int64_t i = 0; i < e->N; ++i)
            for (int j = 0; j < 2; ++j) { /* depths = action[:, 0:2], de-normalised   ModelDynamics.py:50-51,128-130 */
                REAL x = actions_in[i * A + j];
                if (c->normalise_action) x = (x + (REAL)1) * (REAL)c->act_grad[j] + (REAL)c->act_low[j];
                m[j] = (i == 0) ? x : FN(orc_nanmax)(m[j], x); /* np.max(depths, 0): NaN propagates */
            }
        if (c->fill == MBT_FILL_TRIANGULAR) { /* np.max(1 - np.max(depths, 0) / max_fill_depth, 0): a scalar  :82 */
            REAL pb = (REAL)1 - m[0] / (REAL)c->fill_max_depth, pa = (REAL)1 - m[1] / (REAL)c->fill_max_depth;
            p_fill_batch[0] = p_fill_batch[1] = FN(orc_nanmax)(pb, pa);
        } else /* (1 + (fill_multiplier * np.max(depths, 0)) ** fill_exponent) ** -1: one value per side  :113 */
            for (int j = 0; j < 2; ++j)
                p_fill_batch[j] = (REAL)1 / ((REAL)1 + FN(orc_pow)((REAL)c->fill_multiplier * m[j], (REAL)c->fill_exponent));
    }

    for (int64_t i = 0; i < e->N; ++i) {
        REAL *s = e->state + i * D;
        const REAL *cs = e->cur + i * D;
        REAL a[MBT_MAX_ACTION_DIM];
        for (int j = 0; j < A; ++j) { /* normalise_action(inverse=True)   TradingEnvironment.py:104,120-126 */
            REAL x = actions_in[i * A + j];
            if (c->normalise_action) x = (x + (REAL)1) * (REAL)c->act_grad[j] + (REAL)c->act_low[j];
            a[j] = x;
        }
        const REAL S = cs[3]; /* ModelDynamics.midprice: pre-update price   ModelDynamics.py:82-84 */
        REAL arr[2] = {0, 0};
        REAL fil_bid = 0, fil_ask = 0; /* fills after max-inventory suppression (what the processes' update() receives) */

        if (c->dynamics == MBT_DYN_LIMIT || c->dynamics == MBT_DYN_AT_TOUCH || c->dynamics == MBT_DYN_LIMIT_AND_MARKET) {
            /* ---- get_arrivals_and_fills                                ModelDynamics.py:127-131,169-172 */
            REAL fil[2];
            for (int j = 0; j < 2; ++j) {
                REAL p = (c->arrival == MBT_ARR_HAWKES) ? cs[mc + j] * (REAL)c->arr_step /* arrival_models.py:123 */
                                                        : p_arr[j];
                arr[j] = (u[i * 4 + j] < p) ? (REAL)1 : (REAL)0; /* unif < p   arrival_models.py:55-56 */
            }
            if (c->dynamics == MBT_DYN_AT_TOUCH) {
                fil[0] = a[0]; /* fills = action[:, 0:2]   ModelDynamics.py:157-158,171 */
                fil[1] = a[1];
            } else {
                for (int j = 0; j < 2; ++j) { /* unif < exp(-kappa * depth)   fill_probability_models.py:33,58 */
                    REAL p;
                    if (c->fill == MBT_FILL_EXPONENTIAL) {
                        p = FN(orc_exp)(-(REAL)c->fill_exponent * a[j]);
                    } else if (c->fill == MBT_FILL_EXOGENOUS_MM) {
                        /* (depths > d) * base * exp(-kappa * (depths - d)) + (depths <= d), d = the model's (constant)
                         * state column of the side                      fill_probability_models.py:160-163 */
                        REAL d = cs[mc + (c->arrival == MBT_ARR_HAWKES ? 2 : 0) + j];
                        REAL x = -(REAL)c->fill_exponent * (a[j] - d);
                        REAL ex = (x >= (ORC_IS_F64 ? (REAL)709.782712893384 : (REAL)88.72284)) ? (REAL)INFINITY : FN(orc_exp)(x);
                        p = ((REAL)(a[j] > d) * (REAL)c->fill_base) * ex + (REAL)(a[j] <= d);
                    } else {
                        p = p_fill_batch[j];
                    }
                    fil[j] = (u[i * 4 + 2 + j] < p) ? (REAL)1 : (REAL)0;
                }
            }
            /* ---- _remove_max_inventory_fills (pre-step inventory)      TradingEnvironment.py:146-152,323-327 */
            fil[0] = (REAL)(1 - (cs[1] >= qmax)) * fil[0];
            fil[1] = (REAL)(1 - (cs[1] <= -qmax)) * fil[1];
            fil_bid = fil[0];
            fil_ask = fil[1];
            /* ---- update_state                                          ModelDynamics.py:108-116,151-160,208-224 */
            REAL hs = (REAL)c->half_spread;
            if (c->dynamics == MBT_DYN_LIMIT_AND_MARKET) {
                REAL mo_buy = (a[2] > (REAL)0.5) ? (REAL)1 : (REAL)0;
                REAL mo_sell = (a[3] > (REAL)0.5) ? (REAL)1 : (REAL)0;
                REAL best_bid = S - hs, best_ask = S + hs;
                s[0] = s[0] + (mo_sell * best_bid - mo_buy * best_ask);
                s[1] = s[1] + (mo_buy - mo_sell);
            }
            REAL off_b = (c->dynamics == MBT_DYN_AT_TOUCH) ? hs : a[0];
            REAL off_a = (c->dynamics == MBT_DYN_AT_TOUCH) ? hs : a[1];
            /* fill_multiplier = (-1, +1)                                  ModelDynamics.py:71-73 */
            REAL dq = (arr[0] * fil[0]) * (REAL)1 + (arr[1] * fil[1]) * (REAL)-1;
            REAL dc = (((REAL)-1 * arr[0]) * fil[0]) * (S + off_b * (REAL)-1) +
                      (((REAL)1 * arr[1]) * fil[1]) * (S + off_a * (REAL)1);
            s[1] = s[1] + dq;
            s[0] = s[0] + dc;
        } else { /* MBT_DYN_SPEED                                          ModelDynamics.py:262-267 */
            REAL nu = a[0];
            REAL impact;
            if (c->impact == MBT_IMP_TEMP_PERM)
                impact = (REAL)c->imp_temp * nu + cs[mc]; /* price_impact_models.py:91-92 */
            else if (c->impact == MBT_IMP_TEMP_TRANSIENT) /* price_impact_models.py:133-134 */
                impact = (REAL)c->imp_temp * nu + (REAL)c->imp_transient * cs[mc];
            else if (c->impact == MBT_IMP_TRANSIENT) /* price_impact_models.py:174-175 */
                impact = (REAL)c->imp_transient * cs[mc];
            else
                impact = (REAL)c->imp_temp * FN(orc_pow)(nu, (REAL)c->imp_exponent); /* :55-56 */
            REAL px = S + impact;
            REAL vol = nu * (REAL)c->mid_step;
            s[0] = s[0] - vol * px;
            s[1] = s[1] + vol;
        }
        /* ---- _clip_inventory_and_cash                                  TradingEnvironment.py:283-297 */
        REAL qc = FN(orc_clip)(s[1], -qmax, qmax), cc = FN(orc_clip)(s[0], -cmax, cmax);
        if (qc != s[1] || cc != s[0]) e->clipped += 1;
        s[1] = qc;
        s[0] = cc;
        s[2] = (REAL)t_next;
        /* ---- _update_market_state: midprice, arrival, fill, impact     TradingEnvironment.py:206-211,303-309 */
        switch (c->midprice) {
        case MBT_MID_BM: /* midprice_models.py:60-65 */
            s[3] = (S + drift_dt) + vol_sqdt * z[i];
            break;
        case MBT_MID_GBM: /* midprice_models.py:97-105 */
            s[3] = (S + ((REAL)c->mid_drift * S) * (REAL)c->mid_step) + ((((REAL)c->mid_vol * S) * sqdt) * z[i]);
            break;
        case MBT_MID_OU: /* midprice_models.py:140-143 (drift NOT scaled by dt, as written there) */
            s[3] = S + ((-(REAL)c->ou_speed) * (S - (REAL)c->ou_level) + vol_sqdt * z[i]);
            break;
        case MBT_MID_BM_JUMP: { /* midprice_models.py:222-230: jumps with the agent's own (post-suppression) fills */
            REAL fb = fil_bid * arr[0], fa = fil_ask * arr[1];
            s[3] = ((S + drift_dt) + vol_sqdt * z[i]) + ((REAL)c->mid_jump * fa - (REAL)c->mid_jump * fb);
            break;
        }
        case MBT_MID_OU_JUMP: { /* midprice_models.py:262-270 */
            REAL fb = fil_bid * arr[0], fa = fil_ask * arr[1];
            s[3] = ((S - (REAL)c->ou_speed * (S - (REAL)c->ou_level)) + vol_sqdt * z[i]) +
                   ((REAL)c->mid_jump * fa - (REAL)c->mid_jump * fb);
            break;
        }
        case MBT_MID_HESTON: { /* midprice_models.py:354-369; weiners = (z, rho z + sqrt(1-rho^2) z2): the draw contract
                                * for np.random.multivariate_normal(0, [[1, rho], [rho, 1]]) */
            REAL v = cs[4], w_s = z[i], w_v = rho * z[i] + rho_c * z2[i];
#if ORC_IS_F64
            REAL vol = sqrt(v * (REAL)c->mid_step);
#else
            REAL vol = sqrtf(v * (REAL)c->mid_step);
#endif
            s[3] = (S + ((REAL)c->mid_drift * S) * (REAL)c->mid_step) + ((vol * S) * w_s);
            REAL nv = (v + ((REAL)c->heston_speed * ((REAL)c->heston_level - v)) * (REAL)c->mid_step) +
                      (((REAL)c->heston_volvol * vol) * w_v);
            s[4] = nv < (REAL)0 ? -nv : (nv == (REAL)0 ? (REAL)0 : nv); /* np.abs */
            break;
        }
        default: /* MBT_MID_CONSTANT  midprice_models.py:32-33 */
            break;
        }
        if (c->arrival == MBT_ARR_HAWKES) { /* arrival_models.py:110-119 */
            for (int j = 0; j < 2; ++j) {
                REAL lam = cs[mc + j];
                s[mc + j] = (lam + (((REAL)c->hawkes_speed * ((REAL)c->arr_rate[j] - lam)) * (REAL)c->arr_step)) +
                           (REAL)c->hawkes_jump * arr[j];
            }
        }
        if (c->impact == MBT_IMP_TEMP_PERM) /* price_impact_models.py:88-89 */
            s[mc] = cs[mc] + ((REAL)c->imp_perm * a[0]) * (REAL)c->imp_step;
        if (c->impact == MBT_IMP_TEMP_TRANSIENT || c->impact == MBT_IMP_TRANSIENT) /* price_impact_models.py:129-131,170-172 */
            s[mc] = (cs[mc] - ((REAL)c->imp_resilience * cs[mc]) * (REAL)c->imp_step) +
                   ((REAL)c->imp_kernel * a[0]) * (REAL)c->imp_step;

        /* ---- reward_function.calculate(current_state, action, next_state, dones[0])  :108 */
        REAL r = FN(orc_reward)(c, cs, s, a, done, (REAL)(t_next - t_cur), e->q0[i],
                                (REAL)(c->rew_terminal_time - e->t0));
        if (c->normalise_rewards) r = (REAL)c->reward_scaling * r; /* TradingEnvironment.py:128-129 */
        if (rew_out) rew_out[i] = r;
    }
    e->t = t_next;
    e->k += 1;
    e->n_step += 1;
    if (obs_out) FN(orc_write_obs)(e, obs_out);
    if (done_out) *done_out = (uint8_t)done;
}

/* The step's random numbers from the Philox draw contract (include/mbt_philox.h). */
static void FN(orc_fill_draws)(uint64_t seed, int64_t traj_offset, int64_t N, int64_t n_step, REAL *u, REAL *z, REAL *z2) {
    for (int64_t i = 0; i < N; ++i) {
        if (z2) /* second normal of the step (Heston variance): the normal bits of the block of stream STEP2 */
            z2[i] = FN(orc_normal)(mbt_normal_bits(mbt_draw(seed, (uint64_t)(traj_offset + i), (uint64_t)n_step, MBT_STREAM_STEP2)));
        if (!u) continue;
        mbt_u32x4 r = mbt_draw(seed, (uint64_t)(traj_offset + i), (uint64_t)n_step, MBT_STREAM_STEP);
        u[i * 4 + 0] = (REAL)mbt_uniform_bits24(r.x) * (REAL)5.9604644775390625e-08; /* 2^-24 */
        u[i * 4 + 1] = (REAL)mbt_uniform_bits24(r.y) * (REAL)5.9604644775390625e-08;
        u[i * 4 + 2] = (REAL)mbt_uniform_bits24(r.z) * (REAL)5.9604644775390625e-08;
        u[i * 4 + 3] = (REAL)mbt_uniform_bits24(r.w) * (REAL)5.9604644775390625e-08;
        z[i] = FN(orc_normal)(mbt_normal_bits(r));
    }
}

static void FN(orc_step)(FN(orc_env) * e, const REAL *actions, REAL *obs_out, REAL *rew_out, uint8_t *done_out) {
    REAL *u = (REAL *)malloc(sizeof(REAL) * 4 * (size_t)e->N);
    REAL *z = (REAL *)malloc(sizeof(REAL) * (size_t)e->N);
    REAL *z2 = (e->cfg.midprice == MBT_MID_HESTON) ? (REAL *)malloc(sizeof(REAL) * (size_t)e->N) : NULL;
    FN(orc_fill_draws)(e->seed, e->cfg.traj_offset, e->N, e->n_step, u, z, z2);
    FN(orc_step_core)(e, actions, u, z, z2, obs_out, rew_out, done_out);
    free(u);
    free(z);
    free(z2);
}

#undef FN
#undef CAT
#undef CAT_
