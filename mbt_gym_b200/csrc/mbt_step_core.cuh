/*
 * mbt_step_core.cuh -- one env-step of ONE trajectory, state in registers.
 *
 * This is the arithmetic of the reference's TradingEnvironment.step() call tree
 * (mbt_gym/gym/TradingEnvironment.py:103-110,198-220,283-297,323-327 and the model classes it calls)
 * collapsed into a single function over scalars: no (N,2) temporaries, no state copies.  It is used by
 * the step kernel (one call per launch) and by the fused rollout kernel (one call per time step, state
 * never leaving registers).
 *
 * Exactness rules (DESIGN.md "Numerics"): the expression trees below are the reference's, operation
 * for operation, so in float64 every intermediate rounds like numpy's.  The build uses -fmad=false; the
 * only fused operations are the explicit fma() calls inside include/mbt_math.h.
 *
 * The `V` template argument fixes model kinds at compile time (V::dyn etc. >= 0) or leaves them to the
 * runtime config (-1).  The kernels are issue-bound (profiles/), so the BASELINE configurations get fully
 * specialised instantiations -- action/observation widths, reward kind and "no normalisation" known to
 * the compiler -- and everything else runs the generic one.
 */
#ifndef MBT_STEP_CORE_CUH
#define MBT_STEP_CORE_CUH

#include "mbt_b200.h" /* include/ is on the include path (nvcc -I; NVRTC: in-memory headers by these names) */
#include "mbt_math.h"
#include "mbt_philox.h"

/*
 * Compile-time description of a configuration: every member is the MBT_* enum / flag value, or -1 = "read it from the
 * runtime StepParams".  VariantFull fixes any subset; the ahead-of-time table (mbt_variants.h) uses the short form
 * Variant<DYN, MID, ARR, IMP, REW, NORM>, and the run-time specialiser (mbt_jit.h, NVRTC) instantiates VariantFull with
 * EVERYTHING fixed for the handle's configuration -- no model switch is left in the kernel.
 *   na / no / nr   normalise action / observation / rewards compiled in (1), out (0), runtime (-1)
 *   sel            observation column mask of the fused ReduceStateSizeWrapper (0 = all columns), or -1 = runtime
 */
MBT_HD constexpr int mbt_popcount8(int m) {
    return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + ((m >> 3) & 1) + ((m >> 4) & 1) + ((m >> 5) & 1) + ((m >> 6) & 1) + ((m >> 7) & 1);
}

template <int DYN, int MID, int ARR, int IMP, int REW, int FILL, int NA, int NO, int NR, int SEL>
struct VariantFull {
    static constexpr int dyn = DYN, mid = MID, arr = ARR, imp = IMP, rew = REW, fill = FILL, na = NA, no = NO, nr = NR, sel = SEL;
    /* action width when the dynamics are fixed (0 = runtime) */
    static constexpr int A = DYN < 0 ? 0 : (DYN == MBT_DYN_SPEED ? 1 : (DYN == MBT_DYN_LIMIT_AND_MARKET ? 4 : 2));
    /* state-matrix width D when every model that owns columns is fixed (0 = runtime): [cash, inventory, time, price],
     * Heston's variance, Hawkes' two intensities, the exogenous-MM fill model's two depths, one impact column */
    static constexpr bool limit_fills = DYN == MBT_DYN_LIMIT || DYN == MBT_DYN_LIMIT_AND_MARKET;
    static constexpr int D = (DYN < 0 || MID < 0 || ARR < 0 || IMP < 0 || FILL < 0) ? 0
                             : 4 + (MID == MBT_MID_HESTON ? 1 : 0) + (ARR == MBT_ARR_HAWKES ? 2 : 0) +
                                   ((limit_fills && FILL == MBT_FILL_EXOGENOUS_MM) ? 2 : 0) +
                                   ((IMP == MBT_IMP_TEMP_PERM || IMP == MBT_IMP_TEMP_TRANSIENT || IMP == MBT_IMP_TRANSIENT) ? 1 : 0);
    /* emitted observation width (0 = runtime) */
    static constexpr int Dout = (D == 0 || SEL < 0) ? 0 : (SEL == 0 ? D : mbt_popcount8(SEL));
};

/* short form of the ahead-of-time table.  NORM: 0 = no action/observation/reward normalisation and no column selection
 * compiled in; 1 = action AND observation normalisation compiled in (the reference's default constructor, what SB3
 * sees), no reward scaling, all columns; -1 = runtime flags.  Fixed dynamics imply the exponential fill function (the
 * table only routes that one to the specialised variants, mbt_variants.h). */
template <int DYN, int MID, int ARR, int IMP, int REW, int NORM>
using Variant = VariantFull<DYN, MID, ARR, IMP, REW, (DYN >= 0 ? MBT_FILL_EXPONENTIAL : -1), (NORM < 0 ? -1 : NORM), (NORM < 0 ? -1 : NORM),
                            (NORM < 0 ? -1 : 0), (NORM < 0 ? -1 : 0)>;
using VariantGeneric = Variant<-1, -1, -1, -1, -1, -1>;

/* math dispatch on the arithmetic type */
MBT_HD float mbt_exp_t(float x) { return mbt_exp_f32(x); }
MBT_HD double mbt_exp_t(double x) { return mbt_exp_f64(x); }
MBT_HD float mbt_pow_t(float x, float p) { return mbt_pow_f32(x, p); }
MBT_HD double mbt_pow_t(double x, double p) { return mbt_pow_f64(x, p); }
MBT_HD float mbt_exp2k_t(float x, int k) { return mbt_exp2k_f32(x, k); }
MBT_HD double mbt_exp2k_t(double x, int k) { return mbt_exp2k_f64(x, k); }
MBT_HD float mbt_sqrt_t(float x) { return sqrtf(x); }
MBT_HD double mbt_sqrt_t(double x) { return sqrt(x); }
MBT_HD int mbt_u24_below_exp_t(uint32_t k, float x) { return mbt_u24_below_exp_f32(k, x); }
MBT_HD int mbt_u24_below_exp_t(uint32_t k, double x) { return mbt_u24_below_exp_f64(k, x); }
MBT_HD float mbt_div_rcp_t(float a, float b, float y) { return mbt_div_rcp_f32(a, b, y); }
MBT_HD double mbt_div_rcp_t(double a, double b, double y) { return mbt_div_rcp_f64(a, b, y); }
MBT_HD float mbt_rcp_for_div_t(float b) { return mbt_rcp_for_div_f32(b); }
MBT_HD double mbt_rcp_for_div_t(double b) { return mbt_rcp_for_div_f64(b); }
MBT_HD void mbt_real_t(uint32_t k, float *u) { *u = mbt_u24_to_real_f32(k); }
MBT_HD void mbt_real_t(uint32_t k, double *u) { *u = mbt_u24_to_real_f64(k); }

/*
 * Uniform (per-episode) quantities, already in the arithmetic type T.  Built on the host by
 * mbt_make_params() from mbt_config, forming each derived constant the way the reference's Python floats
 * do (e.g. volatility * sqrt(step_size) in float64, then cast).
 */
template <typename T>
struct StepParams {
    /* model selectors (runtime copies; compile-time Variant wins when >= 0) */
    int dyn, mid, arr, imp, rew, fill;
    int action_dim, obs_dim;
    int obs_select, obs_out_dim; /* ReduceStateSizeWrapper fused into the store: column bitmask (0 = all), emitted width */
    int normalise_action, normalise_obs, normalise_rewards;
    int q0_per_traj; /* 1: read q0 column (random initial inventories), 0: q0_uniform */

    T ep_len;    /* reward_function.episode_length = T - t0              RewardFunctions.py:73,112 */
    T q0_uniform;
    T qmax, cmax;

    T p_arr[2];  /* Poisson: intensity*step_size ; NonLinear: 1-exp(-intensity*step_size) */
    uint32_t arr_thr[2]; /* ceil(p_arr * 2^24) clamped to [0, 2^24]:  k*2^-24 < p_arr  <=>  k < arr_thr */
    T arr_step, arr_step_2p24 /* arr_step * 2^24 */, arr_rate[2], hawkes_speed, hawkes_jump;
    T neg_kappa; /* -fill_exponent */
    T fill_max_depth, fill_mult, fill_pexp; /* Triangular: max_fill_depth;  Power: fill_multiplier, fill_exponent */
    T fill_base, fill_depth0[2];            /* ExogenousMm: base_fill_probability, exogenous best depth (bid, ask) */
    T exp_overflow;                         /* smallest argument whose exp() is +inf in numpy's arithmetic type */
    T drift_dt, vol_sqdt, sqdt, mid_drift, mid_vol, mid_step, ou_neg_speed, ou_speed, ou_level, mid_jump;
    T heston_speed, heston_level, heston_rho, heston_rho_c /* sqrt(1 - rho^2) */, heston_xi;
    T imp_temp, imp_perm, imp_exp, imp_step, imp_transient, imp_resilience, imp_kernel, half_spread;
    T phi, alpha, pexp, risk_aversion, reward_scaling;
    T act_low[MBT_MAX_ACTION_DIM], act_grad[MBT_MAX_ACTION_DIM];
    T obs_low[MBT_MAX_OBS_DIM], obs_grad[MBT_MAX_OBS_DIM];
    T obs_rcp[MBT_MAX_OBS_DIM]; /* RN(1 / obs_grad[d]) in T, or 0: see mbt_div_rcp_f64 (include/mbt_math.h) */
};

/* The uniform clock of ONE step (every trajectory shares it, TradingEnvironment.py:216-220). */
template <typename T>
struct StepClock {
    T t_next; /* time column after the step                           TradingEnvironment.py:216 */
    T t_obs;  /* t_next as the observation shows it: normalised on the host when normalise_obs (:112-118) -- the clock is
               * uniform, so one host evaluation of the same IEEE expression replaces one division per trajectory */
    T dt_r;   /* next[TIME] - current[TIME] as the rewards read it    RewardFunctions.py:58,99,131 */
    int done; /* this step is the terminal one                        TradingEnvironment.py:218-220 */
    /* products / quotient of UNIFORM quantities that the reward formulas contain: formed once per step (on the host for
     * the step kernel) with the same IEEE operation each thread would otherwise repeat -- bit-identical results */
    T dt_phi;      /* dt_r * phi                RewardFunctions.py:60,101,133 */
    T dt_alpha;    /* dt_r * alpha              RewardFunctions.py:61-62      */
    T alpha_done;  /* alpha * int(is_terminal)  RewardFunctions.py:134-136    */
    T dt_over_len; /* dt_r / episode_length     RewardFunctions.py:107        */
};

template <typename T>
MBT_HD void clock_derive(StepClock<T> &ck, T phi, T alpha, T ep_len) {
    ck.dt_phi = ck.dt_r * phi;
    ck.dt_alpha = ck.dt_r * alpha;
    ck.alpha_done = alpha * (T)ck.done;
    ck.dt_over_len = ck.dt_r / ep_len;
}

/* Per-trajectory state carried between steps (the SoA columns of DESIGN.md "Layout"). */
template <typename T>
struct Traj {
    T cash, inv, mid;
    T x0, x1; /* Hawkes: (lambda_bid, lambda_ask);  Temp+Perm impact: x0 = accumulated permanent impact */
    T var;    /* Heston: the variance column of the midprice model */
};

template <int CT>
MBT_HD int pick(int runtime) { return CT >= 0 ? CT : runtime; }

/* price-impact models that carry one state column (permanent impact I, or transient impact Y) */
MBT_HD bool imp_has_state(int imp) { return imp == MBT_IMP_TEMP_PERM || imp == MBT_IMP_TEMP_TRANSIENT || imp == MBT_IMP_TRANSIENT; }

/* fill functions whose probability is a reduction over the whole batch (see include/mbt_b200.h, MBT_FILL_TRIANGULAR) */
MBT_HD bool fill_is_batch(int fill) { return fill == MBT_FILL_TRIANGULAR || fill == MBT_FILL_POWER; }

/* np.max / np.maximum of two values: NaN wins */
template <typename T>
MBT_HD T nanmax(T a, T b) {
    if (a != a) return a;
    if (b != b) return b;
    return a > b ? a : b;
}

/*
 * The step's fill probabilities from the deepest quote of the batch on each side (m_bid, m_ask = np.max(depths, 0)),
 * scaled by 2^24 for the integer-valued uniforms (exact: power-of-two scaling; NaN stays NaN = never filled).
 *   Triangular  np.max(1 - np.max(depths, 0) / max_fill_depth, 0)                 fill_probability_models.py:82
 *   Power       (1 + (fill_multiplier * np.max(depths, 0)) ** fill_exponent) ** -1   (`** -1` is 1/x in numpy)  :113
 */
/* one side's threshold (the step kernel computes the two sides on two threads): `m` = that side's batch maximum, `m_other`
 * the other side's (the triangular function reduces over both sides) */
template <typename T>
MBT_HD T fill_batch_threshold_side(const StepParams<T> &p, T m, T m_other, int side) {
    if (p.fill == MBT_FILL_TRIANGULAR) {
        const T pm = (T)1 - m / p.fill_max_depth, po = (T)1 - m_other / p.fill_max_depth;
        return (side == 0 ? nanmax<T>(pm, po) : nanmax<T>(po, pm)) * (T)16777216.0; /* argument order of fill_batch_thresholds */
    }
    return ((T)1 / ((T)1 + mbt_pow_t(p.fill_mult * m, p.fill_pexp))) * (T)16777216.0;
}

template <typename T>
MBT_HD void fill_batch_thresholds(const StepParams<T> &p, T m_bid, T m_ask, T *thr) {
    if (p.fill == MBT_FILL_TRIANGULAR) {
        const T pb = (T)1 - m_bid / p.fill_max_depth, pa = (T)1 - m_ask / p.fill_max_depth;
        thr[0] = thr[1] = nanmax<T>(pb, pa) * (T)16777216.0;
    } else {
        thr[0] = ((T)1 / ((T)1 + mbt_pow_t(p.fill_mult * m_bid, p.fill_pexp))) * (T)16777216.0;
        thr[1] = ((T)1 / ((T)1 + mbt_pow_t(p.fill_mult * m_ask, p.fill_pexp))) * (T)16777216.0;
    }
}

/*
 * ExogenousMmFillProbabilityModel._get_fill_probabilities, one side, times 2^24            fill_probability_models.py:160-163
 *   (depths > d) * base * np.exp(-kappa * (depths - d)) + (depths <= d)
 * evaluated term by term like numpy does (so an overflowing exp gives 0 * inf = NaN = "never filled", and a NaN depth
 * gives 0 * NaN + 0 = NaN as well).
 */
template <typename T>
MBT_HD T fill_exogenous_threshold(const StepParams<T> &p, T depth, int side) {
    const T d = p.fill_depth0[side];
    const T gt = depth > d ? (T)1 : (T)0, le = depth <= d ? (T)1 : (T)0;
    const T x = p.neg_kappa * (depth - d);
    /* exp(x) * 2^24; numpy's exp overflows to +inf where the clamped kernel exp stays finite */
    const T e = (x >= p.exp_overflow) ? (T)INFINITY : mbt_exp2k_t(x, 24);
    return (gt * p.fill_base) * e + le * (T)16777216.0;
}

/* reward_function.calculate for one row; (c0, q_cur, S0) = current_state, s = next_state. */
template <typename T, class V>
MBT_HD T reward_one(const StepParams<T> &p, const StepClock<T> &ck, T c0, T q_cur, T S0, const Traj<T> &s, const T *a, T q_init) {
    const int rew = pick<V::rew>(p.rew);
    T pnl = (s.cash + s.inv * s.mid) - (c0 + q_cur * S0); /* PnL  RewardFunctions.py:26-33 */
    if (rew == MBT_REW_PNL) return pnl;
    if (rew == MBT_REW_EXP_UTILITY) /* RewardFunctions.py:156-163 */
        return ck.done ? -mbt_exp_t(-p.risk_aversion * (s.cash + s.inv * s.mid)) : (T)0;
    T qp = mbt_pow_t(s.inv, p.pexp);
    T base = pnl - ck.dt_phi * qp;
    if (rew == MBT_REW_RUNNING_INVENTORY_PENALTY) /* RewardFunctions.py:128-138 */
        return base - ck.alpha_done * qp;
    if (rew == MBT_REW_CJ_MM) /* RewardFunctions.py:96-109 */
        return base - p.alpha * ((qp - mbt_pow_t(q_cur, p.pexp)) + ck.dt_over_len * mbt_pow_t(q_init, p.pexp));
    /* MBT_REW_CJ_OE  RewardFunctions.py:55-70 */
    return base - ck.dt_alpha *
                      ((p.pexp * a[0]) * mbt_pow_t(q_cur, p.pexp - (T)1) + mbt_pow_t(q_init, p.pexp) * p.ep_len);
}

/*
 * Advance one trajectory by one step.
 *   s       in/out  state
 *   a       raw (de-normalised) action, p.action_dim values
 *   r       the 128 random bits of this (trajectory, step)   include/mbt_philox.h draw contract
 *   q_init  initial inventory of the episode (for CjMm / CjOe)
 *   fill_thr  batch-reduced fill models only: the step's two fill probabilities * 2^24 (fill_batch_thresholds)
 *   nbits2  Heston only: the 32 normal bits of the step's SECOND Philox block (stream MBT_STREAM_STEP2)
 * returns the (scaled) reward; *clipped is set when inventory or cash hit their bounds.
 */
/* FILTER: decide exponential fills through the float filter (mbt_u24_below_exp_f64).  The fused rollout with a FIXED action
 * turns it off: there the fill probability is loop-invariant and the plain comparison lets the compiler hoist the
 * exponential out of the time loop altogether (same decisions either way). */
template <typename T, class V, bool FILTER = true>
MBT_HD T step_one(const StepParams<T> &p, const StepClock<T> &ck, Traj<T> &s, const T *a, mbt_u32x4 r, T q_init, int *clipped,
                  const T *fill_thr = nullptr, uint32_t nbits2 = 0u) {
    const int dyn = pick<V::dyn>(p.dyn), mid = pick<V::mid>(p.mid), arr_kind = pick<V::arr>(p.arr),
              imp = pick<V::imp>(p.imp);
    const int fill = pick<V::fill>(p.fill);
    const T c0 = s.cash, q_cur = s.inv, S = s.mid; /* current_state = state.copy()   TradingEnvironment.py:105 */
    T arr_b = 0, arr_a = 0;
    T own_b = 0, own_a = 0; /* the agent's own executed fills (fill * arrival), for the jump midprice models */

    if (dyn != MBT_DYN_SPEED) {
        /* get_arrivals_and_fills                                    ModelDynamics.py:127-131,169-172 */
        /* every "unif < p" below is evaluated as  k < p * 2^24  on the 24-bit integer k of the draw: the same
         * predicate as  k * 2^-24 < p  (power-of-two scaling is exact), minus a multiply per uniform */
        const uint32_t kb = mbt_uniform_bits24(r.x), ka = mbt_uniform_bits24(r.y);
        if (arr_kind == MBT_ARR_HAWKES) { /* unif < lambda_t * step   arrival_models.py:121-123 */
            T ub, ua;
            mbt_real_t(kb, &ub);
            mbt_real_t(ka, &ua);
            arr_b = (ub < s.x0 * p.arr_step_2p24) ? (T)1 : (T)0;
            arr_a = (ua < s.x1 * p.arr_step_2p24) ? (T)1 : (T)0;
        } else { /* unif < p, p uniform over the batch   arrival_models.py:54-56,81-83 */
            arr_b = (kb < p.arr_thr[0]) ? (T)1 : (T)0;
            arr_a = (ka < p.arr_thr[1]) ? (T)1 : (T)0;
        }
        T fil_b, fil_a, off_b, off_a;
        if (dyn == MBT_DYN_AT_TOUCH) { /* fills = action[:, 0:2]      ModelDynamics.py:157-158,171 */
            fil_b = a[0];
            fil_a = a[1];
            off_b = p.half_spread;
            off_a = p.half_spread;
        } else { /* unif < exp(-kappa*depth)   fill_probability_models.py:28-34,57-58 */
            const uint32_t kvb = mbt_uniform_bits24(r.z), kva = mbt_uniform_bits24(r.w);
            if (fill_is_batch(fill)) { /* unif < p, p one value per side for the whole batch   :82,113 */
                T vb, va;
                mbt_real_t(kvb, &vb);
                mbt_real_t(kva, &va);
                fil_b = (vb < fill_thr[0]) ? (T)1 : (T)0;
                fil_a = (va < fill_thr[1]) ? (T)1 : (T)0;
            } else if (fill == MBT_FILL_EXOGENOUS_MM) { /* :160-163 */
                T vb, va;
                mbt_real_t(kvb, &vb);
                mbt_real_t(kva, &va);
                fil_b = (vb < fill_exogenous_threshold<T>(p, a[0], 0)) ? (T)1 : (T)0;
                fil_a = (va < fill_exogenous_threshold<T>(p, a[1], 1)) ? (T)1 : (T)0;
            } else if (FILTER) { /* k < exp(-kappa * depth) * 2^24 (mbt_u24_below_exp_*: float filter, double decision) */
                fil_b = mbt_u24_below_exp_t(kvb, p.neg_kappa * a[0]) ? (T)1 : (T)0;
                fil_a = mbt_u24_below_exp_t(kva, p.neg_kappa * a[1]) ? (T)1 : (T)0;
            } else {
                T vb, va;
                mbt_real_t(kvb, &vb);
                mbt_real_t(kva, &va);
                fil_b = (vb < mbt_exp2k_t(p.neg_kappa * a[0], 24)) ? (T)1 : (T)0;
                fil_a = (va < mbt_exp2k_t(p.neg_kappa * a[1], 24)) ? (T)1 : (T)0;
            }
            off_b = a[0];
            off_a = a[1];
        }
        /* _remove_max_inventory_fills, on the pre-step inventory     TradingEnvironment.py:146-152,323-327 */
        fil_b = (q_cur >= p.qmax) ? (T)0 * fil_b : fil_b;
        fil_a = (q_cur <= -p.qmax) ? (T)0 * fil_a : fil_a;
        if (dyn == MBT_DYN_LIMIT_AND_MARKET) { /* market orders first   ModelDynamics.py:208-215 */
            T mo_buy = (a[2] > (T)0.5) ? (T)1 : (T)0, mo_sell = (a[3] > (T)0.5) ? (T)1 : (T)0;
            s.cash = s.cash + (mo_sell * (S - p.half_spread) - mo_buy * (S + p.half_spread));
            s.inv = s.inv + (mo_buy - mo_sell);
        }
        /* update_state with fill_multiplier (-1,+1)                   ModelDynamics.py:71-73,108-116 */
        T b = arr_b * fil_b, k = arr_a * fil_a;
        own_b = fil_b * arr_b;
        own_a = fil_a * arr_a;
        s.inv = s.inv + (b - k);
        s.cash = s.cash + (k * (S + off_a) - b * (S - off_b));
    } else { /* TradinghWithSpeedModelDynamics.update_state            ModelDynamics.py:262-267 */
        T nu = a[0];
        T impact;
        if (imp == MBT_IMP_TEMP_PERM) impact = p.imp_temp * nu + s.x0;                               /* price_impact_models.py:91-92 */
        else if (imp == MBT_IMP_TEMP_TRANSIENT) impact = p.imp_temp * nu + p.imp_transient * s.x0;   /* :133-134 */
        else if (imp == MBT_IMP_TRANSIENT) impact = p.imp_transient * s.x0;                          /* :174-175 */
        else impact = p.imp_temp * mbt_pow_t(nu, p.imp_exp);                                         /* :55-56 */
        T vol = nu * p.mid_step;
        s.cash = s.cash - vol * (S + impact);
        s.inv = s.inv + vol;
    }
    /* _clip_inventory_and_cash                                        TradingEnvironment.py:283-297 */
    T qc = s.inv < -p.qmax ? -p.qmax : s.inv;
    qc = qc > p.qmax ? p.qmax : qc;
    T cc = s.cash < -p.cmax ? -p.cmax : s.cash;
    cc = cc > p.cmax ? p.cmax : cc;
    if (qc != s.inv || cc != s.cash) *clipped = 1;
    s.inv = qc;
    s.cash = cc;

    /* _update_market_state: midprice, arrival, fill, impact            TradingEnvironment.py:206-211 */
    if (mid != MBT_MID_CONSTANT) {
        const T z = (T)mbt_normal_from_bits_f32(mbt_normal_bits(r)); /* draw contract: float quantile, widened */
        if (mid == MBT_MID_BM) /* midprice_models.py:60-65 */
            s.mid = (S + p.drift_dt) + p.vol_sqdt * z;
        else if (mid == MBT_MID_GBM) /* midprice_models.py:97-105 */
            s.mid = (S + (p.mid_drift * S) * p.mid_step) + (((p.mid_vol * S) * p.sqdt) * z);
        else if (mid == MBT_MID_OU) /* midprice_models.py:140-143: drift not scaled by dt, as written there */
            s.mid = S + (p.ou_neg_speed * (S - p.ou_level) + p.vol_sqdt * z);
        else if (mid == MBT_MID_HESTON) {
            /* midprice_models.py:354-369.  The reference draws (W_S, W_v) ~ N(0, [[1, rho], [rho, 1]]) from the GLOBAL
             * np.random; the draw contract here: W_S = z, W_v = rho * z + sqrt(1 - rho^2) * z2 (Cholesky), z2 the normal of
             * the step's second Philox block.  Both updates use the variance BEFORE the step. */
            const T z2 = (T)mbt_normal_from_bits_f32(nbits2);
            const T w_v = p.heston_rho * z + p.heston_rho_c * z2;
            const T v = s.var;
            const T vol = mbt_sqrt_t(v * p.mid_step); /* np.sqrt(variance * step_size): correctly rounded on both sides */
            s.mid = (S + (p.mid_drift * S) * p.mid_step) + ((vol * S) * z);
            const T nv = (v + (p.heston_speed * (p.heston_level - v)) * p.mid_step) + ((p.heston_xi * vol) * w_v);
            s.var = nv < (T)0 ? -nv : (nv == (T)0 ? (T)0 : nv); /* np.abs (also turns -0.0 into +0.0) */
        } else if (mid == MBT_MID_BM_JUMP) /* midprice_models.py:222-230: jumps on the agent's own fills */
            s.mid = ((S + p.drift_dt) + p.vol_sqdt * z) + (p.mid_jump * own_a - p.mid_jump * own_b);
        else /* MBT_MID_OU_JUMP  midprice_models.py:262-270 */
            s.mid = ((S - p.ou_speed * (S - p.ou_level)) + p.vol_sqdt * z) + (p.mid_jump * own_a - p.mid_jump * own_b);
    }
    if (arr_kind == MBT_ARR_HAWKES) { /* arrival_models.py:110-119 (jump on arrival, not on fill) */
        s.x0 = (s.x0 + ((p.hawkes_speed * (p.arr_rate[0] - s.x0)) * p.arr_step)) + p.hawkes_jump * arr_b;
        s.x1 = (s.x1 + ((p.hawkes_speed * (p.arr_rate[1] - s.x1)) * p.arr_step)) + p.hawkes_jump * arr_a;
    }
    if (imp == MBT_IMP_TEMP_PERM) /* price_impact_models.py:88-89 */
        s.x0 = s.x0 + (p.imp_perm * a[0]) * p.imp_step;
    else if (imp == MBT_IMP_TEMP_TRANSIENT || imp == MBT_IMP_TRANSIENT) /* price_impact_models.py:129-131,170-172 */
        s.x0 = (s.x0 - (p.imp_resilience * s.x0) * p.imp_step) + (p.imp_kernel * a[0]) * p.imp_step;

    /* rewards = reward_function.calculate(current_state, action, next_state, dones[0])   :108 */
    T rwd = reward_one<T, V>(p, ck, c0, q_cur, S, s, a, q_init);
    return pick<V::nr>(p.normalise_rewards) ? p.reward_scaling * rwd : rwd; /* :128-129 */
}

/* normalise_action(inverse=True)                                       TradingEnvironment.py:120-126 */
template <typename T, class V>
MBT_HD T denorm_action(const StepParams<T> &p, T x, int j) {
    return pick<V::na>(p.normalise_action) ? (x + (T)1) * p.act_grad[j] + p.act_low[j] : x;
}
/* normalise_observation                                                 TradingEnvironment.py:112-118 */
template <typename T, class V>
MBT_HD T norm_obs(const StepParams<T> &p, T x, int d) {
    /* (x - low) / grad - 1 with the division done through the host's reciprocal: bit-identical to `/` (mbt_div_rcp_f64) */
    return pick<V::no>(p.normalise_obs) ? mbt_div_rcp_t(x - p.obs_low[d], p.obs_grad[d], p.obs_rcp[d]) - (T)1 : x;
}

#endif /* MBT_STEP_CORE_CUH */
