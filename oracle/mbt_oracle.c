/*
 * mbt_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, scalar loops) of the reference's hot path,
 * JJJerome/mbt_gym `TradingEnvironment.step()/reset()`; every block in mbt_oracle_impl.h cites the
 * reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may load this library; nothing under mbt_gym_b200/ does, and the product has no CPU path.
 *
 * How it is pinned (see DESIGN.md "Oracle"):
 *   1. oracle/ref_shim.py runs the UNMODIFIED reference from /root/reference with its numpy Generators
 *      replaced by an object that serves THIS file's Philox draws, and checks every state column,
 *      reward and done flag of orc_step_core_f64 against it, trajectory by trajectory; the resulting
 *      vectors are committed under tests/golden/ (the reference is not present on the GPU box).
 *   2. orc_reward_eval reproduces the reference's own five unit tests
 *      (mbt_gym/rewards/tests/testRewardFunctions.py:33-135) in tests/test_oracle_rewards.py.
 *   3. statistically, against the notebook goldens (Test_1 AS table, Test_2 CJP value function).
 *
 * The only product headers it shares are the RNG / math primitives (include/mbt_philox.h,
 * include/mbt_math.h), which are themselves checked against Random123 known answers and libm.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mbt_b200.h" /* mbt_config POD + enums only */
#include "../include/mbt_math.h"
#include "../include/mbt_philox.h"

/* action dim A, observation dim D (= state columns of the reference's matrix) */
static int orc_dims(const mbt_config *c, int32_t *A, int32_t *D, int32_t *S) {
    int a, d = 4; /* cash, inventory, time, midprice            TradingEnvironment.py:131-140 */
    switch (c->dynamics) {
    case MBT_DYN_LIMIT: a = 2; break;            /* ModelDynamics.py:118-121 */
    case MBT_DYN_AT_TOUCH: a = 2; break;         /* ModelDynamics.py:166-167 */
    case MBT_DYN_LIMIT_AND_MARKET: a = 4; break; /* ModelDynamics.py:226-232 */
    case MBT_DYN_SPEED: a = 1; break;            /* ModelDynamics.py:269-271 */
    default: return -1;
    }
    if (c->midprice == MBT_MID_HESTON) d += 1;   /* (price, variance)  midprice_models.py:346 */
    if (c->arrival == MBT_ARR_HAWKES) d += 2;    /* arrival_models.py:99-103 */
    if (c->fill == MBT_FILL_EXOGENOUS_MM && (c->dynamics == MBT_DYN_LIMIT || c->dynamics == MBT_DYN_LIMIT_AND_MARKET))
        d += 2; /* exogenous best depths (bid, ask)  fill_probability_models.py:144-152 */
    if (c->impact == MBT_IMP_TEMP_PERM || c->impact == MBT_IMP_TEMP_TRANSIENT || c->impact == MBT_IMP_TRANSIENT)
        d += 1; /* one state column: price_impact_models.py:79-83,119-127,162-170 */
    *A = a;
    *D = d;
    *S = d - 1;
    return 0;
}

static int orc_obs_out_dim(const mbt_config *c, int D) {
    if (!c->obs_select) return D;
    int n = 0;
    for (int d = 0; d < D; ++d) n += (c->obs_select >> d) & 1u;
    return n;
}

#define REAL double
#define SFX f64
#define ORC_IS_F64 1
#include "mbt_oracle_impl.h"
#undef REAL
#undef SFX
#undef ORC_IS_F64

#define REAL float
#define SFX f32
#define ORC_IS_F64 0
#include "mbt_oracle_impl.h"
#undef REAL
#undef SFX
#undef ORC_IS_F64

/* ------------------------------------------------------------------ exported, precision-dispatched API */
typedef struct orc_handle {
    int precision;
    orc_env_f64 *d;
    orc_env_f32 *f;
} orc_handle;

int orc_config_dims(const mbt_config *c, int32_t *A, int32_t *D, int32_t *S) { return orc_dims(c, A, D, S); }
int orc_config_obs_out_dim(const mbt_config *c) {
    int32_t A, D, S;
    if (orc_dims(c, &A, &D, &S) != 0) return -1;
    return orc_obs_out_dim(c, D);
}

orc_handle *orc_create(const mbt_config *cfg) {
    if (!cfg || cfg->struct_size != (int32_t)sizeof(mbt_config)) return NULL;
    orc_handle *h = (orc_handle *)calloc(1, sizeof *h);
    h->precision = cfg->precision;
    if (cfg->precision == MBT_F64) h->d = orc_create_f64(cfg); else h->f = orc_create_f32(cfg);
    if (!h->d && !h->f) { free(h); return NULL; }
    return h;
}
void orc_destroy(orc_handle *h) {
    if (!h) return;
    orc_destroy_f64(h->d);
    orc_destroy_f32(h->f);
    free(h);
}
void orc_seed(orc_handle *h, uint64_t seed) {
    if (h->d) { h->d->seed = seed; h->d->n_step = 0; h->d->n_episode = 0; }
    if (h->f) { h->f->seed = seed; h->f->n_step = 0; h->f->n_episode = 0; }
}
void orc_reset(orc_handle *h, const mbt_reset_args *args, void *obs_out) {
    if (h->d) orc_reset_f64(h->d, args, (double *)obs_out); else orc_reset_f32(h->f, args, (float *)obs_out);
}
void orc_step(orc_handle *h, const void *actions, void *obs_out, void *rew_out, uint8_t *done_out) {
    if (h->d) orc_step_f64(h->d, (const double *)actions, (double *)obs_out, (double *)rew_out, done_out);
    else orc_step_f32(h->f, (const float *)actions, (float *)obs_out, (float *)rew_out, done_out);
}
/* step with caller-supplied random numbers: u (N,4), z (N,) in the handle's precision */
void orc_step_draws(orc_handle *h, const void *actions, const void *u, const void *z, void *obs_out, void *rew_out,
                    uint8_t *done_out) {
    /* the second normal (Heston only) always comes from the draw contract here */
    if (h->d) {
        double *z2 = h->d->cfg.midprice == MBT_MID_HESTON ? (double *)malloc(sizeof(double) * (size_t)h->d->N) : NULL;
        if (z2) orc_fill_draws_f64(h->d->seed, h->d->cfg.traj_offset, h->d->N, h->d->n_step, NULL, NULL, z2);
        orc_step_core_f64(h->d, (const double *)actions, (const double *)u, (const double *)z, z2, (double *)obs_out,
                          (double *)rew_out, done_out);
        free(z2);
    } else {
        float *z2 = h->f->cfg.midprice == MBT_MID_HESTON ? (float *)malloc(sizeof(float) * (size_t)h->f->N) : NULL;
        if (z2) orc_fill_draws_f32(h->f->seed, h->f->cfg.traj_offset, h->f->N, h->f->n_step, NULL, NULL, z2);
        orc_step_core_f32(h->f, (const float *)actions, (const float *)u, (const float *)z, z2, (float *)obs_out,
                          (float *)rew_out, done_out);
        free(z2);
    }
}
/* the draws step `n_step` would use (so a test can hand the SAME numbers to the reference) */
void orc_draws(int precision, uint64_t seed, int64_t traj_offset, int64_t N, int64_t n_step, void *u, void *z) {
    if (precision == MBT_F64) orc_fill_draws_f64(seed, traj_offset, N, n_step, (double *)u, (double *)z, NULL);
    else orc_fill_draws_f32(seed, traj_offset, N, n_step, (float *)u, (float *)z, NULL);
}
/* the SECOND normal step `n_step` uses (Heston variance; stream MBT_STREAM_STEP2 of the draw contract) */
void orc_draws2(int precision, uint64_t seed, int64_t traj_offset, int64_t N, int64_t n_step, void *z2) {
    if (precision == MBT_F64) orc_fill_draws_f64(seed, traj_offset, N, n_step, NULL, NULL, (double *)z2);
    else orc_fill_draws_f32(seed, traj_offset, N, n_step, NULL, NULL, (float *)z2);
}
/* the initial inventories reset number `n_episode` draws for MBT_Q0_UNIFORM_INT (same formula as orc_reset) */
void orc_q0_draws(uint64_t seed, int64_t traj_offset, int64_t N, int64_t n_episode, int64_t lo, int64_t hi, int64_t *out) {
    for (int64_t i = 0; i < N; ++i) {
        mbt_u32x4 r = mbt_draw(seed, (uint64_t)(traj_offset + i), (uint64_t)n_episode, MBT_STREAM_RESET);
        out[i] = lo + (int64_t)(((uint64_t)r.x * (uint64_t)(hi - lo)) >> 32);
    }
}
void orc_get_state(orc_handle *h, void *out) {
    if (h->d) memcpy(out, h->d->state, sizeof(double) * (size_t)h->d->N * h->d->D);
    else memcpy(out, h->f->state, sizeof(float) * (size_t)h->f->N * h->f->D);
}
void orc_set_state(orc_handle *h, const void *in) {
    if (h->d) { memcpy(h->d->state, in, sizeof(double) * (size_t)h->d->N * h->d->D); h->d->t = h->d->state[2]; }
    else { memcpy(h->f->state, in, sizeof(float) * (size_t)h->f->N * h->f->D); h->f->t = (double)h->f->state[2]; }
}
void orc_get_clock(orc_handle *h, double *t, int64_t *k, int64_t *n_step, int64_t *n_episode, int64_t *clipped) {
    if (h->d) { *t = h->d->t; *k = h->d->k; *n_step = h->d->n_step; *n_episode = h->d->n_episode; *clipped = h->d->clipped; }
    else { *t = h->f->t; *k = h->f->k; *n_step = h->f->n_step; *n_episode = h->f->n_episode; *clipped = h->f->clipped; }
}
/* reward_function.calculate on raw rows (mirrors the reference's unit tests); q0 / L as captured by reset */
void orc_reward_eval(const mbt_config *c, int64_t n, const void *cur, const void *act, const void *next,
                     int is_terminal, double q0, double L, void *out) {
    int32_t A = 0, D = 0, S = 0;
    if (orc_dims(c, &A, &D, &S) != 0) return;
    for (int64_t i = 0; i < n; ++i) {
        if (c->precision == MBT_F64) {
            const double *cs = (const double *)cur + i * D, *s = (const double *)next + i * D;
            ((double *)out)[i] = orc_reward_f64(c, cs, s, (const double *)act + i * A, is_terminal, s[2] - cs[2], q0, L);
        } else {
            const float *cs = (const float *)cur + i * D, *s = (const float *)next + i * D;
            ((float *)out)[i] = orc_reward_f32(c, cs, s, (const float *)act + i * A, is_terminal, s[2] - cs[2], (float)q0, (float)L);
        }
    }
}

/* ------------------------------------------------------------------ primitive wrappers (tests) */
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    mbt_u32x4 c = {ctr[0], ctr[1], ctr[2], ctr[3]};
    mbt_u32x4 r = mbt_philox4x32_10(c, key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
void orc_vec_exp_f32(int64_t n, const float *x, float *y) { for (int64_t i = 0; i < n; ++i) y[i] = mbt_exp_f32(x[i]); }
void orc_vec_log_f32(int64_t n, const float *x, float *y) { for (int64_t i = 0; i < n; ++i) y[i] = mbt_log_f32(x[i]); }
void orc_vec_exp_f64(int64_t n, const double *x, double *y) { for (int64_t i = 0; i < n; ++i) y[i] = mbt_exp_f64(x[i]); }
void orc_vec_log_f64(int64_t n, const double *x, double *y) { for (int64_t i = 0; i < n; ++i) y[i] = mbt_log_f64(x[i]); }
void orc_vec_normal_f32(int64_t n, const uint32_t *b, float *y) { for (int64_t i = 0; i < n; ++i) y[i] = mbt_normal_from_bits_f32(b[i]); }
void orc_vec_normal_f64(int64_t n, const uint32_t *b, double *y) { for (int64_t i = 0; i < n; ++i) y[i] = mbt_normal_from_bits_f64(b[i]); }
/* the kernels' division-through-reciprocal (include/mbt_math.h) next to the plain division, for tests/test_primitives.py */
int64_t orc_count_div_rcp_mismatch_f64(int64_t n, const double *a, const double *b) {
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        double q = mbt_div_rcp_f64(a[i], b[i], mbt_rcp_for_div_f64(b[i])), w = a[i] / b[i];
        bad += memcmp(&q, &w, sizeof q) != 0 && !(q != q && w != w);
    }
    return bad;
}
int64_t orc_count_div_rcp_mismatch_f32(int64_t n, const float *a, const float *b) {
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        float q = mbt_div_rcp_f32(a[i], b[i], mbt_rcp_for_div_f32(b[i])), w = a[i] / b[i];
        bad += memcmp(&q, &w, sizeof q) != 0 && !(q != q && w != w);
    }
    return bad;
}
/* the float-filtered fill decision (mbt_u24_below_exp_f64) next to its definition, and the float estimate's relative error */
int64_t orc_count_fill_filter_mismatch(int64_t n, const uint32_t *k, const double *x, double *max_rel_err_out) {
    int64_t bad = 0;
    double worst = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const int want = mbt_u24_to_real_f64(k[i]) < mbt_exp2k_f64(x[i], 24);
        bad += mbt_u24_below_exp_f64(k[i], x[i]) != want;
        if (x[i] <= 0.0 && x[i] >= -16.0) {
            const double exact = exp(x[i]) * 16777216.0, est = (double)mbt_exp2k_f32((float)x[i], 24);
            const double rel = fabs(est - exact) / exact;
            if (rel > worst) worst = rel;
        }
    }
    if (max_rel_err_out) *max_rel_err_out = worst;
    return bad;
}
void orc_vec_pow_f64(int64_t n, const double *x, double p, double *y) { for (int64_t i = 0; i < n; ++i) y[i] = mbt_pow_f64(x[i], p); }
