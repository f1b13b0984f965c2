/*
 * hostsim.cpp -- TEST-ONLY host compile of the kernel's per-trajectory core (mbt_step_core.cuh).
 *
 * The build container has no GPU; compiling step_one<T,V>() for the host lets tests/test_hostsim.py check
 * the exact expression trees the kernels execute against the oracle before any GPU time is spent.  This is
 * NOT a CPU path of the product: it lives under tests/, is built only by the test, and nothing in
 * mbt_gym_b200/ can reach it.
 */
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../mbt_gym_b200/csrc/mbt_host_params.h"
#include "../../mbt_gym_b200/csrc/mbt_variants.h" /* MBT_FOR_EACH_VARIANT, variant_of: the table mbt_capi.cu launches from */

template <typename T, class V>
static void run(const mbt_config &c, uint64_t seed, int64_t n_step0, double t0, double t_start, int q0_per_traj,
                double q0_uniform, T *state, const T *q0, int steps, const T *actions, T *obs, T *rew,
                uint8_t *dones) {
    int32_t A, D, S;
    mbt_dims(&c, &A, &D, &S);
    const int64_t N = c.num_trajectories;
    StepParams<T> p = mbt_make_params<T>(c, t0, q0_per_traj, q0_uniform);
    const mbt_philox_keys keys = mbt_philox_expand(seed);
    double t = t_start;
    for (int k = 0; k < steps; ++k) {
        const double t_next = t + c.step_size;
        StepClock<T> ck = mbt_make_clock<T>(c, t, t_next, t0);
        /* what mbt_fill_batch_kernel does in front of the step kernel: deepest quote of the batch -> thresholds */
        T fill_thr[2] = {0, 0};
        if (fill_is_batch(c.fill) && c.dynamics != MBT_DYN_AT_TOUCH && c.dynamics != MBT_DYN_SPEED) {
            T m[2] = {0, 0};
            for (int64_t i = 0; i < N; ++i)
                for (int j = 0; j < 2; ++j) {
                    const T x = denorm_action<T, V>(p, actions[((int64_t)k * N + i) * A + j], j);
                    m[j] = i == 0 ? x : nanmax<T>(m[j], x);
                }
            fill_batch_thresholds<T>(p, m[0], m[1], fill_thr);
        }
        for (int64_t i = 0; i < N; ++i) {
            T *row = state + i * D;
            Traj<T> s;
            s.cash = row[0]; s.inv = row[1]; s.mid = row[3]; s.x0 = 0; s.x1 = 0; s.var = 0;
            const int mc = 4 + (c.midprice == MBT_MID_HESTON ? 1 : 0); /* first column of the arrival / impact model */
            if (c.midprice == MBT_MID_HESTON) s.var = row[4];
            if (c.arrival == MBT_ARR_HAWKES) { s.x0 = row[mc]; s.x1 = row[mc + 1]; }
            if (imp_has_state(c.impact)) s.x0 = row[mc];
            T a[MBT_MAX_ACTION_DIM] = {0, 0, 0, 0};
            for (int j = 0; j < A; ++j) a[j] = denorm_action<T, V>(p, actions[((int64_t)k * N + i) * A + j], j);
            mbt_u32x4 r = mbt_draw_keyed(keys, (uint64_t)(c.traj_offset + i), (uint64_t)(n_step0 + k), MBT_STREAM_STEP);
            int clipped = 0;
            uint32_t nbits2 = 0; /* second normal (Heston): what second_normal_bits() computes in the kernels */
            if (V::mid < 0 && c.midprice == MBT_MID_HESTON)
                nbits2 = mbt_normal_bits(mbt_draw_keyed(keys, (uint64_t)(c.traj_offset + i), (uint64_t)(n_step0 + k), MBT_STREAM_STEP2));
            T rw = step_one<T, V>(p, ck, s, a, r, q0_per_traj ? q0[i] : (T)q0_uniform, &clipped, fill_thr, nbits2);
            row[0] = s.cash; row[1] = s.inv; row[2] = ck.t_next; row[3] = s.mid;
            if (c.midprice == MBT_MID_HESTON) row[4] = s.var;
            if (c.arrival == MBT_ARR_HAWKES) { row[mc] = s.x0; row[mc + 1] = s.x1; }
            if (imp_has_state(c.impact)) row[mc] = s.x0;
            for (int d = 0; d < D; ++d) obs[((int64_t)k * N + i) * D + d] = norm_obs<T, V>(p, row[d], d);
            rew[(int64_t)k * N + i] = rw;
        }
        dones[k] = (uint8_t)ck.done;
        t = t_next;
    }
}

template <typename T>
static void dispatch(int variant, const mbt_config &c, uint64_t seed, int64_t n_step0, double t0, double t_start,
                     int q0_per_traj, double q0_uniform, void *state, const void *q0, int steps, const void *actions,
                     void *obs, void *rew, uint8_t *dones) {
#define GO(...) run<T, __VA_ARGS__>(c, seed, n_step0, t0, t_start, q0_per_traj, q0_uniform, (T *)state, (const T *)q0, steps, (const T *)actions, (T *)obs, (T *)rew, dones)
    switch (variant) {
#define X(id, ...) case id: GO(__VA_ARGS__); break;
        MBT_FOR_EACH_VARIANT(X)
#undef X
    default: GO(VariantGeneric); break;
    }
#undef GO
}

/* the variant the library would launch for this config (variant < 0 in hostsim_run selects it) */
extern "C" int hostsim_variant_of(const mbt_config *c) { return variant_of(*c); }

extern "C" int hostsim_run(const mbt_config *c, int variant, uint64_t seed, int64_t n_step0, double t0, double t_start,
                           int q0_per_traj, double q0_uniform, void *state, const void *q0, int steps,
                           const void *actions, void *obs, void *rew, uint8_t *dones) {
    std::string err;
    int rc = mbt_validate_config(c, err);
    if (rc) return rc;
    if (c->precision == MBT_F64)
        dispatch<double>(variant, *c, seed, n_step0, t0, t_start, q0_per_traj, q0_uniform, state, q0, steps, actions, obs, rew, dones);
    else
        dispatch<float>(variant, *c, seed, n_step0, t0, t_start, q0_per_traj, q0_uniform, state, q0, steps, actions, obs, rew, dones);
    return 0;
}
