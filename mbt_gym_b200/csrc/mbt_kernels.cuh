/*
 * mbt_kernels.cuh -- the sm_100a kernels of the hot path.
 *
 *   mbt_step_kernel     one env-step for all trajectories        (TradingEnvironment.step,  :103-110)
 *   mbt_reset_kernel    start an episode                         (TradingEnvironment.reset, :96-101)
 *   mbt_rollout_kernel  whole episode, policy on device, state in registers
 *                                                                (generate_trajectory, helpers/generate_trajectory.py:8-38)
 *   gather/scatter      (N,D) row-major <-> structure-of-arrays  (env.state property)
 *   mbt_reward_kernel   RewardFunction.calculate on caller rows  (RewardFunctions.py:8-13)
 *
 * Data layout in HBM (DESIGN.md "Layout"): the persistent per-trajectory state is structure-of-arrays,
 * one contiguous column per scalar (cash, inventory, midprice, then model columns), so a warp's access to
 * a column is one fully-coalesced 128 B (float) / 256 B (double) transaction.  Time is NOT stored per
 * trajectory: the reference's clock is uniform over the batch (TradingEnvironment.py:216-220), so it lives
 * in the kernel arguments.  Caller-facing buffers keep the reference's layouts: actions (N,A), observations
 * (N,D) row-major; each thread reads/writes its row with the widest aligned vector access (128-bit, or the
 * sm_100 256-bit LDG/STG for a 4-double row).
 *
 * No tensor cores: the path is elementwise + counter-based RNG, bound by HBM bytes and INT32/FP issue.
 */
#ifndef MBT_KERNELS_CUH
#define MBT_KERNELS_CUH

#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif

#include "mbt_step_core.cuh"

#ifndef MBT_BLOCK_THREADS
#define MBT_BLOCK_THREADS 256 /* tuned on B200: 128 / 256 / 512 measured in profiles/r1_step_kernel_history.md */
#endif
constexpr int MBT_BLOCK = MBT_BLOCK_THREADS;

template <typename T>
struct DevState {
    T *cash, *inv, *mid, *x0, *x1, *q0, *var; /* var: Heston variance */
};

/* T = arithmetic / state type; E = element type of the CALLER's buffers (E = T, or float over double arithmetic:
 * `io_precision = MBT_IO_F32`, values converted with round-to-nearest on the way in / out) */
template <typename T, typename E = T>
struct StepArgs {
    StepParams<T> p;
    StepClock<T> ck;
    DevState<T> st;
    const E *actions; /* (N, A) */
    E *obs;           /* (N, D) or NULL */
    E *rew;           /* (N,)   or NULL */
    long long n;
    mbt_philox_keys keys; /* the ten Philox round keys of the seed, expanded on the host */
    unsigned long long traj_offset, n_step;
    unsigned long long *clipped;
    /* batch-reduced fill models: the two running maxima of the quoted depths (order-preserving keys) that
     * mbt_fill_batch_kernel left in front of this launch, and the ticket by which the last block clears them */
    unsigned long long *fill_cells;
    unsigned int *fill_ticket;
    /* CUDA-graph replay (mbt_fold_counters): NULL, or the device-resident base {steps, episodes} that is added to the
     * counters baked into this launch -- a captured episode replays with fresh random numbers every time */
    const unsigned long long *counter_base;
};

/* the specialised variants know the row widths at compile time */
template <typename T, class V>
__device__ __forceinline__ int action_width(const StepParams<T> &p) { return V::A ? V::A : p.action_dim; }
template <typename T, class V>
__device__ __forceinline__ int obs_width(const StepParams<T> &p) { return V::Dout ? V::Dout : p.obs_out_dim; }

/* ------------------------------------------------------------------ row access helpers */
#ifndef MBT_NO_256BIT
__device__ __forceinline__ void st_v4_f64(double *ptr, double a, double b, double c, double d) {
    /* sm_100: 256-bit global store (SASS STG.E.ENL2.256), 32-byte aligned */
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(ptr), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld_v4_f64(const double *ptr, double &a, double &b, double &c, double &d) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(ptr));
}
#else /* an NVRTC older than CUDA 12.9 (PTX ISA < 8.8) has no 256-bit vector accesses: two 128-bit ones */
__device__ __forceinline__ void st_v4_f64(double *ptr, double a, double b, double c, double d) {
    reinterpret_cast<double2 *>(ptr)[0] = make_double2(a, b);
    reinterpret_cast<double2 *>(ptr)[1] = make_double2(c, d);
}
__device__ __forceinline__ void ld_v4_f64(const double *ptr, double &a, double &b, double &c, double &d) {
    const double2 lo = __ldg(reinterpret_cast<const double2 *>(ptr)), hi = __ldg(reinterpret_cast<const double2 *>(ptr) + 1);
    a = lo.x; b = lo.y; c = hi.x; d = hi.y;
}
#endif

template <typename T>
__device__ __forceinline__ void load_row(const T *__restrict__ base, long long i, int w, T *out, bool vec_ok) {
    const T *row = base + i * w;
    if (vec_ok && w == 2) {
        if constexpr (sizeof(T) == 4) {
            float2 v = __ldg(reinterpret_cast<const float2 *>(row));
            out[0] = v.x; out[1] = v.y;
        } else {
            double2 v = __ldg(reinterpret_cast<const double2 *>(row));
            out[0] = v.x; out[1] = v.y;
        }
    } else if (vec_ok && w == 4) {
        if constexpr (sizeof(T) == 4) {
            float4 v = __ldg(reinterpret_cast<const float4 *>(row));
            out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
        } else {
            double a, b, c, d;
            ld_v4_f64(reinterpret_cast<const double *>(row), a, b, c, d);
            out[0] = a; out[1] = b; out[2] = c; out[3] = d;
        }
    } else {
#pragma unroll
        for (int j = 0; j < MBT_MAX_OBS_DIM; ++j)
            if (j < w) out[j] = __ldg(row + j);
    }
}

template <typename T>
__device__ __forceinline__ void store_row(T *__restrict__ base, long long i, int w, const T *v, bool vec_ok) {
    T *row = base + i * w;
    if (vec_ok && w == 4) {
        if constexpr (sizeof(T) == 4) {
            *reinterpret_cast<float4 *>(row) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            st_v4_f64(reinterpret_cast<double *>(row), v[0], v[1], v[2], v[3]);
        }
    } else if (vec_ok && w == 2) {
        if constexpr (sizeof(T) == 4) {
            *reinterpret_cast<float2 *>(row) = make_float2(v[0], v[1]);
        } else {
            *reinterpret_cast<double2 *>(row) = make_double2(v[0], v[1]);
        }
    } else if (vec_ok && w == 6) {
        if constexpr (sizeof(T) == 4) {
            float2 *r2 = reinterpret_cast<float2 *>(row);
            r2[0] = make_float2(v[0], v[1]); r2[1] = make_float2(v[2], v[3]); r2[2] = make_float2(v[4], v[5]);
        } else {
            double2 *r2 = reinterpret_cast<double2 *>(row);
            r2[0] = make_double2(v[0], v[1]); r2[1] = make_double2(v[2], v[3]); r2[2] = make_double2(v[4], v[5]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < MBT_MAX_OBS_DIM; ++j)
            if (j < w) row[j] = v[j];
    }
}

/*
 * Observation rows whose width is not a power of two (D = 5: price-impact column, D = 6: Hawkes intensities) cannot
 * be written with one aligned vector per thread; per-thread scalar stores would touch every 32 B sector of the warp's
 * 32 x D block several times.  Instead the warp stages its block in shared memory and writes it back as contiguous
 * runs -- lane l stores elements l, l+32, ... of the block -- so every store instruction covers 32 consecutive elements
 * (full 128 B / 256 B lines).  Only used when all 32 lanes hold valid rows (the last partial warp stores per thread).
 */
template <typename T>
__device__ __forceinline__ void store_rows_staged(T *__restrict__ base, long long warp_row0, int w, const T *v,
                                                  T *warp_smem, unsigned lane) {
#pragma unroll
    for (int j = 0; j < MBT_MAX_OBS_DIM; ++j)
        if (j < w) warp_smem[lane * w + j] = v[j];
    __syncwarp();
    T *dst = base + warp_row0 * w;
#pragma unroll
    for (int j = 0; j < MBT_MAX_OBS_DIM; ++j)
        if (j < w) dst[j * 32 + lane] = warp_smem[j * 32 + lane];
    __syncwarp();
}

/* observation row of one trajectory: [cash, inventory, time, midprice | arrival cols | impact col]
 * (TradingEnvironment.py:131-140,303-318), normalised on the way out (:112-118).  `t_obs` is the time column's value as
 * emitted (the host normalises the uniform clock: StepClock::t_obs, mbt_time_obs). */
template <typename T, class V>
__device__ __forceinline__ int make_obs_row(const StepParams<T> &p, const Traj<T> &s, T t_obs, T *row) {
    const int arr = pick<V::arr>(p.arr), imp = pick<V::imp>(p.imp);
    row[0] = norm_obs<T, V>(p, s.cash, 0);
    row[1] = norm_obs<T, V>(p, s.inv, 1);
    row[2] = t_obs;
    row[3] = norm_obs<T, V>(p, s.mid, 3);
    /* model columns at FIXED positions (no `row[d++]` with a runtime d: that would put the row in local memory).  Hawkes
     * intensities and a price-impact column never coexist: Hawkes needs limit-order dynamics, impact models need speed
     * dynamics (mbt_validate_config) */
    int d = 4;
    if (pick<V::mid>(p.mid) == MBT_MID_HESTON) { /* the midprice model owns two columns: (price, variance) */
        row[4] = norm_obs<T, V>(p, s.var, 4);
        d = 5;
        if (arr == MBT_ARR_HAWKES) {
            row[5] = norm_obs<T, V>(p, s.x0, 5);
            row[6] = norm_obs<T, V>(p, s.x1, 6);
            d = 7;
        } else if (imp_has_state(imp)) {
            row[5] = norm_obs<T, V>(p, s.x0, 5);
            d = 6;
        }
    } else if (arr == MBT_ARR_HAWKES) {
        row[4] = norm_obs<T, V>(p, s.x0, 4);
        row[5] = norm_obs<T, V>(p, s.x1, 5);
        d = 6;
    } else if (imp_has_state(imp)) {
        row[4] = norm_obs<T, V>(p, s.x0, 4);
        d = 5;
    }
    if (pick<V::fill>(p.fill) == MBT_FILL_EXOGENOUS_MM) { /* (p.fill is MBT_FILL_NONE for dynamics that draw no fills) */
        /* the fill model's two columns (constants, see include/mbt_b200.h) come after the arrival model's; `d` is a
         * runtime value here, so each is placed by a compile-time-indexed select chain (no dynamic register indexing) */
        const T c0 = norm_obs<T, V>(p, p.fill_depth0[0], d), c1 = norm_obs<T, V>(p, p.fill_depth0[1], d + 1);
#pragma unroll
        for (int k = 4; k < MBT_MAX_OBS_DIM; ++k) {
            if (k == d) row[k] = c0;
            if (k == d + 1) row[k] = c1;
        }
        d += 2;
    }
    const int sel = pick<V::sel>(p.obs_select);
    if (sel) { /* keep the selected columns, in order (gym/wrappers.py:30-38) */
        /* compaction without dynamically indexed registers (a runtime `row[j++]` would push the whole row into local
         * memory for every launch of the runtime-flag variants, selecting or not): output slot j takes column k when k is
         * the j-th set bit of the mask -- all indices below are compile-time after unrolling, the tests are warp-uniform */
        T out[MBT_MAX_OBS_DIM];
        int j = 0;
#pragma unroll
        for (int k = 0; k < MBT_MAX_OBS_DIM; ++k) {
            const bool take = k < d && ((sel >> k) & 1);
#pragma unroll
            for (int slot = 0; slot < MBT_MAX_OBS_DIM; ++slot)
                if (slot <= k && take && slot == j) out[slot] = row[k];
            j += take ? 1 : 0;
        }
#pragma unroll
        for (int slot = 0; slot < MBT_MAX_OBS_DIM; ++slot)
            if (slot < j) row[slot] = out[slot];
        return j;
    }
    return d;
}

template <typename T, class V>
__device__ __forceinline__ void load_traj(const StepParams<T> &p, const DevState<T> &st, long long i, Traj<T> &s) {
    const int arr = pick<V::arr>(p.arr), imp = pick<V::imp>(p.imp);
    s.cash = st.cash[i];
    s.inv = st.inv[i];
    s.mid = st.mid[i];
    s.x0 = (T)0;
    s.x1 = (T)0;
    if (arr == MBT_ARR_HAWKES) { s.x0 = st.x0[i]; s.x1 = st.x1[i]; }
    if (imp_has_state(imp)) s.x0 = st.x0[i];
    s.var = (T)0;
    if (pick<V::mid>(p.mid) == MBT_MID_HESTON) s.var = st.var[i];
}

template <typename T, class V>
__device__ __forceinline__ void store_traj(const StepParams<T> &p, const DevState<T> &st, long long i, const Traj<T> &s) {
    const int arr = pick<V::arr>(p.arr), imp = pick<V::imp>(p.imp), mid = pick<V::mid>(p.mid);
    st.cash[i] = s.cash;
    st.inv[i] = s.inv;
    if (mid != MBT_MID_CONSTANT) st.mid[i] = s.mid;
    if (arr == MBT_ARR_HAWKES) { st.x0[i] = s.x0; st.x1[i] = s.x1; }
    if (imp_has_state(imp)) st.x0[i] = s.x0;
    if (mid == MBT_MID_HESTON) st.var[i] = s.var;
}

/* the 32 normal bits of the step's SECOND Philox block, for the models that need a second normal (Heston) */
template <typename T, class V>
__device__ __forceinline__ uint32_t second_normal_bits(const StepParams<T> &p, const mbt_philox_keys &keys,
                                                       unsigned long long traj, unsigned long long n_step) {
    if (pick<V::mid>(p.mid) != MBT_MID_HESTON) return 0u;
    return mbt_normal_bits(mbt_draw_keyed(keys, traj, n_step, MBT_STREAM_STEP2));
}

/* order-preserving key of a real (float widens exactly): key(a) < key(b) <=> a < b; every NaN -> the largest key */
__device__ __forceinline__ unsigned long long real_to_key(double x) {
    if (x != x) return ~0ull;
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_real(unsigned long long k) {
    if (k == ~0ull) return __longlong_as_double(0x7ff8000000000000ll); /* NaN */
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}

/* ------------------------------------------------------------------ step */
/* One trajectory, one env-step: load state + action row, draw, advance, store state + observation row + reward. */
/* Programmatic dependent launch (sm_90+): `launch_dependents` lets the NEXT kernel in the stream start launching while
 * this one is still running; `wait` blocks until the PREVIOUS kernel has completed and its writes are visible.  Both are
 * no-ops when the kernel was launched without the programmatic-serialization attribute. */
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename T, typename E, class V, bool VEC>
__device__ __forceinline__ void step_row(const StepArgs<T, E> &g, long long i, bool full_warp, E *warp_smem) {
    const StepParams<T> &p = g.p;
    const bool live = i < g.n; /* (the threads past the end still take part in the block-level stage below) */
    /* state-independent prologue: the step's 128 random bits depend only on (seed, trajectory id, step index) */
    unsigned long long n_step = g.n_step;
    if (g.counter_base) { /* warp-uniform; such launches are never programmatic (the base is written by a kernel) */
        pdl_wait();
        n_step += g.counter_base[0];
    }
    const mbt_u32x4 r = mbt_draw_keyed(g.keys, g.traj_offset + (unsigned long long)i, n_step, MBT_STREAM_STEP);
    const uint32_t nbits2 = second_normal_bits<T, V>(p, g.keys, g.traj_offset + (unsigned long long)i, n_step);
    pdl_wait(); /* everything below reads what the previous kernel (previous step, the batch reduction, or the caller's policy) wrote */

    /* the row's own loads are issued first: they do not depend on the batch thresholds computed below */
    const int A = action_width<T, V>(p);
    E a_io[MBT_MAX_ACTION_DIM] = {0, 0, 0, 0};
    Traj<T> s;
    s.cash = s.inv = s.mid = s.x0 = s.x1 = s.var = (T)0;
    const int rew_kind = pick<V::rew>(p.rew);
    T q_init = p.q0_uniform;
    if (live) {
        load_row<E>(g.actions, i, A, a_io, VEC);
        load_traj<T, V>(p, g.st, i, s);
        if ((rew_kind == MBT_REW_CJ_MM || rew_kind == MBT_REW_CJ_OE) && p.q0_per_traj) q_init = g.st.q0[i];
    }

    T fill_thr[2] = {(T)0, (T)0};
    if (fill_is_batch(pick<V::fill>(p.fill))) {
        /* per block, two threads turn the batch maxima (keys left by mbt_fill_batch_kernel) into the step's two fill
         * thresholds -- one pow each for the power function, side by side -- and share them */
        __shared__ T s_thr[2];
        if (threadIdx.x < 2) {
            const T m_own = (T)key_to_real(__ldcg(g.fill_cells + threadIdx.x));
            const T m_other = (T)key_to_real(__ldcg(g.fill_cells + (threadIdx.x ^ 1u)));
            s_thr[threadIdx.x] = fill_batch_threshold_side<T>(p, m_own, m_other, (int)threadIdx.x);
        }
        __syncthreads();
        fill_thr[0] = s_thr[0];
        fill_thr[1] = s_thr[1];
        /* the last block to get here has seen every other block read the cells: clear them for the next step's reduction */
        if (threadIdx.x == 0 && atomicAdd(g.fill_ticket, 1u) == gridDim.x - 1) {
            g.fill_cells[0] = 0ull;
            g.fill_cells[1] = 0ull;
            *g.fill_ticket = 0u;
        }
    }
    if (!live) return;

    T a[MBT_MAX_ACTION_DIM] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < MBT_MAX_ACTION_DIM; ++j)
        if (j < A) a[j] = denorm_action<T, V>(p, (T)a_io[j], j);

    int clipped = 0;
    const T rwd = step_one<T, V>(p, g.ck, s, a, r, q_init, &clipped, fill_thr, nbits2);

    store_traj<T, V>(p, g.st, i, s);
    if (g.obs) {
        T row[MBT_MAX_OBS_DIM];
        make_obs_row<T, V>(p, s, g.ck.t_obs, row);
        const int D = obs_width<T, V>(p);
        E row_io[MBT_MAX_OBS_DIM];
#pragma unroll
        for (int d = 0; d < MBT_MAX_OBS_DIM; ++d)
            if (d < D) row_io[d] = (E)row[d];
        if (D != 4 && D != 2 && full_warp)
            store_rows_staged<E>(g.obs, i - (long long)(threadIdx.x & 31u), D, row_io, warp_smem, threadIdx.x & 31u);
        else
            store_row<E>(g.obs, i, D, row_io, VEC);
    }
    if (g.rew) g.rew[i] = (E)rwd;
    if (clipped) atomicAdd(g.clipped, 1ull);
}

/*
 * One thread per trajectory, 256-thread blocks.  Two alternatives were measured in round 1 and dropped
 * (profiles/r1_step_kernel_history.md): a persistent single-wave warp-tile loop (cost 8-12 registers, occupancy
 * 64 -> 40-48 warps/SM, slower) and two trajectories per thread (no change).  The marginal cost per trajectory
 * already equals the measured HBM rate; what is left at N = 2^20 is a fixed ~5 us of launch / ramp / drain.
 * VEC: the caller's action/obs pointers are aligned for whole-row vector access.
 */
template <typename T, typename E, class V, bool VEC>
__device__ __forceinline__ void mbt_step_body(const StepArgs<T, E> &g) {
    /* staging for non-power-of-two observation rows: one 32 x MBT_MAX_OBS_DIM tile per warp (unused when D == 4) */
    constexpr bool FIXED_W = V::Dout > 0;                    /* emitted row width known at compile time */
    constexpr int SW = FIXED_W ? V::Dout : MBT_MAX_OBS_DIM;  /* staged row width */
    constexpr bool NO_STAGE = FIXED_W && (V::Dout == 4 || V::Dout == 2 || V::Dout == 1); /* one aligned vector per row */
    __shared__ E smem[NO_STAGE ? 1 : (MBT_BLOCK / 32) * 32 * SW];
    const long long i = (long long)blockIdx.x * MBT_BLOCK + threadIdx.x;
    const long long warp_row0 = i - (long long)(threadIdx.x & 31u);
    const bool full_warp = !NO_STAGE && (warp_row0 + 32 <= g.n); /* warp-uniform */
    E *warp_smem = NO_STAGE ? smem : smem + (threadIdx.x >> 5) * 32 * SW;
    pdl_launch_dependents();
    step_row<T, E, V, VEC>(g, i, full_warp, warp_smem);
}

/* ahead-of-time instantiations (mbt_variants.h); the run-time specialiser wraps mbt_step_body in extern "C" kernels */
template <typename T, typename E, class V, bool VEC>
__global__ void __launch_bounds__(MBT_BLOCK) mbt_step_kernel(const __grid_constant__ StepArgs<T, E> g) {
    mbt_step_body<T, E, V, VEC>(g);
}

/* ------------------------------------------------------------------ batch reduction in front of the step */
/*
 * Triangular / Power fill functions (fill_probability_models.py:82,113): the reference's `np.max(depths, 0)` reduces
 * over the trajectory axis, so a step's fill probability depends on the deepest quote of the whole batch.  This kernel
 * is that reduction: NaN-propagating max of the (de-normalised) depth columns, eight independent row loads in flight per
 * thread, warp shuffle -> shared -> ONE 64-bit atomicMax per block and side on an order-preserving integer key of the
 * value (NaN = the largest key, like np.max).  max is order-independent, so the result is deterministic and equal to
 * numpy's.  The maxima stay in `cells` as keys: the step kernel that follows decodes them (one thread per block) and its
 * last block clears them; a group of handles all-reduces (max) the cells over NCCL in between.  Both kernels are launched
 * programmatically dependent: this one starts while the previous step drains and waits before it reads the actions; the
 * step's Philox prologue overlaps this kernel's tail.
 * Measured alternatives at N = 2^20, float64, reduction + step per env-step (profiles/r2_session_notes.md): thresholds by
 * this kernel's last block (ticket + fence, the round-1 design) 34.3 us; a third one-warp kernel between the two 33.0 us;
 * this design 25.5 us (round 1: 35 us) against 15.3 us for a market without the batch reduction -- every additional
 * grid-wide dependency costs more than the per-block threshold stage.
 */
template <typename T, typename E>
struct FillBatchArgs {
    StepParams<T> p;
    const E *actions; /* (N, A) */
    long long n;
    unsigned long long *cells; /* [2] running maxima as keys; zero on entry (cleared by the step kernel) */
};

template <typename T>
__device__ __forceinline__ T warp_nanmax(T v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = nanmax<T>(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

/* the two depths of row i: one 2-element vector load when the rows are aligned for it (A = 2 or 4), else two scalars */
template <typename E, bool VEC>
__device__ __forceinline__ void load_depths(const E *__restrict__ actions, long long i, int A, E &d0, E &d1) {
    const E *row = actions + i * A;
    if constexpr (VEC) {
        if constexpr (sizeof(E) == 4) {
            const float2 v = __ldg(reinterpret_cast<const float2 *>(row));
            d0 = v.x; d1 = v.y;
        } else {
            const double2 v = __ldg(reinterpret_cast<const double2 *>(row));
            d0 = v.x; d1 = v.y;
        }
    } else {
        d0 = __ldg(row + 0);
        d1 = __ldg(row + 1);
    }
}

template <typename T, typename E, bool VEC>
__global__ void __launch_bounds__(MBT_BLOCK) mbt_fill_batch_kernel(const __grid_constant__ FillBatchArgs<T, E> g) {
    const StepParams<T> &p = g.p;
    const int A = p.action_dim;
    const T neg_inf = -(T)INFINITY; /* identity of the NaN-propagating max */
    T m0 = neg_inf, m1 = neg_inf;
    constexpr int UNROLL = 8; /* independent row loads in flight per thread: few blocks (few atomics), deep loads */
    const long long nth = (long long)gridDim.x * MBT_BLOCK;
    pdl_launch_dependents();
    pdl_wait(); /* the actions (caller's policy / H2D copy) and the cleared cells (previous step kernel) are visible */
    for (long long i = (long long)blockIdx.x * MBT_BLOCK + threadIdx.x; i < g.n; i += UNROLL * nth) {
        E d0[UNROLL], d1[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { /* depths = action[:, 0:2]   ModelDynamics.py:50-51,128-130 */
            const long long r = i + u * nth;
            d0[u] = d1[u] = (E)0;
            if (r < g.n) load_depths<E, VEC>(g.actions, r, A, d0[u], d1[u]);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            if (i + u * nth < g.n) {
                m0 = nanmax<T>(m0, denorm_action<T, VariantGeneric>(p, (T)d0[u], 0));
                m1 = nanmax<T>(m1, denorm_action<T, VariantGeneric>(p, (T)d1[u], 1));
            }
    }
    __shared__ T sm[MBT_BLOCK / 32][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    m0 = warp_nanmax<T>(m0);
    m1 = warp_nanmax<T>(m1);
    if (lane == 0) { sm[warp][0] = m0; sm[warp][1] = m1; }
    __syncthreads();
    if (threadIdx.x >= 2) return;
    T b = sm[0][threadIdx.x]; /* thread 0: bid side, thread 1: ask side */
    for (int w = 1; w < MBT_BLOCK / 32; ++w) b = nanmax<T>(b, sm[w][threadIdx.x]);
    atomicMax(g.cells + threadIdx.x, real_to_key((double)b));
}

/* ------------------------------------------------------------------ reset */
template <typename T, typename E = T>
struct ResetArgs {
    StepParams<T> p; /* for normalisation + model kinds */
    DevState<T> st;
    E *obs;
    long long n;
    unsigned long long seed, traj_offset, n_episode;
    const unsigned long long *counter_base; /* see StepArgs */
    T cash0, t0_obs /* start time as the observation shows it */, mid0, lam0[2], imp0, var0;
    int q0_mode;
    T q0_const;
    long long q0_lo;
    unsigned long long q0_span;
};

template <typename T, typename E>
__global__ void __launch_bounds__(MBT_BLOCK) mbt_reset_kernel(const __grid_constant__ ResetArgs<T, E> g) {
    const long long i = (long long)blockIdx.x * MBT_BLOCK + threadIdx.x;
    if (i >= g.n) return;
    const StepParams<T> &p = g.p;
    Traj<T> s;
    s.cash = g.cash0;
    if (g.q0_mode == MBT_Q0_UNIFORM_INT) { /* rng.integers(lo, hi)   TradingEnvironment.py:271-272 */
        const unsigned long long n_episode = g.n_episode + (g.counter_base ? g.counter_base[1] : 0ull);
        const mbt_u32x4 r = mbt_draw(g.seed, g.traj_offset + (unsigned long long)i, n_episode, MBT_STREAM_RESET);
        s.inv = (T)(g.q0_lo + (long long)(((unsigned long long)r.x * g.q0_span) >> 32));
    } else if (g.q0_mode == MBT_Q0_PER_TRAJ) { /* a callable's array, uploaded into the q0 column   TradingEnvironment.py:275-279 */
        s.inv = g.st.q0[i];
    } else {
        s.inv = g.q0_const;
    }
    s.mid = g.mid0;
    s.x0 = (T)0;
    s.x1 = (T)0;
    s.var = g.var0; /* Heston initial_variance (unused otherwise) */
    if (p.arr == MBT_ARR_HAWKES) { s.x0 = g.lam0[0]; s.x1 = g.lam0[1]; }
    if (imp_has_state(p.imp)) s.x0 = g.imp0; /* 0 for the permanent impact, initial_transient_impact otherwise */
    g.st.cash[i] = s.cash;
    g.st.inv[i] = s.inv;
    g.st.mid[i] = s.mid;
    if (p.arr == MBT_ARR_HAWKES) { g.st.x0[i] = s.x0; g.st.x1[i] = s.x1; }
    if (imp_has_state(p.imp)) g.st.x0[i] = s.x0;
    if (p.mid == MBT_MID_HESTON) g.st.var[i] = s.var;
    if (g.q0_mode == MBT_Q0_UNIFORM_INT) g.st.q0[i] = s.inv; /* reward_function.reset  RewardFunctions.py:72,111 */
    if (g.obs) {
        T row[MBT_MAX_OBS_DIM];
        const int d = make_obs_row<T, VariantGeneric>(p, s, g.t0_obs, row);
        E row_io[MBT_MAX_OBS_DIM];
#pragma unroll
        for (int k = 0; k < MBT_MAX_OBS_DIM; ++k)
            if (k < d) row_io[k] = (E)row[k];
        store_row<E>(g.obs, i, d, row_io, false);
    }
}

/* ------------------------------------------------------------------ counters for CUDA-graph replay */
/* base[0] += steps, base[1] += episodes: what mbt_fold_counters enqueues (one thread; capturable) */
__global__ void mbt_fold_counters_kernel(unsigned long long *base, unsigned long long steps, unsigned long long episodes) {
    base[0] += steps;
    base[1] += episodes;
}

/* ------------------------------------------------------------------ state gather / scatter */
template <typename T>
__global__ void __launch_bounds__(MBT_BLOCK) mbt_gather_state_kernel(StepParams<T> p, DevState<T> st, T t, T *out, long long n) {
    const long long i = (long long)blockIdx.x * MBT_BLOCK + threadIdx.x;
    if (i >= n) return;
    Traj<T> s;
    load_traj<T, VariantGeneric>(p, st, i, s);
    StepParams<T> raw = p;
    raw.normalise_obs = 0;
    raw.obs_select = 0;
    T row[MBT_MAX_OBS_DIM];
    const int d = make_obs_row<T, VariantGeneric>(raw, s, t, row);
    store_row<T>(out, i, d, row, false);
}

template <typename T>
__global__ void __launch_bounds__(MBT_BLOCK) mbt_scatter_state_kernel(StepParams<T> p, DevState<T> st, const T *in, long long n) {
    const long long i = (long long)blockIdx.x * MBT_BLOCK + threadIdx.x;
    if (i >= n) return;
    const T *row = in + i * p.obs_dim;
    st.cash[i] = row[0];
    st.inv[i] = row[1];
    st.mid[i] = row[3];
    int d = 4;
    if (p.mid == MBT_MID_HESTON) { st.var[i] = row[d]; d += 1; }
    if (p.arr == MBT_ARR_HAWKES) { st.x0[i] = row[d]; st.x1[i] = row[d + 1]; d += 2; }
    if (imp_has_state(p.imp)) st.x0[i] = row[d];
}

/* ------------------------------------------------------------------ inventory histogram */
/*
 * Histogram of the inventory column over integer bins lo .. hi (the reference plots np.histogram of terminal inventories,
 * gym/helpers/plotting.py:94-110; north_star: "inventory distribution").  counts[0] = below lo, counts[1 + (q - lo)] for
 * lo <= round(q) <= hi, counts[hi - lo + 2] = above hi or NaN.  Shared-memory histogram per block (bins <= 4096), one global
 * 64-bit atomicAdd per non-empty bin and block.  Integer counts: the result does not depend on the order of the atomics.
 */
constexpr int MBT_HIST_MAX_BINS = 4096;
template <typename T>
__global__ void __launch_bounds__(MBT_BLOCK) mbt_inventory_hist_kernel(const T *__restrict__ inv, long long n, long long lo, int bins,
                                                                       unsigned long long *counts) {
    extern __shared__ unsigned int sh[];
    for (int b = threadIdx.x; b < bins + 2; b += MBT_BLOCK) sh[b] = 0u;
    __syncthreads();
    const long long stride = (long long)gridDim.x * MBT_BLOCK;
    for (long long i = (long long)blockIdx.x * MBT_BLOCK + threadIdx.x; i < n; i += stride) {
        const double q = (double)inv[i];
        int slot = bins + 1; /* NaN / above */
        if (q == q) {
            const double r = rint(q) - (double)lo;
            slot = r < 0.0 ? 0 : (r < (double)bins ? 1 + (int)r : bins + 1);
        }
        atomicAdd(&sh[slot], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < bins + 2; b += MBT_BLOCK)
        if (sh[b]) atomicAdd(counts + b, (unsigned long long)sh[b]);
}

/* ------------------------------------------------------------------ reward on caller rows */
template <typename T>
__global__ void __launch_bounds__(MBT_BLOCK) mbt_reward_kernel(StepParams<T> p, int is_terminal, const T *cur, const T *act, const T *nxt,
                                                               T *out, long long n) {
    const long long i = (long long)blockIdx.x * MBT_BLOCK + threadIdx.x;
    if (i >= n) return;
    const T *c = cur + i * p.obs_dim, *x = nxt + i * p.obs_dim;
    T a[MBT_MAX_ACTION_DIM] = {0, 0, 0, 0};
    for (int j = 0; j < p.action_dim; ++j) a[j] = act[i * p.action_dim + j];
    Traj<T> s;
    s.cash = x[0]; s.inv = x[1]; s.mid = x[3]; s.x0 = 0; s.x1 = 0; s.var = 0;
    StepClock<T> ck;
    ck.t_next = x[2];
    ck.dt_r = x[2] - c[2];
    ck.done = is_terminal;
    clock_derive<T>(ck, p.phi, p.alpha, p.ep_len);
    out[i] = reward_one<T, VariantGeneric>(p, ck, c[0], c[1], c[3], s, a, p.q0_uniform);
}

/* ------------------------------------------------------------------ fused rollout */
constexpr int MBT_SUMMARY_DOUBLES = 7; /* sum R, sum R^2, sum q, sum q^2, sum action, sum r^2, clip events */

template <typename T>
struct RolloutClock {
    StepClock<T> ck;
    T t_cur;     /* time column BEFORE the step: what the policy sees in its observation */
    T t_cur_obs; /* ... as a recorded observation shows it (normalised when normalise_obs) */
};

template <typename T>
struct RolloutArgs {
    StepParams<T> p;
    DevState<T> st;
    long long n;
    mbt_philox_keys keys;
    unsigned long long traj_offset, n_step0;
    const unsigned long long *counter_base; /* see StepArgs */
    int steps;            /* env-steps to run (until the episode ends) */
    /* device, `steps` entries: the uniform clock of every step, formed on the host exactly like the step kernel's
     * (mbt_make_clock on the host-accumulated t += dt) -- the loop loads it instead of redoing the arithmetic per thread */
    const RolloutClock<T> *clocks;
    /* policy */
    int pol_kind, table_rows, table_cols, inv_offset;
    T fixed[MBT_MAX_ACTION_DIM];
    T as_gamma, as_sigma_sq, as_fill_comp, as_terminal_time;
    const T *table; /* device */
    /* outputs */
    T *returns; /* (N,) or NULL */
    T *term_q;  /* (N,) or NULL */
    /* optional full recording, TIME-MAJOR so every store is a coalesced row write:
     * rec_obs (steps+1, N, D) normalised like reset()/step() return them, rec_act (steps, N, A) as the agent
     * returned them, rec_rew (steps, N).  The host exposes them transposed, in the shapes of generate_trajectory. */
    T *rec_obs, *rec_act, *rec_rew;
    double *block_sums; /* (gridDim.x, MBT_SUMMARY_DOUBLES) */
    double *summary;    /* (MBT_SUMMARY_DOUBLES,): the block rows folded in a fixed order by the last block to finish */
    unsigned int *ticket; /* zero on entry, zero again on exit */
    unsigned long long *clipped;
};

/* POL >= 0: the policy kind is known at compile time (the action row stays in registers, no branches per step);
 * POL < 0: runtime g.pol_kind (the recording kernels) */
template <typename T, int POL>
__device__ __forceinline__ void policy_action(const RolloutArgs<T> &g, int A, int k, T t, const Traj<T> &s, T *a) {
    const int kind = POL >= 0 ? POL : g.pol_kind;
    if (kind == MBT_POL_AVELLANEDA_STOIKOV) { /* BaselineAgents.py:62-83 */
        const T tau = g.as_terminal_time - t;
        const T adj = ((s.inv * g.as_gamma) * g.as_sigma_sq) * tau;
        const T spread = (g.as_gamma == (T)0) ? g.as_fill_comp : (g.as_gamma * g.as_sigma_sq) * tau + g.as_fill_comp;
        a[0] = adj + spread / (T)2;
        a[1] = -adj + spread / (T)2;
    } else if (kind == MBT_POL_CJ_MM_TABLE) { /* BaselineAgents.py:116-137 */
        T fi = (T)g.inv_offset + s.inv;
        const T hi = (T)(2 * g.inv_offset);
        fi = fi < (T)0 ? (T)0 : fi;
        fi = fi > hi ? hi : fi;
        const int idx = (int)fi; /* .astype(int) truncates */
        const T *e = g.table + ((long long)k * g.table_cols + idx) * 2;
        a[0] = e[0];
        a[1] = e[1];
    } else if (kind == MBT_POL_SCHEDULE) {
#pragma unroll
        for (int j = 0; j < MBT_MAX_ACTION_DIM; ++j)
            if (j < A) a[j] = g.table[(long long)k * A + j];
    } else { /* MBT_POL_FIXED  BaselineAgents.py:25-42 */
#pragma unroll
        for (int j = 0; j < MBT_MAX_ACTION_DIM; ++j) a[j] = g.fixed[j];
    }
}

/* REC: compile the trajectory-recording stores in (mbt_rollout_record) or out (mbt_rollout, the fast path);
 * POL: policy kind fixed at compile time (fast path) or -1 = runtime */
template <typename T, class V, bool REC, int POL>
__device__ __forceinline__ void mbt_rollout_body(const RolloutArgs<T> &g) {
    const long long i = (long long)blockIdx.x * MBT_BLOCK + threadIdx.x;
    const bool live = i < g.n;
    const StepParams<T> &p = g.p;
    double acc[MBT_SUMMARY_DOUBLES] = {0, 0, 0, 0, 0, 0, 0};
    if (live) {
        Traj<T> s;
        load_traj<T, V>(p, g.st, i, s);
        const int rew_kind = pick<V::rew>(p.rew);
        T q_init = p.q0_uniform;
        if ((rew_kind == MBT_REW_CJ_MM || rew_kind == MBT_REW_CJ_OE) && p.q0_per_traj) q_init = g.st.q0[i];
        const int A = action_width<T, V>(p);
        T ret = (T)0;
        int n_clipped = 0; /* clip EVENTS (steps in which inventory or cash hit a bound), like a loop of mbt_step counts them */
        const int D = obs_width<T, V>(p);
        const unsigned long long n_step0 = g.n_step0 + (g.counter_base ? g.counter_base[0] : 0ull);
        if (REC && g.rec_obs) {
            T row[MBT_MAX_OBS_DIM];
            make_obs_row<T, V>(p, s, g.clocks[0].t_cur_obs, row);
            store_row<T>(g.rec_obs, i, D, row, false);
        }
        for (int k = 0; k < g.steps; ++k) {
            const StepClock<T> ck = g.clocks[k].ck; /* warp-uniform loads (one transaction), L1-resident */
            T a[MBT_MAX_ACTION_DIM] = {0, 0, 0, 0};
            policy_action<T, POL>(g, A, k, g.clocks[k].t_cur, s, a);
            if (REC && g.rec_act) store_row<T>(g.rec_act, (long long)k * g.n + i, A, a, false);
#pragma unroll
            for (int j = 0; j < MBT_MAX_ACTION_DIM; ++j)
                if (j < A) {
                    acc[4] += (double)a[j];
                    a[j] = denorm_action<T, V>(p, a[j], j);
                }
            const mbt_u32x4 r = mbt_draw_keyed(g.keys, g.traj_offset + (unsigned long long)i, n_step0 + (unsigned long long)k, MBT_STREAM_STEP);
            /* batch-reduced fill models: only policies whose action is uniform over the batch reach here (do_rollout),
             * so the deepest quote of the batch is this trajectory's own */
            T fill_thr[2] = {(T)0, (T)0};
            if (fill_is_batch(pick<V::fill>(p.fill))) fill_batch_thresholds<T>(p, a[0], a[1], fill_thr);
            const uint32_t nbits2 = second_normal_bits<T, V>(p, g.keys, g.traj_offset + (unsigned long long)i, n_step0 + (unsigned long long)k);
            int clipped = 0;
            const T rwd = step_one<T, V, (POL != MBT_POL_FIXED)>(p, ck, s, a, r, q_init, &clipped, fill_thr, nbits2);
            n_clipped += clipped;
            ret = ret + rwd;
            acc[5] += (double)rwd * (double)rwd;
            if (REC && g.rec_rew) g.rec_rew[(long long)k * g.n + i] = rwd;
            if (REC && g.rec_obs) {
                T row[MBT_MAX_OBS_DIM];
                make_obs_row<T, V>(p, s, ck.t_obs, row);
                store_row<T>(g.rec_obs, (long long)(k + 1) * g.n + i, D, row, false);
            }
        }
        store_traj<T, V>(p, g.st, i, s);
        if (g.returns) g.returns[i] = ret;
        if (g.term_q) g.term_q[i] = s.inv;
        acc[0] = (double)ret;
        acc[1] = (double)ret * (double)ret;
        acc[2] = (double)s.inv;
        acc[3] = (double)s.inv * (double)s.inv;
        acc[6] = (double)n_clipped;
    }
    /* episode-return summaries: warp shuffle -> shared -> one row per block; the LAST block to finish (ticket) folds the
     * rows in a fixed order into g.summary, so the result is deterministic and stays on the device (a group all-reduces
     * it over NCCL from there).  This is the only cross-thread step of the path. */
    __shared__ double sm[MBT_BLOCK / 32][MBT_SUMMARY_DOUBLES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int m = 0; m < MBT_SUMMARY_DOUBLES; ++m) {
        double v = acc[m];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) sm[warp][m] = v;
    }
    __syncthreads();
    if (threadIdx.x < MBT_SUMMARY_DOUBLES) {
        double v = 0;
        for (int w = 0; w < MBT_BLOCK / 32; ++w) v += sm[w][threadIdx.x];
        g.block_sums[(long long)blockIdx.x * MBT_SUMMARY_DOUBLES + threadIdx.x] = v;
        if (threadIdx.x == 6 && v != 0.0) atomicAdd(g.clipped, (unsigned long long)v); /* cumulative counter of the handle */
        __threadfence(); /* this block's row is visible before its ticket is taken */
    }
    __shared__ bool is_last;
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(g.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    /* thread t folds rows t, t + 256, ... in increasing order; then the same shuffle -> shared tree as above */
    double part[MBT_SUMMARY_DOUBLES] = {0, 0, 0, 0, 0, 0, 0};
    for (unsigned b = threadIdx.x; b < gridDim.x; b += MBT_BLOCK) {
#pragma unroll
        for (int m = 0; m < MBT_SUMMARY_DOUBLES; ++m) part[m] += __ldcg(g.block_sums + (long long)b * MBT_SUMMARY_DOUBLES + m);
    }
#pragma unroll
    for (int m = 0; m < MBT_SUMMARY_DOUBLES; ++m) {
        double v = part[m];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) sm[warp][m] = v;
    }
    __syncthreads();
    if (threadIdx.x < MBT_SUMMARY_DOUBLES) {
        double v = 0;
        for (int w = 0; w < MBT_BLOCK / 32; ++w) v += sm[w][threadIdx.x];
        g.summary[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) *g.ticket = 0u;
}

template <typename T, class V, bool REC, int POL>
__global__ void __launch_bounds__(MBT_BLOCK) mbt_rollout_kernel(const __grid_constant__ RolloutArgs<T> g) {
    mbt_rollout_body<T, V, REC, POL>(g);
}

#endif /* MBT_KERNELS_CUH */
