/*
 * mbt_philox.h -- Philox4x32-10 counter-based RNG, shared by the sm_100a kernels
 * (mbt_gym_b200/csrc) and by the CPU oracle (oracle/mbt_oracle.c).  Plain C99 / CUDA C++.
 *
 * Why it exists: the reference draws from one numpy PCG64 `Generator` per stochastic
 * process (mbt_gym/stochastic_processes/StochasticProcessModel.py:27,37-39; consumers at
 * arrival_models.py:55,122, fill_probability_models.py:33, midprice_models.py:64,143).
 * A sequential generator cannot be evaluated per (trajectory, step) on a GPU, so the
 * B200 path replaces it with the counter-based Philox4x32-10 of Salmon et al.,
 * "Parallel random numbers: as easy as 1, 2, 3" (SC'11): 10 rounds of
 *     (c0,c1,c2,c3) <- (hi(M1*c2)^c1^k0, lo(M1*c2), hi(M0*c0)^c3^k1, lo(M0*c0))
 * with the key bumped by the Weyl constants each round.  Known-answer vectors from the
 * Random123 distribution are checked in tests/test_primitives.py.
 *
 * Draw contract (one call per trajectory per env-step):
 *     key     = (seed_lo, seed_hi)
 *     counter = (traj_lo, traj_hi, n_lo, (stream << 24) | (n_hi & 0xFFFFFF))
 * where traj is the GLOBAL trajectory id (so results do not depend on how trajectories
 * are sharded over GPUs), n is the draw index (env-step counter since seed(), or the
 * episode counter for stream MBT_STREAM_RESET) and stream separates uses.
 */
#ifndef MBT_PHILOX_H
#define MBT_PHILOX_H

#include "mbt_rtc.h" /* <stdint.h> or its NVRTC stand-in */

#if defined(__CUDACC__)
#define MBT_HD __host__ __device__ __forceinline__
#else
#define MBT_HD static inline
#endif

#define MBT_PHILOX_M0 0xD2511F53u
#define MBT_PHILOX_M1 0xCD9E8D57u
#define MBT_PHILOX_W0 0x9E3779B9u
#define MBT_PHILOX_W1 0xBB67AE85u

#define MBT_STREAM_STEP 0u  /* per-step draws: arrivals, fills, midprice normal */
#define MBT_STREAM_RESET 1u /* per-episode draws: random initial inventory      */
#define MBT_STREAM_STEP2 2u /* per-step draws of models that need a second normal (Heston variance) */

typedef struct mbt_u32x4 {
    uint32_t x, y, z, w;
} mbt_u32x4;

MBT_HD void mbt_mulhilo32(uint32_t a, uint32_t b, uint32_t *hi, uint32_t *lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b; /* one IMAD.WIDE.U32 on the device */
    *lo = (uint32_t)p;
    *hi = (uint32_t)(p >> 32);
}

/* The ten round keys of a seed (key bumped by the Weyl constants each round).  The step kernels get them
 * precomputed from the host in their argument block (constant bank) instead of re-deriving them per thread. */
typedef struct mbt_philox_keys {
    uint32_t k0[10], k1[10];
} mbt_philox_keys;

MBT_HD mbt_philox_keys mbt_philox_expand(uint64_t seed) {
    mbt_philox_keys K;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        K.k0[r] = k0;
        K.k1[r] = k1;
        k0 += MBT_PHILOX_W0;
        k1 += MBT_PHILOX_W1;
    }
    return K;
}

#if defined(__cplusplus)
MBT_HD mbt_u32x4 mbt_philox4x32_10_keyed(mbt_u32x4 c, const mbt_philox_keys &K) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mbt_mulhilo32(MBT_PHILOX_M0, c.x, &hi0, &lo0);
        mbt_mulhilo32(MBT_PHILOX_M1, c.z, &hi1, &lo1);
        mbt_u32x4 n;
        n.x = hi1 ^ c.y ^ K.k0[r];
        n.y = lo1;
        n.z = hi0 ^ c.w ^ K.k1[r];
        n.w = lo0;
        c = n;
    }
    return c;
}

/* mbt_draw with pre-expanded keys: identical output. */
MBT_HD mbt_u32x4 mbt_draw_keyed(const mbt_philox_keys &K, uint64_t traj, uint64_t n, uint32_t stream) {
    mbt_u32x4 c;
    c.x = (uint32_t)traj;
    c.y = (uint32_t)(traj >> 32);
    c.z = (uint32_t)n;
    c.w = (stream << 24) | ((uint32_t)(n >> 32) & 0x00FFFFFFu);
    return mbt_philox4x32_10_keyed(c, K);
}
#endif

MBT_HD mbt_u32x4 mbt_philox4x32_10(mbt_u32x4 c, uint32_t k0, uint32_t k1) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mbt_mulhilo32(MBT_PHILOX_M0, c.x, &hi0, &lo0);
        mbt_mulhilo32(MBT_PHILOX_M1, c.z, &hi1, &lo1);
        mbt_u32x4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += MBT_PHILOX_W0;
        k1 += MBT_PHILOX_W1;
    }
    return c;
}

/* The draw contract above, in one place. */
MBT_HD mbt_u32x4 mbt_draw(uint64_t seed, uint64_t traj, uint64_t n, uint32_t stream) {
    mbt_u32x4 c;
    c.x = (uint32_t)traj;
    c.y = (uint32_t)(traj >> 32);
    c.z = (uint32_t)n;
    c.w = (stream << 24) | ((uint32_t)(n >> 32) & 0x00FFFFFFu);
    return mbt_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}

/*
 * The five random quantities one env-step consumes, carved out of ONE 128-bit draw:
 *   top 24 bits of each word -> the four uniforms in [0,1) on the 2^-24 grid
 *       (x: bid arrival, y: ask arrival, z: bid fill, w: ask fill -- the consumption
 *        order of the reference, arrival_models.py:55 then fill_probability_models.py:33);
 *   low 8 bits of the four words, concatenated -> 32 bits for the midprice normal
 *       (midprice_models.py:64,143), mapped through mbt_normal_from_bits_f32 (mbt_math.h) in BOTH
 *       precisions: the float64 path widens that float (the draw is a definition, its 24-bit resolution
 *       is far below any statistical resolution; state arithmetic stays float64).
 * All 128 output bits of a Philox block are independent, so the five values are too.
 */
MBT_HD uint32_t mbt_uniform_bits24(uint32_t word) { return word >> 8; }

MBT_HD uint32_t mbt_normal_bits(mbt_u32x4 r) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(__byte_perm(r.x, r.y, 0x0040), __byte_perm(r.z, r.w, 0x0040), 0x5410); /* 3 PRMT */
#else
    return (r.x & 0xFFu) | ((r.y & 0xFFu) << 8) | ((r.z & 0xFFu) << 16) | ((r.w & 0xFFu) << 24);
#endif
}

#endif /* MBT_PHILOX_H */
