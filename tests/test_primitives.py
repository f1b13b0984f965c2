"""RNG / math primitives shared by the kernels and the oracle (include/mbt_philox.h, include/mbt_math.h)."""
import numpy as np
from scipy import special

from oracle import oracle as O


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        assert [int(x) for x in O.philox(ctr, key)] == want


def _ulp_err(got, want, eps):
    return np.max(np.abs(got.astype(np.float64) - want) / np.abs(want)) / eps


def test_exp_log_accuracy_vs_libm():
    rng = np.random.default_rng(0)
    x = rng.uniform(-87, 88, 400_000).astype(np.float32)
    assert _ulp_err(O.vec("exp_f32", x), np.exp(x.astype(np.float64)), 2.0 ** -24) < 2.0
    x = np.exp(rng.uniform(-80, 80, 400_000)).astype(np.float32)
    assert _ulp_err(O.vec("log_f32", x), np.log(x.astype(np.float64)), 2.0 ** -24) < 2.0
    x = rng.uniform(-700, 700, 400_000)
    assert _ulp_err(O.vec("exp_f64", x), np.exp(x), 2.0 ** -53) < 4.0
    x = np.exp(rng.uniform(-700, 700, 400_000))
    assert _ulp_err(O.vec("log_f64", x), np.log(x), 2.0 ** -53) < 4.0
    # range ends: branch-free clamp to [-87, 88] / [-700, 700]; NaN counts as the lower end
    ends = O.vec("exp_f32", np.array([-100.0, -87.0, 88.0, 100.0, np.nan], np.float32))
    assert ends[0] == ends[1] == ends[4] and ends[2] == ends[3] and 1e-38 < ends[0] < 2e-38 and ends[2] > 1e38
    ends = O.vec("exp_f64", np.array([-1000.0, -700.0, 700.0, 1000.0]))
    assert ends[0] == ends[1] and ends[2] == ends[3] and ends[0] < 1e-300 and ends[2] > 1e300


def test_pow_matches_numpy_fast_paths():
    x = np.array([-3.0, -1.0, 0.0, 0.5, 2.0, 17.0])
    assert np.array_equal(O.vec("pow_f64", x, 2.0), x ** 2.0)
    assert np.array_equal(O.vec("pow_f64", x, 1.0), x ** 1.0)
    np.testing.assert_allclose(O.vec("pow_f64", x, 4.0), x ** 4.0, rtol=1e-13)
    np.testing.assert_allclose(O.vec("pow_f64", x, 3.0), x ** 3.0, rtol=1e-13)


def test_normal_from_bits_is_the_normal_quantile():
    rng = np.random.default_rng(1)
    bits = rng.integers(0, 2 ** 32, 1_000_000, dtype=np.uint64).astype(np.uint32)
    ks = np.arange(1, 32.01, 0.01)
    mc = np.clip(np.round(2.0 ** -ks * 2 ** 31 - 0.5), 0, 2 ** 31 - 1).astype(np.uint64)
    bits = np.concatenate([bits, (0x7FFFFFFF - mc).astype(np.uint32), np.array([0, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF], np.uint32)])
    t = ((0x7FFFFFFF - (bits & 0x7FFFFFFF)).astype(np.float64) + 0.5) / 2 ** 31
    want = np.where(bits >> 31, 1.0, -1.0) * special.ndtri(t / 2)
    z64 = O.vec("normal_f64", bits)
    z32 = O.vec("normal_f32", bits).astype(np.float64)
    assert np.max(np.abs(z64 - want) / np.maximum(np.abs(want), 1e-9)) < 1e-13
    assert np.max(np.abs(z32 - want)) < 3e-6
    assert np.max(np.abs(z64)) < 6.34 and np.all(np.isfinite(z32))
    main = z64[:1_000_000]
    assert abs(main.mean()) < 5e-3 and abs(main.var() - 1) < 5e-3 and abs((main ** 4).mean() - 3) < 0.05


def test_draws_are_uniform_and_independent_across_trajectories_and_steps():
    from mbt_gym_b200 import _abi
    u0, z0 = O.draws(_abi.MBT_F64, 123, 0, 50_000, 0)
    u1, z1 = O.draws(_abi.MBT_F64, 123, 0, 50_000, 1)
    assert u0.min() >= 0 and u0.max() < 1
    assert np.all(np.abs(u0.mean(axis=0) - 0.5) < 0.01)
    assert abs(np.corrcoef(z0, z1)[0, 1]) < 0.02 and abs(np.corrcoef(u0[:, 0], u0[:, 2])[0, 1]) < 0.02
    # global trajectory ids: a shard at offset 1000 sees the same numbers as rows 1000.. of the full batch
    u_part, z_part = O.draws(_abi.MBT_F64, 123, 1000, 100, 0)
    assert np.array_equal(u_part, u0[1000:1100]) and np.array_equal(z_part, z0[1000:1100])
    # float32 uniforms are the same grid points
    u32, _ = O.draws(_abi.MBT_F32, 123, 0, 1000, 0)
    assert np.array_equal(u32.astype(np.float64), u0[:1000])


def _hard_division_cases(p, n, rng):
    """Pairs (a, b) of p-bit significands whose quotient lies within s * 2^-(2p) (relative) of a MIDPOINT between two
    floating-point numbers -- the inputs on which a division that is not correctly rounded goes wrong.  With Qm odd
    (p+1 bits, the midpoint scaled by 2^(p+1)) pick B = s * Qm^-1 mod 2^(p+1); then B * Qm = s (mod 2^(p+1)) and
    A = (B * Qm - s) / 2^(p+1) is an integer with A / B = Qm / 2^(p+1) - s / (B * 2^(p+1))."""
    mod = 1 << (p + 1)
    out_a, out_b = [], []
    while len(out_a) < n:
        qm = int(rng.integers(1 << p, mod)) | 1
        s = int(rng.integers(1, 64)) * (1 if rng.random() < 0.5 else -1)
        b = (s * pow(qm, -1, mod)) % mod
        if not (1 << (p - 1)) <= b < (1 << p):
            continue
        a = (b * qm - s) >> (p + 1)
        if (a << (p + 1)) != b * qm - s or not (1 << (p - 1)) <= a < (1 << p):
            continue
        out_a.append(a); out_b.append(b)
    return np.array(out_a, dtype=np.float64), np.array(out_b, dtype=np.float64)


def test_division_through_the_hosts_reciprocal_equals_ieee_division_bitwise():
    """mbt_div_rcp_* (include/mbt_math.h): what the kernels use for (obs - low) / grad -- bit-identical to `/` on random
    pairs over the whole exponent range, on the special values, and on quotients adversarially close to rounding midpoints."""
    rng = np.random.default_rng(2024)
    for dt, p, emax in ((np.float64, 53, 600), (np.float32, 24, 70)):
        n = 20_000_000 if dt == np.float64 else 10_000_000
        a = (rng.standard_normal(n) * np.exp2(rng.uniform(-emax, emax, n))).astype(dt)
        b = (rng.standard_normal(n) * np.exp2(rng.uniform(-emax, emax, n))).astype(dt)
        assert O.div_rcp_mismatches(a, b) == 0
        # realistic divisors (observation-space half-widths), numerators of any size incl. zeros, infinities, NaN
        b = rng.choice(np.array([21600.0, 200.0, 0.5, 8.0, 1e4, 3.0, 100.0 / 3, 1e-3], dtype=dt), n)
        a[: n // 50] = rng.choice(np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-300, 1e300, 5e-324], dtype=dt), n // 50)
        assert O.div_rcp_mismatches(a, b) == 0
        ha, hb = _hard_division_cases(p, 60_000, rng)
        scale = np.exp2(rng.integers(-40, 40, ha.size)).astype(np.float64)
        assert O.div_rcp_mismatches((ha * scale).astype(dt), hb.astype(dt)) == 0
        assert O.div_rcp_mismatches((-ha).astype(dt), (hb * scale).astype(dt)) == 0


def test_float_filtered_fill_decision_equals_its_float64_definition():
    """mbt_u24_below_exp_f64 decides `k < exp(x) * 2^24` with the float exponential unless k falls inside a band around the
    float estimate: same decisions as the float64 definition for uniform draws AND for k placed right at the threshold,
    and the float estimate's error stays an order of magnitude inside the band (1.6e-5)."""
    rng = np.random.default_rng(77)
    n = 20_000_000
    x = rng.uniform(-18.0, 1.0, n)
    x[: n // 10] = -1.5 * rng.uniform(0.0, 3.2, n // 10)  # the quoting range of the reference's markets
    k = rng.integers(0, 1 << 24, n, dtype=np.uint32)
    bad, worst = O.fill_filter_check(k, x)
    assert bad == 0 and worst < 2.0e-6, (bad, worst)
    # adversarial: k within a few integers / a few 1e-5 (relative) of the threshold
    thr = np.exp(np.clip(x, -50, 1)) * 16777216.0
    near = thr * (1 + rng.choice([0, 1e-7, -1e-7, 1e-6, -1e-6, 1.5e-5, -1.5e-5, 1.7e-5, -1.7e-5, 3e-5, -3e-5], n)) + rng.integers(-2, 3, n)
    k2 = np.clip(np.rint(near), 0, (1 << 24) - 1).astype(np.uint32)
    bad, _ = O.fill_filter_check(k2, x)
    assert bad == 0
