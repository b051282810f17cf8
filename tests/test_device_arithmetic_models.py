"""CPU models of the device-side arithmetic shortcuts, checked exhaustively or adversarially against exact integers.

The CUDA kernels replace divisions by quotient estimates (float reciprocal with the 1.5*2^23 rounding trick in the plane
split, 32-bit Barrett variants in the CRT and panel kernels).  Each shortcut comes with a hand-derived error bound in the
source; these tests restate the device formulas operation by operation (same widths, same rounding) and check the bounds
by exhaustion / adversarial inputs, so a wrong bound cannot hide behind "the random GPU tests passed".
Sources: gpufinitefieldmatrices.jl_b200/csrc/gemm_tc.cu (Encoder::plane, crt_fast_kernel), csrc/pluq.cu (pa_reduce, ll_reduce).
"""
import numpy as np
import pytest

MODULI = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197]  # kModuli, gemm_tc.cu
MAGIC = np.float32(12582912.0)  # 1.5 * 2^23


def fma32(a, b, c):
    """fmaf on float32 arrays: the product of two float32 is exact in float64, one rounding at the end."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


@pytest.mark.parametrize("m", MODULI[1:])
def test_split_fp32_digit_is_exact_for_every_reachable_t(m):
    """Encoder<2>::plane: t = hi*(2^14 mod m) + lo with |hi| <= 2^13, 0 <= lo < 2^14, so |t| <= 8192*254 + 16383.
    For EVERY integer t in that range the digit must be congruent to t and lie in [-127, 127]."""
    tmax = 8192 * (m - 1) + 16383
    t = np.arange(-tmax, tmax + 1, dtype=np.int64)
    tf = t.astype(np.float32)
    assert np.array_equal(tf.astype(np.int64), t)  # exact in fp32
    inv = np.float32(1.0 / m)
    q = fma32(tf, np.full_like(tf, inv), np.full_like(tf, MAGIC)) - MAGIC
    r = fma32(q, np.full_like(tf, np.float32(-m)), tf)
    rb = (r + MAGIC).view(np.uint32) & 0xFF  # low mantissa byte, as the kernel extracts it
    digit = rb.astype(np.int64)
    digit[digit >= 128] -= 256
    assert np.array_equal(digit, r.astype(np.int64))
    assert int(np.abs(digit).max()) <= 127
    assert np.all((digit - t) % m == 0)


def test_split_prep_and_full_digit_on_samples():
    """prepare(): v' = v + 2^27, hi = (v' >> 14) - 2^13, lo = v' & 16383 for |v| <= 2^27; digit == v (mod m)."""
    rng = np.random.default_rng(1)
    v = np.concatenate([rng.integers(-2 ** 27, 2 ** 27 + 1, size=200000), [-2 ** 27, 2 ** 27, 0, -1, 1, 2 ** 26, -2 ** 26]]).astype(np.int64)
    vp = v + 2 ** 27
    hi = (vp >> 14) - 8192
    lo = vp & 16383
    assert np.array_equal(hi * 16384 + lo, v)
    for m in MODULI[1:]:
        t = hi * (16384 % m) + lo
        assert np.all((t - v) % m == 0) and np.abs(t).max() < 2 ** 22
    assert np.all(((vp & 255) - v) % 256 == 0)  # m = 256: low byte


def _crt_setup(s, P):
    mods = MODULI[:s]
    M = 1
    for m in mods:
        M *= m
    w = [(M // m) % P for m in mods]
    u = [pow((M // m) % m, -1, m) for m in mods]
    f23 = [(1 << 23) // m for m in mods]
    Wq = [(P - (q * (M % P)) % P) % P for q in range(s + 2)]
    return mods, M, w, u, f23, Wq


@pytest.mark.parametrize("s,P,balanced", [(8, 33554393, True), (8, 33554393, False), (5, 65537, True), (15, 4294967291, True), (3, 2 ** 29 - 3, False)])
def test_crt_fast_quotient_estimate_and_reduction(s, P, balanced):
    """crt_fast_kernel: q = (sum e_t*floor(2^23/m_t) + rnd) >> 23 must equal round(S/M) (balanced) / floor(S/M) for every x the
    plan admits (|x|/M < 1/2 - 2^-7 resp. x/M < 1 - 2^-6), and the single-multiply reduction must return x mod P."""
    mods, M, w, u, f23, Wq = _crt_setup(s, P)
    lim = M // 2 - M // 128 if balanced else M - M // 64
    xs = [0, 1, lim - 1, lim // 2, lim // 3]
    if balanced:
        xs += [-1, -(lim - 1), -(lim // 2)]
    import random
    pr = random.Random(s)
    xs += [pr.randrange(-lim + 1 if balanced else 0, lim) for _ in range(3000)]
    # values right at the decision boundary of the quotient (x/M close to +-1/2 resp. 1): the plan's margin must cover them
    xs += [lim - 1 - pr.randrange(0, 1000) for _ in range(200)] + ([-(lim - 1) + pr.randrange(0, 1000) for _ in range(200)] if balanced else [])
    rnd = (1 << 22) + (1 << 11) if balanced else (1 << 12)
    fast = (1 << 16) < P < (1 << 30)
    mu48 = (1 << 48) // P
    for x in xs:
        e = [((x % m) * u[t]) % m for t, m in enumerate(mods)]  # what the GEMM epilogue stores
        S = sum(et * (M // m) for et, m in zip(e, mods))
        q_true = (S - x) // M
        assert S - q_true * M == x
        F = sum(et * f for et, f in zip(e, f23))
        assert F < 2 ** 32
        q = (F + rnd) >> 23
        assert q == q_true, (x, q, q_true)
        acc = sum(et * wt for et, wt in zip(e, w)) + Wq[q]
        assert acc < 2 ** 46 or not fast
        if fast:
            xh = acc >> 16
            assert xh < 2 ** 32
            r = (acc - ((xh * mu48) >> 32) * P)
            assert 0 <= r < 3 * P and 3 * P < 2 ** 32
            r %= 2 ** 32
            for _ in range(2):
                r = min(r, (r - P) % 2 ** 32)
        else:
            r = acc % P
        assert r == x % P


def _bits(P):
    return P.bit_length()


@pytest.mark.parametrize("P", [2, 3, 7, 11, 251, 256, 257, 65521, 65535, 65537, 131071, 33554393, 67108859, 2 ** 26 - 1, 2 ** 25, 536870909])
def test_panel_quotient_estimates(P):
    """pa_reduce<1/4> (x < P^2 + 2.6 P, remainder < 2.6 P after the estimate) and ll_reduce (x < 32 P^2, three conditional
    subtractions, only for bits(P) <= 26); pa_reduce<0/3> (x < 2^32, remainder < 2P) for P < 2^16."""
    rng = np.random.default_rng(P % 1000)
    b = _bits(P)
    sh = b - 1
    mu = min((1 << (32 + sh)) // P, 2 ** 32 - 1)

    def adversarial(limit):
        xs = [0, 1, P - 1, P, P + 1, limit - 1, limit - P, limit // 2]
        xs += [k * P - 1 for k in (1, 2, 3, limit // P)] + [k * P for k in (1, 2, limit // P - 1)]
        xs += [int(v) % limit for v in rng.integers(0, 2 ** 62, size=4000)]
        return [x for x in xs if 0 <= x < limit]

    if (1 << 16) < P < (1 << 30):  # mid path: lazy inputs up to 2.6 P
        limit = P * P + (26 * P) // 10
        assert limit >> sh < 2 ** 32
        for x in adversarial(limit):
            r = x - (((x >> sh) * mu) >> 32) * P
            assert 0 <= r < (26 * P) // 10 + 1, (x, r)
    if b <= 26:  # left-looking dot products: up to 31 products + slack
        limit = 32 * P * P
        for x in adversarial(limit):
            xh = x >> sh
            assert xh < 2 ** 32
            r = x - ((xh * mu) >> 32) * P
            assert 0 <= r < 4 * P and 4 * P <= 2 ** 32
            for _ in range(3):
                r = min(r, (r - P) % 2 ** 32)
            assert r == x % P
    if P < (1 << 16):  # small path, lazy: x = nl*u + a <= (P-1)^2 + 2P - 1 < 2^32
        mu32 = (1 << 32) // P
        limit = (P - 1) * (P - 1) + 2 * P
        assert limit <= 2 ** 32 or P == 65535
        for x in adversarial(min(limit, 2 ** 32)):
            r = x - ((x * mu32) >> 32) * P
            assert 0 <= r < 2 * P


# ---------------------------------------------------------------- round 2: GEMV accumulation budget, 64-bit Barrett, wide products ----
def _mod_u64(v, P):
    """common.cuh mod_u64: q = mulhi64(v, floor((2^64-1)/P)), r = v - q*P (mod 2^64), two conditional subtractions"""
    mu = (2 ** 64 - 1) // P
    q = (v * mu) >> 64
    r = (v - q * P) % 2 ** 64
    if r >= P:
        r -= P
    if r >= P:
        r -= P
    return r


@pytest.mark.parametrize("P", [1, 2, 3, 11, 65521, 33554393, 2 ** 32 - 5, 2 ** 32 - 1, 8191 * 8191, 2 ** 52 - 1, 2 ** 52, (1 << 26) ** 2])
def test_barrett_u64_two_corrections_suffice(P):
    """mod_u64 is used with ARBITRARY 64-bit inputs (raw uint64 accumulators in gemv.cu, Karatsuba sums): the quotient estimate is at
    most 2 short for every v < 2^64 and every P the library passes (P <= 2^52), so two conditional subtractions finish the reduction."""
    rng = np.random.default_rng(P % 1000)
    vs = [0, 1, P - 1, P, P + 1, 2 * P - 1, 2 * P, 2 ** 64 - 1, 2 ** 64 - P, 2 ** 63, 2 ** 63 - 1, (2 ** 64 - 1) // P * P, (2 ** 64 - 1) // P * P - 1]
    vs += [int(x) for x in rng.integers(0, 2 ** 63, size=4000, dtype=np.int64)] + [int(x) + 2 ** 63 for x in rng.integers(0, 2 ** 63, size=4000, dtype=np.int64)]
    for v in vs:
        if 0 <= v < 2 ** 64:
            assert _mod_u64(v, P) == v % P, (v, P)


def _terms_budget(bound2):
    """gemv.cu terms_budget: floor(9.0e18 / bound2), capped at 2^20"""
    if bound2 < 1:
        return 1 << 20
    return min(1 << 20, int(9.0e18 / bound2))


@pytest.mark.parametrize("R,P", [(11, 11), (65521, 65521), (33554393, 33554393), (2 ** 32 - 5, 65521), (2 ** 32, 2 ** 32 - 5), (2 ** 31, 3), (4294967291, 4294967291),
                                 (2 ** 29 + 1, 2 ** 29 + 1), (1518500250, 7), (3037000500, 7)])
def test_gemv_accumulation_budget_never_wraps(R, P):
    """gemv_kernel<0>: raw products (< (R-1)^2) are added to a uint64 accumulator, reduced whenever the NEXT batch of GV_UNROLL terms
    could exceed the budget T, the reduced value counting as one term.  Worst case (every product maximal): the accumulator must stay
    below 2^64 at every step; when the budget cannot hold two batches the kernel switches to MODE 1 (every product reduced)."""
    GV_UNROLL = 8
    b2 = (R - 1) ** 2
    budget = _terms_budget(float(max(b2, P)))
    mode = 1 if budget < 2 * GV_UNROLL else 0
    T = (1 << 20) if mode else budget
    term = (P - 1) if mode else b2  # MODE 1 adds mod_u64(product) < P
    acc, since, peak = 0, 0, 0
    for _ in range(40000 // GV_UNROLL):
        acc += GV_UNROLL * term
        peak = max(peak, acc)
        since += GV_UNROLL
        if since + GV_UNROLL > T:
            acc = P - 1  # worst reduced value
            since = 1
    assert peak < 2 ** 64, (R, P, mode, T, peak.bit_length())
    if mode == 1:
        assert (1 << 20) * (P - 1) < 2 ** 64
    # the cross-warp and K-slice sums add at most 8 + 64 reduced values
    assert 72 * (P - 1) < 2 ** 64


@pytest.mark.parametrize("P", [2 ** 32 + 15, 2 ** 45 + 59, 2 ** 52 - 47, 2 ** 52, 8191 * 8191 * 8191])
def test_wide_mulmod_through_byte_steps(P):
    """wide.cu mulmod_u64 (2^32 < P <= 2^52): hi = (a*b) >> 64 < 2^40, acc = hi % P, then eight steps acc = ((acc << 8) | byte) % P over
    the low word -- every intermediate stays below 2^60 and the result is (a*b) mod P."""
    rng = np.random.default_rng(7)
    cases = [(P - 1, P - 1), (P - 1, 1), (0, P - 1), (2 ** 32, 2 ** 32 % P), (P // 2, P // 2 + 1)]
    cases += [(int(a) % P, int(b) % P) for a, b in rng.integers(0, 2 ** 62, size=(3000, 2), dtype=np.int64)]
    for a, b in cases:
        prod = a * b
        hi, lo = prod >> 64, prod % 2 ** 64
        assert hi < 2 ** 40
        acc = hi % P
        for k in range(0, 64, 8):
            sh = (acc << 8) | ((lo >> (56 - k)) & 0xFF)
            assert sh < 2 ** 60
            acc = sh % P
        assert acc == prod % P
