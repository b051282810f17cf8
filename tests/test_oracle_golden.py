"""Pins oracle/oracle.py against the reference's own known-answer fixtures
(tests/golden/reference_fixtures.json, transcribed from /root/reference/test/**) and against
exact python-int arithmetic.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

FX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_fixtures.json")))


def test_de_rham_inverse():
    f = FX["de_rham"]
    A = np.array(f["A"])
    ok, inv = O.is_invertible_with_inverse(A, f["N"])
    assert ok is True
    assert np.array_equal(O.exact_matmul_mod(np.mod(A, f["N"]), inv, f["N"]), np.array(f["A_times_inverse"]))


def test_matmul_literals():
    f = FX["matmul_2x3_3x2"]
    A, B = np.array(f["A"]), np.array(f["B"])
    assert np.array_equal(O.matmul_mod(A, B, f["N"]), np.array(f["C_mod11"]))
    assert np.array_equal(O.matmul_mod(A, B, f["override_N"]), np.array(f["C_mod7"]))
    assert np.array_equal(np.mod(np.array(f["C_literal"]), 11), np.array(f["C_mod11"]))
    g = FX["matmul_inplace"]
    A, B = np.array(g["A"]), np.array(g["B"])
    assert np.array_equal(O.matmul_mod(A, B, g["N"]), np.array(g["C_mod9"]))
    assert np.array_equal(O.matmul_mod(A, B, g["override_N"]), np.array(g["C_mod3"]))


def test_basic_3x3():
    f = FX["basic_3x3"]
    A, B, N, s = np.array(f["A"]), np.array(f["B"]), f["N"], f["scalar"]
    assert np.array_equal(O.ew_add(A, B, N), np.array(f["add"]))
    assert np.array_equal(O.ew_sub(A, B, N), np.array(f["sub"]))
    assert np.array_equal(O.matmul_mod(A, B, N), np.array(f["matmul"]))
    assert np.array_equal(O.ew_mul(A, B, N), np.array(f["elementwise_multiply"]))
    assert np.array_equal(O.ew_scalar_add(A, s, N), np.array(f["scalar_add"]))
    assert np.array_equal(O.ew_scalar_sub(A, s, N), np.array(f["scalar_sub"]))
    assert np.array_equal(O.ew_scalar_mul(A, s, N), np.array(f["scalar_mul"]))
    assert np.array_equal(O.ew_negate(A, N), np.array(f["negate"]))
    assert np.array_equal(O.matmul_mod(A, A, N), np.array(f["pow2"]))


def test_permutations():
    f = FX["permutation_3x3"]
    A = np.array(f["A"])
    P = [tuple(p) for p in f["P"]]
    Ac = O.apply_col_perm(P, A)
    assert np.array_equal(Ac, np.array(f["col_perm"]))
    assert np.array_equal(O.apply_col_perm(P, Ac, inverse=True), A)
    Ar = O.apply_row_perm(P, A)
    assert np.array_equal(Ar, np.array(f["row_perm"]))
    assert np.array_equal(O.apply_row_perm(P, Ar, inverse=True), A)


def test_triangular_2x2_and_sweep():
    f = FX["triangular_2x2"]
    N = f["N"]
    U = np.array(f["upper"]); L = np.array(f["lower"])
    assert np.array_equal(O.exact_matmul_mod(U, O.upper_triangular_inverse(U, N), N), np.eye(2, dtype=np.int64))
    assert np.array_equal(O.exact_matmul_mod(L, O.lower_triangular_inverse(L, N), N), np.eye(2, dtype=np.int64))
    rng = np.random.default_rng(0)
    # triangular_test.jl:82-88 sweep (subset of primes/sizes to keep CPU time low)
    for p in (3, 13, 97):
        for n in (33, 47, 65):
            T = np.tril(rng.integers(1, p, size=(n, n)))
            Ti = O.lower_triangular_inverse(T, p)
            assert np.array_equal(O.exact_matmul_mod(T, Ti, p), np.eye(n, dtype=np.int64))
            assert np.array_equal(O._fast_tri_inverse_upper(T.T, p).T, Ti)
    # wide upper (triangular_test.jl:43-45): A * A_inv == I_rows
    T = np.triu(rng.integers(1, 13, size=(20, 50)))
    Ti = O.upper_triangular_inverse(T, 13)
    assert Ti.shape == (50, 20)
    assert np.array_equal(O.exact_matmul_mod(T, Ti, 13), np.eye(20, dtype=np.int64))
    with pytest.raises(ValueError):
        O.lower_triangular_inverse(np.ones((5, 3), dtype=np.int64), 13)


def test_fill_and_ctor():
    f = FX["fill_122_mod_11"]
    assert f["value"] % f["N"] == f["expect"] == 1
    img = O.construct(np.array([[-3, 12], [5, 6]]), 7)
    assert img.shape == (34, 34)
    assert img[0, 0] == 4 and img[0, 1] == 5 and img[2:, :].sum() == 0 and img[:, 2:].sum() == 0
    with pytest.raises(ValueError):
        O.construct(np.array([[0.5]]), 7)
    with pytest.raises(ValueError):
        O.construct(np.array([[1]]), 2 ** 52 + 1)


@pytest.mark.parametrize("case", FX["stripe_cases"]["cases"][:2])
def test_stripe_cases_vs_exact(case):
    rng = np.random.default_rng(case["n"] + case["N"])
    n, N = case["n"], case["N"]
    A = rng.integers(case["lo"], case["hi"] + 1, size=(n, n))
    B = rng.integers(case["lo"], case["hi"] + 1, size=(n, n))
    for (i, j, v, which) in case.get("poke", []):
        (A if which == "A" else B)[i - 1, j - 1] = v
    exact = np.array((A.astype(object) @ B.astype(object)) % N, dtype=np.int64)
    assert np.array_equal(O.stripe_mul(A, B, N), exact)
    assert np.array_equal(O.exact_matmul_mod(A, B, N), exact)


def test_stripe_width_domain():
    # N=33554393 -> M=7 (F64); N=65521 -> 2098175; N=2^26 -> 1; N=2^26+1 -> 0
    assert O.find_max_stripe_ops(53, 33554393) == 7
    assert O.find_max_stripe_ops(53, 65521) == 2098175
    assert O.find_max_stripe_ops(53, 2 ** 26) == 1
    assert O.find_max_stripe_ops(53, 2 ** 26 + 1) == 0
    assert O.find_max_stripe_ops(24, 11) == 167771


def test_large_modulus_matmul_vs_python_ints():
    N = 33554393
    A = O.synth_matrix(3, 64, 200, N)
    B = O.synth_matrix(4, 200, 48, N)
    exact = np.array((A.astype(object) @ B.astype(object)) % N, dtype=np.int64)
    assert np.array_equal(O.stripe_mul(A, B, N), exact)
    assert np.array_equal(O.exact_matmul_mod(A, B, N), exact)
    A[:] = N - 1; B[:] = N - 1
    exact = np.array((A.astype(object) @ B.astype(object)) % N, dtype=np.int64)
    assert np.array_equal(O.matmul_mod(A, B, N), exact)


def test_mod_inv():
    for P in (7, 11, 65521, 33554393):
        for p in (1, 2, 3, P - 1, P // 2):
            assert (O.mod_inv(p, P) * p) % P == 1


def _check_pluq(A, N, U, L, pr, pc):
    PA = O.apply_row_perm(pr, np.mod(A, N))
    PAQ = O.apply_col_perm(pc, PA)
    assert np.array_equal(O.exact_matmul_mod(L[:, : U.shape[0]], U, N), PAQ)


def test_pluq_full_rank_matches_reference_loop():
    N = 65521
    for n, seed in ((17, 1), (64, 2), (90, 3)):
        A = O.synth_matrix(seed, n, n, N)
        U, L, pr, pc = O.pluq_reference(A, N)
        assert pc == []
        U2, L2, pr2, pc2, r = O.pluq(A, N)
        assert r == n and pc2 == [] and pr2 == pr
        assert np.array_equal(U, U2) and np.array_equal(L, L2)
        _check_pluq(A, N, U, L, pr, pc)
        assert all(U[i, i] == 1 for i in range(n))
        assert np.array_equal(np.tril(U, -1), np.zeros_like(U))
        assert np.array_equal(np.triu(L, 1), np.zeros_like(L))


def test_pluq_rank_deficient_correct_mode():
    N = 7
    rng = np.random.default_rng(5)
    X = rng.integers(0, N, size=(30, 9)); Y = rng.integers(0, N, size=(9, 40))
    A = O.exact_matmul_mod(X, Y, N)
    A[:, 3] = 0
    A[:, 11] = (3 * A[:, 1] + A[:, 2]) % N
    U, L, pr, pc, r = O.pluq(A, N)
    assert r == O.rank(A, N) <= 9
    _check_pluq(A, N, U, L, pr, pc)
    assert np.count_nonzero(U[r:, :]) == 0
    assert all(U[i, i] == 1 for i in range(r))
    R, piv = O.rref(A, N)
    assert len(piv) == r
    # rref idempotent and row space preserved: rref(R) == R
    R2, piv2 = O.rref(R, N)
    assert np.array_equal(R, R2) and piv == piv2
    for t, c in enumerate(piv):
        col = np.zeros(A.shape[0], dtype=np.int64); col[t] = 1
        assert np.array_equal(R[:, c], col)


def test_inverse_random():
    N = 7
    rng = np.random.default_rng(9)
    n = 40
    while True:
        A = rng.integers(0, N, size=(n, n))
        if O.rank(A, N) == n:
            break
    inv = O.inverse(A, N)
    assert np.array_equal(O.exact_matmul_mod(A, inv, N), np.eye(n, dtype=np.int64))
    A[:, 5] = A[:, 6]
    ok, none = O.is_invertible_with_inverse(A, N)
    assert ok is False and none is None


def test_karatsuba_reference_kernels_equal_direct():
    f = FX["karatsuba_params"]
    N1, N2 = f["N1"], f["N2"]
    rng = np.random.default_rng(11)
    n = 24
    A1 = rng.integers(0, N1, size=(n, n)); A2 = rng.integers(0, N2, size=(n, n))
    B1 = rng.integers(0, N1, size=(n, n)); B2 = rng.integers(0, N2, size=(n, n))
    C1, C2 = O.karatsuba_matmul(A1, A2, B1, B2, N1, N2)
    D1, D2 = O.karatsuba_matmul_direct(A1, A2, B1, B2, N1, N2)
    assert np.array_equal(C1, D1) and np.array_equal(C2, D2)
    x1 = rng.integers(0, N1, size=n); x2 = rng.integers(0, N2, size=n)
    C1, C2 = O.karatsuba_matmul(A1, A2, x1, x2, N1, N2)
    D1, D2 = O.karatsuba_matmul_direct(A1, A2, x1, x2, N1, N2)
    assert np.array_equal(C1, D1) and np.array_equal(C2, D2)
    # split/join round trip (KaratsubaMatrix.jl:372-397, :318-336)
    full = O.karatsuba_join(A1, A2, N1)
    s1, s2 = O.karatsuba_split(full, N1, N2)
    assert np.array_equal(s1, A1) and np.array_equal(s2, A2)


def test_c_oracle_matches_numpy_oracle():
    from oracle import oracle_c as OC
    for (m, k, n, N) in [(33, 70, 21, 11), (100, 129, 64, 65521), (64, 200, 48, 33554393), (17, 40, 9, 4294967291)]:
        A = O.synth_matrix(3, m, k, N); B = O.synth_matrix(4, k, n, N)
        assert np.array_equal(OC.matmul_mod(A, B, N), O.exact_matmul_mod(A, B, N))
    for (m, n, N) in [(20, 20, 7), (40, 25, 65521), (25, 40, 11)]:
        A = O.synth_matrix(5, m, n, N); A[:, 3] = 0
        E, L, pr, piv = OC.echelon(A, N)
        Eo, Lo, pro, pivo = O.echelon(A, N)
        assert np.array_equal(E, Eo) and np.array_equal(L, Lo) and pr == pro and piv == pivo


def test_hensel_lift_oracle():
    """hensel.jl:13-21 restated: the lifted T satisfies A*T == I modulo N^precision and reduces to the start value mod N."""
    rng = np.random.default_rng(5)
    for (p, prec, n) in ((7, 4, 6), (13, 7, 5), (2, 20, 4), (65521, 3, 4)):
        while True:
            A = rng.integers(0, p ** prec, size=(n, n))
            ok, T0 = O.is_invertible_with_inverse(A % p, p)
            if ok:
                break
        T = O.hensel_pseudoinverse(p, prec, A, T0)
        M = p ** prec
        I = np.eye(n, dtype=object)
        assert np.array_equal(np.array(A, dtype=object).dot(T) % M, I)
        assert np.array_equal(np.array(T % p, dtype=np.int64), T0)


def test_sampled_row_helpers_match_full_oracle():
    """oracle/sampled.py (used by the BASELINE-size GPU tests and bench.py's parity_check) agrees with the full-matrix oracle."""
    from oracle import sampled as S
    N = 33554393
    A = O.synth_matrix(5, 300, 200, N); B = O.synth_matrix(6, 200, 150, N)
    rows = S.pick_rows(300, 64, seed=1)
    assert len(rows) == 64 and rows[0] == 0 and rows[-1] == 299 and len(set(rows.tolist())) == 64
    assert np.array_equal(S.synth_rows(5, rows, 300, 200, N), A[rows])
    assert np.array_equal(S.synth_cols(6, [0, 7, 149], 200, N), B[:, [0, 7, 149]])
    C = O.exact_matmul_mod(A, B, N)
    assert S.check_product_rows(C[rows], A[rows], B, N)["match"]
    bad = C[rows].copy(); bad[3, 5] ^= 1
    assert S.check_product_rows(bad, A[rows], B, N) == {"rows": 64, "cols": 150, "mismatches": 1, "match": False}
    pairs = [(1, 5), (2, 5), (7, 3), (5, 1)]
    M = O.synth_matrix(9, 8, 8, 97)
    assert np.array_equal(M[S.perm_to_map(pairs, 8)], O.apply_row_perm(pairs, M))
    assert np.array_equal(M[:, S.perm_to_map(pairs, 8)], O.apply_col_perm(pairs, M))


def test_oracle_fast_paths_equal_the_python_integer_paths():
    """The oracle's int64 (N < 2^31) / uint64 (N <= 2^32) eliminations, triangular inverses and RREF, its float64-BLAS exact product
    (partial sums < 2^53) and its uint64 per-k product are the same numbers as the python-integer restatements they replace -- checked on
    random and on adversarial (all N-1) inputs up to the edge of their ranges."""
    rng = np.random.default_rng(11)
    for (m, n, N) in [(40, 40, 7), (60, 45, 65521), (45, 60, 33554393), (50, 50, 2 ** 31 - 1), (30, 30, 2), (50, 50, 2147483659), (45, 45, 4294967291)]:
        for adversarial in (False, True):
            A = rng.integers(0, N, size=(m, n), dtype=np.int64)
            if adversarial:
                A[:] = N - 1
                A[::3, ::2] = N - 2
                A[5, :] = 0
            fast = O.echelon(A, N)
            slow = O.echelon(A, N, python_ints=True)
            assert np.array_equal(fast[0], slow[0]) and np.array_equal(fast[1], slow[1]) and fast[2] == slow[2] and fast[3] == slow[3]
            again = O.echelon(A, N)  # served from the one-entry cache: equal, and not aliased with the first answer
            assert np.array_equal(again[0], fast[0]) and again[0] is not fast[0]
            fr, sr = O.pluq_reference(A, N), O.pluq_reference(A, N, python_ints=True)
            assert np.array_equal(fr[0], sr[0]) and np.array_equal(fr[1], sr[1]) and fr[2:] == sr[2:]
            if m == n:
                T = np.triu(rng.integers(0, N, size=(n, n), dtype=np.int64))
                T[np.arange(n), np.arange(n)] = rng.integers(1, N, size=n) if N > 2 else 1
                if adversarial:
                    T[np.triu_indices(n, 1)] = N - 1
                assert np.array_equal(O._fast_tri_inverse_upper(T, N), O._fast_tri_inverse_upper(T, N, python_ints=True))
            Eo, _, _, piv = slow  # RREF from python integers
            Ro = Eo.astype(object)
            for t in range(len(piv) - 1, 0, -1):
                fcol = Ro[:t, piv[t]].copy()
                Ro[:t, :] = (Ro[:t, :] - np.outer(fcol, Ro[t, :])) % N
            assert np.array_equal(O.rref(A, N)[0], np.array(Ro, dtype=np.int64))
    for (m, k, n, hi, N) in [(30, 700, 20, 1331, 1331), (20, 300, 25, 2 ** 26 - 1, 2 ** 26 - 5), (10, 50, 10, 2 ** 31, 4294967291), (12, 40, 9, 2 ** 32 - 1, 4294967291),
                              (8, 30, 8, 2 ** 32 - 1, 2 ** 52 - 47)]:
        A = rng.integers(0, hi + 1, size=(m, k), dtype=np.int64); B = rng.integers(0, hi + 1, size=(k, n), dtype=np.int64)
        A[0, :] = hi; B[:, 0] = hi
        want = np.array((A.astype(object) @ B.astype(object)) % N, dtype=np.int64)
        assert np.array_equal(O.exact_matmul_mod(A, B, N), want)
