"""GPU parity tests: every call goes through the C ABI (ctypes mirror == Julia shim) and is compared bit-for-bit
with the CPU oracle on the same seeded inputs, with the reference's own literal fixtures, and -- at BASELINE sizes --
through size-independent properties.  Run with `pytest -m gpu` on a B200."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402  (checker only)

FX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_fixtures.json")))


@pytest.fixture(scope="module")
def g():
    import gffm_b200
    gffm_b200.default_context()
    return gffm_b200


# ---------------------------------------------------------------- container ----------------------------------------------
def test_ctor_padding_mod_and_errors(g):
    A = np.array([[-3, 12], [5, 6], [7, -1]])
    M = g.CuModMatrix(A, 7)
    assert M.size() == (3, 2) and M.shape == (3, 2)
    img = M.unsafe_Array(np.int64)
    assert img.shape == (35, 34)  # +32 per dimension, CuModMatrix.jl:62
    assert np.array_equal(img, O.construct(A, 7))
    assert np.array_equal(M.Array(np.int64), np.mod(A, 7))
    for dt in (np.float32, np.float64, np.int64, np.int32, np.uint32):
        Mh = g.CuModMatrix(np.mod(A, 7).astype(dt), 7, elem_type=np.float64)
        assert np.array_equal(Mh.to_int(), np.mod(A, 7))
        assert Mh.Array().dtype == np.float64
    with pytest.raises(g.InexactError):
        g.CuModMatrix(np.array([[0.5, 1.0]]), 7)
    with pytest.raises(g.CuModArrayModulusMismatchException):
        g.CuModMatrix(np.array([[1]]), 2 ** 52 + 1)
    # mod=false keeps the raw values, new_size pads/crops
    assert np.array_equal(g.CuModMatrix(np.array([[9, 10]]), 7, mod=False).to_int(), [[9, 10]])
    assert g.CuModMatrix(np.ones((2, 2)), 7, new_size=(4, 5)).size() == (4, 5)
    v = g.CuModVector(np.array([1, 2, 15]), 11)
    assert v.shape == (3,) and np.array_equal(v.to_int(), [1, 2, 4])
    e = g.CuModMatrix(np.zeros((0, 5)), 7)
    assert e.Array().shape == (0, 5)


def test_basic_3x3_fixture(g):
    f = FX["basic_3x3"]
    A = g.CuModMatrix(np.array(f["A"]), f["N"]); B = g.CuModMatrix(np.array(f["B"]), f["N"]); s = f["scalar"]
    assert np.array_equal((A + B).to_int(), f["add"])
    assert np.array_equal((A - B).to_int(), f["sub"])
    assert np.array_equal((A * B).to_int(), f["matmul"])
    F = g.zeros(np.float32, 3, 3, f["N"]); g.elementwise_multiply_(F, A, B)
    assert np.array_equal(F.to_int(), f["elementwise_multiply"])
    assert np.array_equal((s + A).to_int(), f["scalar_add"])
    assert np.array_equal((A - s).to_int(), f["scalar_sub"])
    assert np.array_equal((s * A).to_int(), f["scalar_mul"])
    assert np.array_equal((A * -1).to_int(), f["negate"])
    assert np.array_equal((A ** 2).to_int(), f["pow2"])
    assert np.array_equal((A ** 0).to_int(), f["pow0"])
    assert A[0, 0] == 1 and A[2, 2] == 9


def test_matmul_fixtures_and_override_modulus(g):
    f = FX["matmul_2x3_3x2"]
    A = g.CuModMatrix(np.array(f["A"]), f["N"]); B = g.CuModMatrix(np.array(f["B"]), f["N"])
    assert np.array_equal((A * B).to_int(), f["C_mod11"])
    assert np.array_equal(g.mat_mul_gpu_type(A, B).to_int(), f["C_mod11"])
    assert np.array_equal(g.mat_mul_gpu_type(A, B, f["override_N"]).to_int(), f["C_mod7"])
    h = FX["matmul_inplace"]
    A = g.CuModMatrix(np.array(h["A"]), h["N"]); B = g.CuModMatrix(np.array(h["B"]), h["N"])
    C = g.zeros(np.float32, 2, 2, h["N"]); g.mat_mul_type_inplace_(C, A, B)
    assert np.array_equal(C.to_int(), h["C_mod9"])
    C2 = g.zeros(np.float32, 2, 2, h["override_N"]); g.mat_mul_type_inplace_(C2, A, B, h["override_N"])
    assert np.array_equal(C2.to_int(), h["C_mod3"])
    # mul! checks: modulus mismatch before size mismatch (CuModMatrix.jl:769-783)
    with pytest.raises(g.CuModArrayModulusMismatchException):
        g.mul_(g.zeros(np.float32, 2, 2, 7), A, B)
    with pytest.raises(g.CuModArraySizeMismatchException):
        g.mul_(g.zeros(np.float32, 3, 2, h["N"]), A, B)


def test_inplace_ops_and_modulus_override(g):
    N = 11
    rng = np.random.default_rng(3)
    A = rng.integers(0, N, size=(37, 53)); B = rng.integers(0, N, size=(37, 53))
    Ag, Bg = g.CuModMatrix(A, N), g.CuModMatrix(B, N)
    C = g.zeros(np.float32, 37, 53, N)
    assert np.array_equal(g.add_(C, Ag, Bg).to_int(), O.ew_add(A, B, N))
    assert np.array_equal(g.sub_(C, Ag, Bg).to_int(), O.ew_sub(A, B, N))
    assert np.array_equal(g.elementwise_multiply_(C, Ag, Bg).to_int(), O.ew_mul(A, B, N))
    assert np.array_equal(g.negate_(C, Ag).to_int(), O.ew_negate(A, N))
    assert np.array_equal(g.scalar_add_(C, Ag, 5).to_int(), O.ew_scalar_add(A, 5, N))
    assert np.array_equal(g.scalar_sub_(C, Ag, 5).to_int(), O.ew_scalar_sub(A, 5, N))
    assert np.array_equal(g.rscalar_sub_(C, Ag, 5).to_int(), O.ew_rscalar_sub(A, 5, N))
    assert np.array_equal(g.mul_(C, Ag, 3).to_int(), O.ew_scalar_mul(A, 3, N))
    assert np.array_equal((Ag / 3).to_int(), O.ew_scalar_div(A, 3, N))
    # mod_N override on mismatched-modulus operands (inplace_operations_test.jl:125-191)
    B13 = g.CuModMatrix(B, 13)
    with pytest.raises(g.CuModArrayModulusMismatchException):
        g.add_(C, Ag, B13)
    assert np.array_equal(g.add_(C, Ag, B13, mod_N=5).to_int(), O.ew_add(A, B, 5))
    # copy!, mod_elements!, fill!, zero!, change_modulus (inplace_operations_test.jl:216-255; basic :232-258)
    D = g.zeros(np.float32, 37, 53, N); g.copy_(D, Ag)
    assert D.equals(Ag)
    f = FX["fill_122_mod_11"]
    g.fill_(D, f["value"]); assert np.all(D.to_int() == f["expect"])
    g.zero_(D); assert np.all(D.to_int() == 0)
    E = g.change_modulus(Ag, 5)
    assert E.N == 5 and np.array_equal(E.to_int(), np.mod(A, 5)) and Ag.N == N
    g.change_modulus_no_alloc_(Ag, 3)
    assert Ag.N == 3 and np.array_equal(Ag.to_int(), np.mod(A, 3))
    assert np.array_equal(g.eye(np.float32, 5, N).to_int(), np.eye(5, dtype=np.int64))
    R = g.rand(np.float32, 40, 30, N, seed=7)
    assert R.to_int().max() < N and R.unsafe_Array(np.int64)[40:, :].sum() == 0  # padding stays zero (unlike CuModMatrix.jl:551-556)
    assert np.array_equal(g.transpose(Bg).to_int(), B.T)
    assert np.array_equal(g.synth(33, 17, 65521, 5).to_int(), O.synth_matrix(5, 33, 17, 65521))
    big = g.CuModMatrix(rng.integers(0, 2 ** 31, size=(9, 9)), 4294967291)
    assert np.array_equal((big + big).to_int(), O.ew_add(big.to_int(), big.to_int(), 4294967291))
    assert np.array_equal(g.elementwise_multiply_(g.zeros(np.float64, 9, 9, 4294967291), big, big).to_int(), O.ew_mul(big.to_int(), big.to_int(), 4294967291))


def test_permutation_fixture(g):
    f = FX["permutation_3x3"]
    P = [tuple(p) for p in f["P"]]
    A = g.CuModMatrix(np.array(f["A"], dtype=np.float64), f["N"])
    g.apply_col_perm_(P, A); assert np.array_equal(A.to_int(), f["col_perm"])
    g.apply_col_inv_perm_(P, A); assert np.array_equal(A.to_int(), f["A"])
    g.apply_row_perm_(P, A); assert np.array_equal(A.to_int(), f["row_perm"])
    g.apply_row_inv_perm_(P, A); assert np.array_equal(A.to_int(), f["A"])
    rng = np.random.default_rng(1)
    X = rng.integers(0, 11, size=(20, 30))
    Pl = [(int(a), int(b)) for a, b in rng.integers(1, 21, size=(15, 2))]
    Xg = g.CuModMatrix(X, 11); g.apply_row_perm_(Pl, Xg)
    assert np.array_equal(Xg.to_int(), O.apply_row_perm(Pl, X))
    g.apply_row_inv_perm_(Pl, Xg); assert np.array_equal(Xg.to_int(), X)
    assert np.array_equal(g.perm_array_to_matrix(Pl, 11, (20, 20), perm_stack=True).to_int(), O.perm_array_to_matrix(Pl, 20, True))
    assert g.mod_inv(3, 7) == O.mod_inv(3, 7) == 5


# ---------------------------------------------------------------- modular GEMM ----------------------------------------------
CASES = [  # (m, k, n, N)
    (1, 1, 1, 11), (3, 5, 2, 7), (100, 100, 100, 2 ** 11), (100, 100, 100, 11 ** 3), (129, 257, 255, 11), (300, 515, 700, 251),
    (257, 1000, 129, 65521), (300, 500, 700, 33554393), (128, 4096, 64, 33554393), (513, 130, 1025, 2 ** 26), (64, 300, 64, 4294967291),
]


@pytest.mark.parametrize("m,k,n,N", CASES)
def test_gemm_vs_oracle_all_algorithms(g, m, k, n, N):
    A = O.synth_matrix(11, m, k, N); B = O.synth_matrix(12, k, n, N)
    want = O.matmul_mod(A, B, N)
    algos = [g.capi.ALGO_AUTO, g.capi.ALGO_SIMT]
    if N <= 65536:
        algos.append(g.capi.ALGO_LIMB)
    if N <= 2 ** 31:
        algos.append(g.capi.ALGO_RNS)
    Ag, Bg = g.CuModMatrix(A, N), g.CuModMatrix(B, N)
    C0 = O.synth_matrix(13, m, n, N)
    for algo in algos:
        C = g.zeros(np.float64, m, n, N)
        g.mul_(C, Ag, Bg, algo=algo)
        assert np.array_equal(C.to_int(), want), f"algo {algo}"
        Cacc = g.CuModMatrix(C0, N); g.mul_(Cacc, Ag, Bg, algo=algo, mode=g.capi.GEMM_ADD)
        assert np.array_equal(Cacc.to_int(), np.mod(C0 + want, N)), f"algo {algo} add"
        Csub = g.CuModMatrix(C0, N); g.mul_(Csub, Ag, Bg, algo=algo, mode=g.capi.GEMM_SUB)
        assert np.array_equal(Csub.to_int(), np.mod(C0 - want, N)), f"algo {algo} sub"


@pytest.mark.parametrize("N,algo_name", [(11, "ALGO_LIMB"), (256, "ALGO_LIMB"), (65521, "ALGO_LIMB"), (65536, "ALGO_LIMB"), (33554393, "ALGO_RNS"), (2 ** 26, "ALGO_RNS")])
def test_gemm_adversarial_all_max(g, N, algo_name):
    """Exact accumulation budget: every entry N-1, K long enough to need more than one K chunk for L=2."""
    m, k, n = 130, 20000, 140
    A = np.full((m, k), N - 1, dtype=np.int64); B = np.full((k, n), N - 1, dtype=np.int64)
    want = np.full((m, n), (k * (N - 1) * (N - 1)) % N, dtype=np.int64)
    C = g.zeros(np.float64, m, n, N)
    g.mul_(C, g.CuModMatrix(A, N), g.CuModMatrix(B, N), algo=getattr(g.capi, algo_name))
    assert np.array_equal(C.to_int(), want)


def test_stripe_mul_cases_from_reference(g):
    for case in FX["stripe_cases"]["cases"]:
        n, N = case["n"], case["N"]
        rng = np.random.default_rng(n * 7 + N)
        A = rng.integers(case["lo"], case["hi"] + 1, size=(n, n)); B = rng.integers(case["lo"], case["hi"] + 1, size=(n, n))
        for (i, j, v, which) in case.get("poke", []):
            (A if which == "A" else B)[i - 1, j - 1] = v
        C = g.zeros(np.float32, n, n, N)
        g.stripe_mul_(C, g.CuModMatrix(A, N), g.CuModMatrix(B, N))
        assert np.array_equal(C.to_int(), O.exact_matmul_mod(A, B, N))


def test_baseline_config1_1024_mod_11(g):
    """BASELINE config 1: 1024x1024 * 1024x1024 mod 11 (seeds 1,2), bit-exact vs the integer oracle; the device-side
    synthetic generator must agree with the oracle's."""
    n, N = 1024, 11
    A = O.synth_matrix(1, n, n, N); B = O.synth_matrix(2, n, n, N)
    Ag = g.synth(n, n, N, 1); Bg = g.synth(n, n, N, 2)
    assert np.array_equal(Ag.to_int(), A) and np.array_equal(Bg.to_int(), B)
    assert np.array_equal((Ag * Bg).to_int(), O.matmul_mod(A, B, N))


def test_gemv(g):
    for (m, k, N) in [(100, 100, 11 ** 3), (500, 333, 33554393), (1, 7, 7), (130, 1, 11)]:
        A = O.synth_matrix(21, m, k, N); x = O.synth_matrix(22, k, 1, N).reshape(-1)
        z = g.zeros(np.float64, m, 1, N)
        g.gemv_(z, g.CuModMatrix(A, N), g.CuModVector(x, N))
        assert np.array_equal(z.to_int().reshape(-1), O.matvec_mod(A, x, N))
        z2 = g.CuModMatrix(A, N) * g.CuModVector(x, N)  # mul!(z,A,x) through the GEMM entry
        assert np.array_equal(z2.to_int().reshape(-1), O.matvec_mod(A, x, N))


def _freivalds(g, A, B, C, N, seed=99):
    """C == A*B (mod N) with probability >= 1 - 1/N per trial, using the independent GEMV kernel: A*(B*x) == C*x."""
    n = B.cols
    x = g.synth(n, 1, N, seed)
    Bx = g.zeros(np.float64, B.rows, 1, N); g.gemv_(Bx, B, x)
    ABx = g.zeros(np.float64, A.rows, 1, N); g.gemv_(ABx, A, Bx)
    Cx = g.zeros(np.float64, C.rows, 1, N); g.gemv_(Cx, C, x)
    return ABx.equals(Cx)


@pytest.mark.parametrize("n,N", [(8192, 33554393), (8192, 7), (16384, 65521), (16384, 33554393)])
def test_gemm_full_size_properties(g, n, N):
    """BASELINE sizes (configs 2 and the n=16384 metric): Freivalds check with the SIMT GEMV, linearity
    (A*(B+B') == A*B + A*B'), and all-(N-1) adversarial input with its closed-form answer."""
    A = g.synth(n, n, N, 5); B = g.synth(n, n, N, 6)
    C = g.zeros(np.float32, n, n, N); g.mul_(C, A, B)
    for s in (1, 2, 3):
        assert _freivalds(g, A, B, C, N, seed=100 + s)
    B2 = g.synth(n, n, N, 16); Bs = B + B2
    Cs = g.zeros(np.float32, n, n, N); g.mul_(Cs, A, Bs)
    g.mul_(C, A, B2, mode=g.capi.GEMM_ADD)
    assert C.equals(Cs)
    g.fill_(A, N - 1); g.fill_(B, N - 1); g.mul_(C, A, B)
    want = (n * (N - 1) * (N - 1)) % N
    ref = g.zeros(np.float32, n, n, N); g.fill_(ref, want)
    assert C.equals(ref)


# ---------------------------------------------------------------- elimination ----------------------------------------------
def _check_pluq_invariant(g, Ag, U, L, pr, pc):
    PA = g.copy(Ag); g.apply_row_perm_(pr, PA); g.apply_col_perm_(pc, PA)
    return (L * U).equals(PA) if L.cols == U.rows else None


@pytest.mark.parametrize("m,n,N,seed", [(1, 1, 7, 1), (10, 10, 7, 1), (33, 33, 2, 2), (100, 100, 65521, 3), (257, 257, 33554393, 4), (300, 200, 65521, 5),
                                        (200, 300, 11, 6), (1000, 1000, 13, 7), (700, 700, 4294967291, 8),
                                        # one case per arithmetic path of the panel kernels: left-looking with the 32-bit quotient
                                        # estimate (17 / 26 bits), left-looking with the 64-bit Barrett (29 bits), right-looking lazy
                                        # (30 bits); tall input whose row slices exceed the register kernels (shared-memory panel)
                                        (200, 200, 131071, 12), (260, 260, 67108859, 11), (300, 300, 536870909, 9),
                                        (300, 300, 805306457, 10), (20000, 40, 65521, 13)])
def test_pluq_lu_rref_inverse_vs_oracle(g, m, n, N, seed):
    A = O.synth_matrix(seed, m, n, N)
    if min(m, n) > 20 and seed in (5, 6, 7):
        A[:, 3] = 0; A[:, 11] = (3 * A[:, 1] + A[:, 2]) % N; A[m // 2, :] = A[0, :]
    Ag = g.CuModMatrix(A, N)
    U, L, pr, pc, rk = g.pluq_gpu_kernel(Ag, return_rank=True)
    Uo, Lo, pro, pco, rko = O.pluq(A, N)
    assert rk == rko and pr == pro and pc == pco
    assert np.array_equal(U.to_int(), Uo) and np.array_equal(L.to_int(), Lo)
    assert _check_pluq_invariant(g, Ag, U, L, pr, pc)
    E, L2, pr2, piv = g.lu(Ag, return_pivots=True)
    Eo, Lo2, pro2, pivo = O.echelon(A, N)
    assert np.array_equal(E.to_int(), Eo) and np.array_equal(L2.to_int(), Lo2) and pr2 == pro2 and piv == pivo
    R, pivr = g.rref(Ag, return_pivots=True)
    Ro, pivro = O.rref(A, N)
    assert np.array_equal(R.to_int(), Ro) and pivr == pivro
    assert g.rank(Ag) == rko
    if m == n:
        ok, inv = g.is_invertible_with_inverse(Ag)
        oko, invo = O.is_invertible_with_inverse(A, N)
        assert ok == oko == g.is_invertible(Ag)
        if ok:
            assert np.array_equal(inv.to_int(), invo)
            assert (Ag * inv).equals(g.eye(np.float32, n, N))
        else:
            with pytest.raises(g.MatrixNotInvertibleException):
                g.inverse(Ag)


def test_pluq_matches_reference_loop_on_full_rank(g):
    """For full-rank input no column swap occurs, so the blocked path must reproduce the reference's own loop
    (pluq_kernels.jl:46-157, restated literally in oracle.pluq_reference) bit for bit."""
    N = 65521
    A = O.synth_matrix(31, 150, 150, N)
    U, L, pr, pc = g.pluq_gpu_kernel(g.CuModMatrix(A, N))
    Uo, Lo, pro, pco = O.pluq_reference(A, N)
    assert pc == pco == [] and pr == pro
    assert np.array_equal(U.to_int(), Uo) and np.array_equal(L.to_int(), Lo)


def test_pluq_reference_quirk_mode(g):
    """GFFM_PIVOT_REFERENCE_QUIRK replays the reference's rank-deficient behaviour literally."""
    N = 7
    rng = np.random.default_rng(5)
    A = O.exact_matmul_mod(rng.integers(0, N, size=(30, 9)), rng.integers(0, N, size=(9, 40)), N)
    A[:, 3] = 0
    for M_ in (A, A[:, :30], O.synth_matrix(2, 25, 25, N)):
        U, L, pr, pc = g.pluq_gpu_kernel(g.CuModMatrix(M_, N), col_pivot_mode=g.capi.PIVOT_REFERENCE_QUIRK)
        Uo, Lo, pro, pco = O.pluq_reference(M_, N)
        assert pr == pro and pc == pco
        assert np.array_equal(U.to_int(), Uo) and np.array_equal(L.to_int(), Lo)


def test_de_rham_known_answer(g):
    f = FX["de_rham"]
    A = g.CuModMatrix(np.array(f["A"]), f["N"])
    flag, B = g.is_invertible_with_inverse(A)
    assert flag is True
    assert np.array_equal((A * B).to_int(), f["A_times_inverse"])


def test_triangular_inverse(g):
    f = FX["triangular_2x2"]
    for upper, key in ((True, "upper"), (False, "lower")):
        T = g.CuModMatrix(np.array(f[key], dtype=np.float64), f["N"])
        Ti = g.upper_triangular_inverse_no_copy(T) if upper else g.lower_triangular_inverse_no_copy(T)
        assert np.array_equal((T * Ti).to_int(), np.eye(2))
    I = g.CuModMatrix(np.eye(1000), 2)
    assert np.array_equal((I * g.upper_triangular_inverse_no_copy(I)).to_int(), np.eye(1000))
    rng = np.random.default_rng(4)
    for p in (3, 13, 97):          # triangular_test.jl:82-88 sweep (subset)
        for n in (33, 47, 64, 65):
            T = np.tril(rng.integers(1, p, size=(n, n)))
            Ti = g.lower_triangular_inverse_no_copy(g.CuModMatrix(T, p))
            assert np.array_equal(Ti.to_int(), O.lower_triangular_inverse(T, p))
    T = np.triu(rng.integers(1, 13, size=(1000, 1000)))
    Tg = g.CuModMatrix(T, 13)
    assert (Tg * g.upper_triangular_inverse_no_copy(Tg)).equals(g.eye(np.float32, 1000, 13))
    Tw = np.triu(rng.integers(1, 13, size=(200, 500)))   # wide upper: [T^-1; 0] (triangular_test.jl:43-45)
    Twg = g.CuModMatrix(Tw, 13)
    Ti = g.upper_triangular_inverse_no_copy(Twg)
    assert Ti.size() == (500, 200) and np.array_equal(Ti.to_int(), O.upper_triangular_inverse(Tw, 13))
    assert np.array_equal((Twg * Ti).to_int(), np.eye(200))
    with pytest.raises(g.InverseNotDefinedException):
        g.lower_triangular_inverse_no_copy(g.CuModMatrix(np.ones((5, 3)), 13))
    with pytest.raises(g.MatrixNotInvertibleException):
        g.lower_triangular_inverse_no_copy(g.CuModMatrix(np.array([[1, 0], [1, 0]]), 13))


def test_baseline_config4_lu_inverse_8192_mod_7(g):
    """BASELINE config 4: A = Pi*(Lo*Up) mod 7, n = 8192; properties A*A^-1 == I and P*A == L*U."""
    n, N = 8192, 7
    lo = O.synth_matrix(9, n, n, N); up = O.synth_matrix(10, n, n, N)
    Lo = np.tril(lo, -1) + np.eye(n, dtype=np.int64)
    Up = np.triu(up, 1) + np.diag(1 + (np.diag(up) % 6))
    Ag = g.CuModMatrix(Lo, N) * g.CuModMatrix(Up, N)
    rot = [(i + 1, ((i + 4097) % n) + 1) for i in range(0, 64)]
    g.apply_row_perm_(rot, Ag)
    ok, inv = g.is_invertible_with_inverse(Ag)
    assert ok and (Ag * inv).equals(g.eye(np.float32, n, N))
    U, L, pr, pc, rk = g.pluq_gpu_kernel(Ag, return_rank=True)
    assert rk == n and pc == [] and _check_pluq_invariant(g, Ag, U, L, pr, pc)


def test_baseline_config3_rref_pluq_16384_rank_deficient(g):
    """BASELINE config 3: 16384^2 rank-deficient mod 65521; rank, P*A*Q == L*U, echelon shape, RREF idempotence."""
    n, N, r = 16384, 65521, 15360
    X = g.synth(n, r, N, 7); Y = g.synth(r, n, N, 8)
    Ag = X * Y
    z = g.zeros(np.float32, n, 1, N)
    for c in (100, 5000, 16383):
        g.capi.check(Ag.lib.gffm_mat_copy_block(Ag.h, 0, c, z.h, 0, 0, n, 1))
    c17 = g.zeros(np.float32, n, 1, N); c42 = g.zeros(np.float32, n, 1, N)
    g.capi.check(Ag.lib.gffm_mat_copy_block(c17.h, 0, 0, Ag.h, 0, 17, n, 1))
    g.capi.check(Ag.lib.gffm_mat_copy_block(c42.h, 0, 0, Ag.h, 0, 4242, n, 1))
    comb = c17 * 3 + c42
    g.capi.check(Ag.lib.gffm_mat_copy_block(Ag.h, 0, 9000, comb.h, 0, 0, n, 1))
    U, L, pr, pc, rk = g.pluq_gpu_kernel(Ag, return_rank=True)
    assert rk <= r and rk >= r - 4
    assert _check_pluq_invariant(g, Ag, U, L, pr, pc)
    R, piv = g.rref(Ag, return_pivots=True)
    assert len(piv) == rk and 100 not in piv and 5000 not in piv and 9000 not in piv
    R2, piv2 = g.rref(R, return_pivots=True)
    assert piv2 == piv and R2.equals(R)


# ---------------------------------------------------------------- Karatsuba ----------------------------------------------
@pytest.mark.parametrize("n,cols,N1,N2", [(500, 1, 13 ** 4, 13 ** 3), (500, 500, 13 ** 4, 13 ** 3), (300, 200, 8191, 8191), (200, 64, 11, 11), (130, 130, 2 ** 26, 2 ** 26)])
def test_karatsuba_matmul(g, n, cols, N1, N2):
    """test/KaratsubaMatrix/basic_operations_test.jl:75-133 (mat x vec, N1=13^4, N2=13^3, n=500) plus mat x mat."""
    rng = np.random.default_rng(n + cols)
    A1 = rng.integers(0, N1, size=(n, n)); A2 = rng.integers(0, N2, size=(n, n))
    B1 = rng.integers(0, N1, size=(n, cols)); B2 = rng.integers(0, N2, size=(n, cols))
    AK = g.KaratsubaMatrix(g.CuModMatrix(A1, N1, elem_type=np.float64), g.CuModMatrix(A2, N1, elem_type=np.float64), N1, N2, N1 * N2)
    BK = g.KaratsubaMatrix(g.CuModMatrix(B1, N1, elem_type=np.float64), g.CuModMatrix(B2, N1, elem_type=np.float64), N1, N2, N1 * N2)
    CK = g.KaratsubaZeros(np.float64, n, cols, N1, N2)
    for K in (AK, BK, CK):
        g.initialize_plan_(K)
    g.KMatMul_(CK, AK, BK)
    full = (O.karatsuba_join(A1, A2, N1) @ O.karatsuba_join(B1, B2, N1)) % (N1 * N2)
    assert np.array_equal(CK.Array(), full)
    if N1 <= 2 ** 20 and n <= 300:
        C1, C2 = O.karatsuba_matmul(A1, A2, B1, B2, N1, N2)  # the reference's own kernels, restated
        assert np.array_equal(CK.data1.to_int(), C1) and np.array_equal(CK.data2.to_int(), C2)


def test_karatsuba_elementwise_and_split(g):
    N1, N2 = 13 ** 4, 13 ** 3
    M = N1 * N2
    rng = np.random.default_rng(8)
    A = rng.integers(0, M, size=(40, 30)); B = rng.integers(0, M, size=(40, 30))
    AK = g.MatToKMat(A, N1, N2); BK = g.MatToKMat(B, N1, N2)
    assert np.array_equal(AK.Array(), A)
    CK = g.KaratsubaZeros(np.float64, 40, 30, N1, N2)
    K = g.karatsuba
    assert np.array_equal(K.add_(CK, AK, BK).Array(), (A.astype(object) + B) % M)
    assert np.array_equal(K.sub_(CK, AK, BK).Array(), (A.astype(object) - B) % M)
    assert np.array_equal(K.scalar_multiply_(CK, AK, 97).Array(), (A.astype(object) * 97) % M)
    assert np.array_equal(K.negate_(CK, AK).Array(), (-A.astype(object)) % M)


# ---------------------------------------------------------------- pipelined host-to-host GEMM ----------------------------------------------
@pytest.mark.parametrize("m,k,n,N", [(100, 100, 100, 11), (700, 900, 1100, 33554393), (3000, 2500, 4100, 33554393), (2600, 3000, 2100, 65521), (4500, 1000, 2300, 251)])
def test_gemm_host_pipelined(g, m, k, n, N):
    """gffm_gemm_host (H2D / split / tcgen05 tiles / D2H on three streams) == the plain upload-mul-download sequence."""
    A = O.synth_matrix(41, m, k, N); B = O.synth_matrix(42, k, n, N)
    C = g.matmul_host(A, B, N)
    ref = (g.CuModMatrix(A, N) * g.CuModMatrix(B, N)).to_int()
    assert C.dtype == np.uint32 and np.array_equal(C.astype(np.int64), ref)
    if m * k * n <= 2 ** 30:
        assert np.array_equal(ref, O.matmul_mod(A, B, N))
    # unreduced inputs are reduced on the device (constructor semantics mod=true)
    if N < 2 ** 31:
        C2 = g.matmul_host(A + N, B, N)
        assert np.array_equal(C2, C)


# ---------------------------------------------------------------- BASELINE config 5 (n = 32768) ----------------------------------------------
def _freivalds_k(g, A1, A2, B1, B2, C1, C2, N1, N2, seed):
    """(A1 + N1*A2) * (B1 + N1*B2) == C1 + N1*C2 (mod N1*N2), checked modulo N2... on a random vector through plain GEMVs mod N1*N2 is not
    available for N1*N2 >= 2^32, so check the two consequences that pin the result: C1 == A1*B1 (mod N1) and
    C1 + N1*C2 == full product (mod N2), both via the independent SIMT GEMV kernel."""
    n = B1.cols
    # mod N1: C1 x == A1 (B1 x)
    x = g.synth(n, 1, N1, seed)
    t = g.zeros(np.float64, B1.rows, 1, N1); g.gemv_(t, B1, x, P=N1)
    l = g.zeros(np.float64, A1.rows, 1, N1); g.gemv_(l, A1, t, P=N1)
    r = g.zeros(np.float64, C1.rows, 1, N1); g.gemv_(r, C1, x, P=N1)
    return l.equals(r)


def test_baseline_config5_plain_32768(g):
    """32768 x 32768 plain mod-p product (9 moduli, 4 GiB operands): Freivalds with the SIMT GEMV + all-(N-1) closed form."""
    n, N = 32768, 33554393
    A = g.synth(n, n, N, 11); B = g.synth(n, n, N, 12)
    C = g.zeros(np.float32, n, n, N); g.mul_(C, A, B)
    for s in (1, 2):
        assert _freivalds(g, A, B, C, N, seed=200 + s)
    g.fill_(A, N - 1); g.fill_(B, N - 1); g.mul_(C, A, B)
    ref = g.zeros(np.float32, n, n, N); g.fill_(ref, (n * (N - 1) * (N - 1)) % N)
    assert C.equals(ref)


def test_baseline_config5_karatsuba_32768(g):
    """32768 x 32768 Karatsuba product, N1 = N2 = 8191 (M ~ 2^26): low limb against A1*B1 mod N1 (Freivalds), and the
    closed form for all-maximal limbs."""
    n, N1, N2 = 32768, 8191, 8191
    A1 = g.synth(n, n, N1, 13); A2 = g.synth(n, n, N2, 14); B1 = g.synth(n, n, N1, 15); B2 = g.synth(n, n, N2, 16)
    AK = g.KaratsubaMatrix(A1, A2, N1, N2); BK = g.KaratsubaMatrix(B1, B2, N1, N2)
    CK = g.KaratsubaZeros(np.float64, n, n, N1, N2)
    g.KMatMul_(CK, AK, BK)
    assert _freivalds_k(g, A1, A2, B1, B2, CK.data1, CK.data2, N1, N2, seed=300)
    # all limbs maximal: every entry of the full product is n * (M-1)^2 mod M = n mod M  (M-1 = -1)
    for X in (A1, B1):
        g.fill_(X, N1 - 1)
    for X in (A2, B2):
        g.fill_(X, N2 - 1)
    g.KMatMul_(CK, AK, BK)
    M = N1 * N2
    want = n % M
    r1 = g.zeros(np.float64, n, n, N1); g.fill_(r1, want % N1)
    r2 = g.zeros(np.float64, n, n, N1); g.fill_(r2, want // N1)
    assert CK.data1.equals(r1) and CK.data2.equals(r2)


# ---------------------------------------------------------------- multi-GPU layer on one GPU ----------------------------------------------
@pytest.mark.parametrize("m,k,n,N,offs", [
    (70, 50, 90, 33554393, [0, 17, 17, 60, 90]),            # small + unaligned + empty panel -> panel-by-panel dispatcher path
    (640, 1024, 1536, 33554393, [0, 256, 768, 1536]),       # pipelined, RNS
    (300, 1000, 1300, 65521, [0, 512, 1024, 1300]),         # pipelined, two positional limbs, ragged edges
    (1000, 640, 2048, 11, [0, 1024, 2048]),                 # pipelined, one limb
    (256, 4096, 512, 2 ** 26, [0, 512]),                    # a single panel
])
def test_gemm_panels_vs_oracle(g, m, k, n, N, offs):
    """gffm_gemm_panels without events == mul! == oracle (the panel pipeline of the multi-GPU layer, fed from resident B)."""
    import ctypes as C
    A = O.synth_matrix(51, m, k, N); B = O.synth_matrix(52, k, n, N)
    dA = g.CuModMatrix(A, N); dB = g.CuModMatrix(B, N); dC = g.zeros(np.float32, m, n, N)
    off = (C.c_int64 * len(offs))(*offs)
    g.capi.check(dC.lib.gffm_gemm_panels(dC.h, dA.h, dB.h, len(offs) - 1, off, None, None, 0, 0))
    assert np.array_equal(dC.to_int(), O.matmul_mod(A, B, N))
    with pytest.raises(g.GffmError):
        bad = (C.c_int64 * 3)(0, 5, n - 1)
        g.capi.check(dC.lib.gffm_gemm_panels(dC.h, dA.h, dB.h, 2, bad, None, None, 0, 0))
    with pytest.raises(g.CuModArraySizeMismatchException):
        g.capi.check(dC.lib.gffm_gemm_panels(dC.h, dB.h, dA.h, len(offs) - 1, off, None, None, 0, 0))


@pytest.mark.parametrize("N", [33554393, 65521])
def test_broadcast_matmul_streamed_panels(g, N):
    """BroadcastMatmul on one GPU: the 'broadcast' is a device copy on the communication stream, B changes every step.
    Checks the ready/consumed event protocol (panel p of step s+1 may overwrite the receive buffer as soon as step s has
    turned it into planes) -- every step must see exactly its own B."""
    import torch
    m, k, n = 1024, 2048, 4096
    ctx = g.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        A = g.synth(m, k, N, 61, ctx=ctx)
        ld = ((k + 31) // 32) * 32
        srcs = []
        for s in range(3):
            Bs = g.synth(k, n, N, 70 + s, ctx=ctx)
            t = torch.zeros((n, ld), dtype=torch.int32, device="cuda")
            W = g.CuModMatrix.wrap_device(t.data_ptr(), k, n, ld, N, ctx=ctx)
            g.copy_(W, Bs)
            srcs.append((t, W))
        Bt = torch.zeros((n, ld), dtype=torch.int32, device="cuda")
        B = g.CuModMatrix.wrap_device(Bt.data_ptr(), k, n, ld, N, ctx=ctx)
        Cs = [g.zeros(np.float32, m, n, N, ctx=ctx) for _ in range(3)]
        ctx.sync()
        cur = {"s": 0}
        panels = g.multigpu.col_panels(n, 4, align=g.multigpu.PANEL_ALIGN)
        assert panels == [(0, 1024), (1024, 2048), (2048, 3072), (3072, 4096)]

        def deliver(c0, c1):
            Bt[c0:c1].copy_(srcs[cur["s"]][0][c0:c1], non_blocking=True)

        for s in range(3):
            cur["s"] = s
            bm = g.multigpu.BroadcastMatmul(torch, None, Cs[s], A, B, Bt, panels, deliver=deliver) if s == 0 else bm
            bm.C = Cs[s]
            A.touch()
            bm.step()
        bm.finish()
        ctx.sync()
        torch.cuda.synchronize()
        for s in range(3):
            ref = g.zeros(np.float32, m, n, N, ctx=ctx)
            g.mul_(ref, A, srcs[s][1])
            assert Cs[s].equals(ref), f"step {s}"
        ctx.sync()


# ---------------------------------------------------------------- Hensel lifting ----------------------------------------------
@pytest.mark.parametrize("p,prec,n", [(7, 4, 40), (13, 7, 150), (2, 30, 64), (181, 4, 300)])
def test_hensel_pseudoinverse(g, p, prec, n):
    """hensel_pseudoinverse (triangular/hensel.jl:13-21) on CuModMatrix: mod-p inverse from the GPU elimination, lifted to
    p^prec (< 2^32) by tensor-core products; bit-exact vs the python-int oracle and A*T == I (mod p^prec)."""
    M = p ** prec
    seed = 90
    while True:
        A = O.synth_matrix(seed, n, n, M)
        ok, T0 = g.is_invertible_with_inverse(g.CuModMatrix(A % p, p))
        if ok:
            break
        seed += 1
    T0h = T0.to_int()
    T = g.hensel_pseudoinverse(p, prec, g.CuModMatrix(A, M), g.CuModMatrix(T0h, M))
    want = O.hensel_pseudoinverse(p, prec, A, T0h)
    assert np.array_equal(T.to_int().astype(object), want)
    I = (g.CuModMatrix(A, M) * T).to_int()
    assert np.array_equal(I, np.eye(n, dtype=np.int64))
    with pytest.raises(g.CuModArrayModulusMismatchException):
        g.hensel_pseudoinverse(p, prec, g.CuModMatrix(A % p, p), T0)


def test_hensel_pseudoinverse_karatsuba(g):
    """The same Newton lift on two-limb matrices: modulus 13^7 = 13^4 * 13^3 (the reference's Karatsuba test moduli,
    test/KaratsubaMatrix/basic_operations_test.jl:77-78), three doublings from a mod-13 inverse."""
    p, N1, N2, n = 13, 13 ** 4, 13 ** 3, 96
    M = N1 * N2
    seed = 95
    while True:
        A = O.synth_matrix(seed, n, n, M)
        ok, T0 = g.is_invertible_with_inverse(g.CuModMatrix(A % p, p))
        if ok:
            break
        seed += 1
    T0h = T0.to_int()
    AK = g.MatToKMat(A, N1, N2); TK = g.MatToKMat(T0h, N1, N2)
    g.karatsuba.hensel_pseudoinverse(3, AK, TK)            # precision 1 -> 2 -> 4 -> 8 >= 7
    want = O.hensel_pseudoinverse(p, 8, A, T0h) % M
    assert np.array_equal(TK.Array(), want)
    assert np.array_equal(np.array(A, dtype=object).dot(TK.Array()) % M, np.eye(n, dtype=object))


def test_karatsuba_array_interface(g):
    """The rest of the KaratsubaArray surface (KaratsubaMatrix.jl:302-316, :338-356, :428-503, :738-744): indexing, copy!, zero!,
    operators, Karatsubacopy, MatToKMat with a single modulus -- against python integers."""
    N1, N2 = 13 ** 4, 13 ** 3
    M = N1 * N2
    rng = np.random.default_rng(8)
    A = rng.integers(0, M, size=(40, 30)); B = rng.integers(0, M, size=(40, 30)); Cm = rng.integers(0, M, size=(30, 20))
    AK = g.MatToKMat(A, N1, N2); BK = g.MatToKMat(B, N1, N2); CK = g.MatToKMat(Cm, N1, N2)
    Ao, Bo, Co = A.astype(object), B.astype(object), Cm.astype(object)
    assert AK.size() == (40, 30)
    assert AK[3, 4] == int(A[3, 4])
    AK[3, 4] = 123456789; Ao[3, 4] = 123456789
    assert AK[3, 4] == 123456789 and int(AK.data1[3, 4]) == 123456789 % N1 and int(AK.data2[3, 4]) == 123456789 // N1
    assert np.array_equal((AK + BK).Array(), (Ao + Bo) % M)
    assert np.array_equal((AK - BK).Array(), (Ao - Bo) % M)
    assert np.array_equal((5 * AK).Array(), (5 * Ao) % M)
    assert np.array_equal((-AK).Array(), (-Ao) % M)
    assert np.array_equal((AK * CK).Array(), Ao.dot(Co) % M)
    assert np.array_equal(g.KMatToMat(AK), Ao)
    D = g.Karatsubacopy(AK)
    g.karatsuba.zero_(AK)
    assert not AK.Array().any() and np.array_equal(D.Array(), Ao)
    g.karatsuba.copy_(AK, D)
    assert np.array_equal(AK.Array(), Ao)
    K1 = g.MatToKMat(A % 8191, 8191)                       # single modulus: N1 = N2 = M (KaratsubaMatrix.jl:367-370)
    assert (K1.N1, K1.N2) == (8191, 8191) and np.array_equal(K1.Array(), (A % 8191).astype(object))
    with pytest.raises(ValueError):
        g.KaratsubaZeros(np.float64, 2, 2, N1, N2, use_gpu=False)


# ---------------------------------------------------------------- BASELINE configs pinned to the CPU oracle ----------------------------------------------
# The properties above (Freivalds, linearity, closed forms, P*A*Q == L*U through our own GEMM) are strong but self-referential.
# Here sampled ROWS of every BASELINE-size result are recomputed by the oracle (exact uint64 host arithmetic, oracle_c.matmul_mod)
# from inputs rebuilt with the oracle's own generator and compared bit for bit -- the reference's own criterion
# (/root/reference/test/CuModMatrix/stripe_mul_test.jl:31-50: `==` against mod.(A*B, N) on the host).
from oracle import sampled as S  # noqa: E402  (checker only)
from oracle import oracle_c as OC  # noqa: E402

NROWS = 64


def _check_device_generator(Bg, seed, n_rows, N, tag):
    """Downloads Bg (uint32) after checking on 32 sampled columns that the device generator equals the oracle's."""
    cols = S.pick_rows(Bg.cols, 32, seed=1000 + seed)
    Bh = Bg.to_u32()
    assert np.array_equal(Bh[:, cols].astype(np.int64), S.synth_cols(seed, cols, n_rows, N)), f"device generator != oracle generator ({tag})"
    return Bh


@pytest.mark.parametrize("n,N,sa,sb", [(8192, 33554393, 3, 4), (16384, 33554393, 5, 6), (16384, 65521, 5, 6), (16384, 11, 5, 6)])
def test_gemm_baseline_sizes_sampled_rows_vs_oracle(g, n, N, sa, sb):
    """BASELINE config 2 (8192^2 mod 33554393, seeds 3,4) and the n = 16384 metric (seeds 5,6; all three moduli): 64 rows of C bit-exact
    vs the oracle."""
    A = g.synth(n, n, N, sa); B = g.synth(n, n, N, sb)
    C = g.zeros(np.float32, n, n, N); g.mul_(C, A, B)
    rows = S.pick_rows(n, NROWS, seed=n + N % 1000)
    A_rows = S.synth_rows(sa, rows, n, n, N)
    assert np.array_equal(A.gather_rows(rows), A_rows)  # device generator == oracle generator on the sampled rows
    Bh = _check_device_generator(B, sb, n, N, f"n={n}")
    rep = S.check_product_rows(C.gather_rows(rows), A_rows, Bh, N)
    assert rep["match"], rep


def test_baseline_config5_plain_32768_sampled_rows_vs_oracle(g):
    """BASELINE config 5, plain product (seeds 11, 12): 64 rows of the 32768^2 result bit-exact vs the oracle."""
    n, N = 32768, 33554393
    A = g.synth(n, n, N, 11); B = g.synth(n, n, N, 12)
    C = g.zeros(np.float32, n, n, N); g.mul_(C, A, B)
    rows = S.pick_rows(n, NROWS, seed=55)
    A_rows = S.synth_rows(11, rows, n, n, N)
    assert np.array_equal(A.gather_rows(rows), A_rows)
    Bh = _check_device_generator(B, 12, n, N, "c5 plain")
    rep = S.check_product_rows(C.gather_rows(rows), A_rows, Bh, N)
    assert rep["match"], rep


def test_baseline_config5_karatsuba_32768_sampled_rows_vs_oracle(g):
    """BASELINE config 5, Karatsuba product N1 = N2 = 8191 (seeds 13-16): 64 rows of C1 + N1*C2 bit-exact vs the oracle's plain product
    of the joined values mod N1*N2 (joined values < 2^26, so oracle_c's uint64 accumulation is exact)."""
    n, N1, N2 = 32768, 8191, 8191
    M = N1 * N2
    A1 = g.synth(n, n, N1, 13); A2 = g.synth(n, n, N2, 14); B1 = g.synth(n, n, N1, 15); B2 = g.synth(n, n, N2, 16)
    CK = g.KaratsubaZeros(np.float64, n, n, N1, N2)
    g.KMatMul_(CK, g.KaratsubaMatrix(A1, A2, N1, N2), g.KaratsubaMatrix(B1, B2, N1, N2))
    rows = S.pick_rows(n, NROWS, seed=56)
    A_rows = S.synth_rows(13, rows, n, n, N1) + N1 * S.synth_rows(14, rows, n, n, N2)
    assert np.array_equal(A1.gather_rows(rows) + N1 * A2.gather_rows(rows), A_rows)
    Bh = _check_device_generator(B1, 15, n, N1, "c5 B1")
    B2h = _check_device_generator(B2, 16, n, N2, "c5 B2")
    Bh = Bh + np.uint32(N1) * B2h  # joined B, < 2^26
    del B2h
    want = S.product_rows(A_rows, Bh, M, in_bound=M)
    got = CK.data1.gather_rows(rows) + N1 * CK.data2.gather_rows(rows)
    assert np.array_equal(got, want)


def test_baseline_config4_sampled_rows_vs_oracle(g):
    """BASELINE config 4 (LU + inverse, 8192^2 mod 7): sampled rows of A*A^-1 (== rows of I) and of L*U (== rows of P*A) recomputed by
    the oracle from the DOWNLOADED factors -- no GEMM of ours in the check."""
    n, N = 8192, 7
    lo = O.synth_matrix(9, n, n, N); up = O.synth_matrix(10, n, n, N)
    Lo = np.tril(lo, -1) + np.eye(n, dtype=np.int64)
    Up = np.triu(up, 1) + np.diag(1 + (np.diag(up) % 6))
    rot = [(i + 1, ((i + 4097) % n) + 1) for i in range(0, 64)]
    Ah = O.apply_row_perm(rot, OC.matmul_mod(Lo, Up, N))  # the input itself comes from the oracle
    Ag = g.CuModMatrix(Ah, N)
    ok, inv = g.is_invertible_with_inverse(Ag)
    assert ok
    rows = S.pick_rows(n, NROWS, seed=57)
    rep = S.check_product_rows(np.eye(n, dtype=np.int64)[rows], Ah[rows], inv.to_u32(), N)
    assert rep["match"], rep
    U, L, pr, pc, rk = g.pluq_gpu_kernel(Ag, return_rank=True)
    assert rk == n and pc == []
    PA = O.apply_row_perm(pr, Ah)
    rep = S.check_product_rows(PA[rows], L.gather_rows(rows), U.to_u32(), N)
    assert rep["match"], rep


def _config3_matrix(g, n, r, N, zero_cols, dep):
    X = g.synth(n, r, N, 7); Y = g.synth(r, n, N, 8)
    Ag = X * Y
    z = g.zeros(np.float32, n, 1, N)
    for c in zero_cols:
        g.capi.check(Ag.lib.gffm_mat_copy_block(Ag.h, 0, c, z.h, 0, 0, n, 1))
    c17 = g.zeros(np.float32, n, 1, N); c42 = g.zeros(np.float32, n, 1, N)
    g.capi.check(Ag.lib.gffm_mat_copy_block(c17.h, 0, 0, Ag.h, 0, dep[1], n, 1))
    g.capi.check(Ag.lib.gffm_mat_copy_block(c42.h, 0, 0, Ag.h, 0, dep[2], n, 1))
    comb = c17 * 3 + c42
    g.capi.check(Ag.lib.gffm_mat_copy_block(Ag.h, 0, dep[0], comb.h, 0, 0, n, 1))
    return Ag


def test_baseline_config3_sampled_rows_vs_oracle(g):
    """BASELINE config 3 (RREF + PLUQ, 16384^2 rank-deficient mod 65521): sampled rows of L*U recomputed by the oracle from the downloaded
    factors equal the same rows of P*A*Q (permutations applied on the host by the oracle); the RREF annihilates sampled null-space
    relations: for the dependent column 9000 = 3*col 17 + col 4242 the RREF columns satisfy the same relation."""
    n, N, r = 16384, 65521, 15360
    Ag = _config3_matrix(g, n, r, N, (100, 5000, 16383), (9000, 17, 4242))
    U, L, pr, pc, rk = g.pluq_gpu_kernel(Ag, return_rank=True)
    Ah = Ag.to_u32()
    rows = S.pick_rows(n, NROWS, seed=58)
    rmap, cmap = S.perm_to_map(pr, n), S.perm_to_map(pc, n)
    PAQ_rows = Ah[rmap[rows]][:, cmap].astype(np.int64)  # rows of P*A*Q, permutations applied on the host (oracle semantics)
    del Ah
    rep = S.check_product_rows(PAQ_rows, L.gather_rows(rows), U.to_u32(), N)
    assert rep["match"], rep
    R, piv = g.rref(Ag, return_pivots=True)
    del U, L
    Rh = R.to_u32().astype(np.int64)
    assert np.array_equal(Rh[:, 9000], (3 * Rh[:, 17] + Rh[:, 4242]) % N)  # row operations preserve column relations
    assert not Rh[:, [100, 5000, 16383]].any()
    assert len(piv) == rk and np.array_equal(Rh[np.arange(rk), piv], np.ones(rk, dtype=np.int64))  # unit pivots ...
    sub = Rh[:, piv]
    assert np.count_nonzero(sub) == rk  # ... alone in their columns: reduced form


def test_config3_shape_2048_rank_pivots_echelon_vs_oracle(g):
    """The config-3 construction at n = 2048 (r = 1920, zero columns 100/1000/2047, column 900 = 3*col 17 + col 424), where the C oracle's
    full echelon elimination takes seconds: rank, pivot columns, row transpositions, U and L are bit-exact vs the oracle, and the RREF is
    the unique solution of T*R == E (T = pivot-column block of the oracle's echelon form E, unit upper triangular), checked by the oracle."""
    n, N, r = 2048, 65521, 1920
    Ag = _config3_matrix(g, n, r, N, (100, 1000, 2047), (900, 17, 424))
    Ah = Ag.to_int()
    W, Lh, perm_rows, pivcols = OC.echelon(Ah, N)
    rk = len(pivcols)
    assert rk <= r and 100 not in pivcols and 1000 not in pivcols and 900 not in pivcols
    U, L, pr, piv = g.lu(Ag, return_pivots=True)
    assert piv == pivcols
    assert pr == perm_rows
    assert np.array_equal(U.to_int(), W)
    assert np.array_equal(L.to_int(), Lh)
    assert g.rank(Ag) == rk
    R, piv2 = g.rref(Ag, return_pivots=True)
    assert piv2 == pivcols
    Rh = R.to_int()
    assert not Rh[rk:].any()
    T = W[:rk][:, pivcols]
    assert np.array_equal(np.triu(T, 1) + np.eye(rk, dtype=np.int64), T)
    assert np.array_equal(OC.matmul_mod(T, Rh[:rk], N), W[:rk])


# ---------------------------------------------------------------- round-2 additions: GEMV, limits ----------------------------------------------
@pytest.mark.parametrize("m,k,N,P", [(1000, 3000, 4294967291, 65521), (700, 5000, 33554393, 7), (4100, 4099, 2 ** 26, 2 ** 26), (513, 70000, 65521, 65521),
                                     (3, 5, 11, 11), (2048, 2048, 4294967291, 4294967291)])
def test_gemv_modulus_override_and_k_split(g, m, k, N, P):
    """mul!(z,A,x;R,P) (CuModMatrix.jl:816-836): entries are bounded by the operands' modulus N, the result is reduced mod P -- also when
    P << N (products near 2^64 must not wrap the accumulator), with the K range split across CTAs and ragged row tails."""
    A = O.synth_matrix(31, m, k, N); x = O.synth_matrix(32, k, 1, N).reshape(-1)
    z = g.zeros(np.float64, m, 1, P)
    g.gemv_(z, g.CuModMatrix(A, N), g.CuModVector(x, N), P=P)
    want = np.array([int(sum(int(a) * int(b) for a, b in zip(A[i], x)) % P) for i in range(0, m, max(1, m // 7))])
    assert np.array_equal(z.to_int().reshape(-1)[::max(1, m // 7)], want)
    if N < 2 ** 31:  # full comparison through the oracle's striped GEMV
        assert np.array_equal(z.to_int().reshape(-1), OC.matmul_mod(A, x.reshape(-1, 1), P, in_bound=N).reshape(-1))


def test_gemv_rejects_aliasing_and_mismatch(g):
    A = g.synth(8, 8, 11, 1); x = g.synth(8, 1, 11, 2)
    with pytest.raises(g.GffmError):
        g.gemv_(x, A, x)
    with pytest.raises(g.CuModArrayModulusMismatchException):
        g.gemv_(g.zeros(np.float32, 8, 1, 13), A, x)


@pytest.mark.parametrize("n,N1,N2", [(777, 13 ** 4, 13 ** 3), (4096, 8191, 8191), (5, 2 ** 26, 2 ** 26), (1500, 11, 11)])
def test_karatsuba_matvec_single_pass(g, n, N1, N2):
    """KMatMul! on vectors / KMatMul_gemv! (KaratsubaMatrix.jl:238-300): the fused one-pass kernel equals the exact product mod N1*N2."""
    M = N1 * N2
    rng = np.random.default_rng(n)
    A = rng.integers(0, M, size=(n, n + 3), dtype=np.int64); x = rng.integers(0, M, size=(n + 3,), dtype=np.int64)
    A[0, :] = M - 1; x[:] = np.where(np.arange(n + 3) % 5 == 0, M - 1, x)
    AK = g.KaratsubaMatrix.from_array(A, N1, N2, M); xK = g.KaratsubaMatrix.from_array(x.reshape(-1, 1), N1, N2, M)
    zK = g.KaratsubaZeros(np.float64, n, 1, N1, N2)
    g.KMatMul_(zK, AK, xK)
    want = np.array([int(sum(int(a) * int(b) for a, b in zip(A[i], x)) % M) for i in range(n)], dtype=object)
    got = np.array([int(v) for v in np.asarray(zK.Array()).reshape(-1)], dtype=object)
    assert np.array_equal(got, want)


def test_karatsuba_inner_dimension_above_65536(g):
    """KMatMul! has no inner-dimension limit (KaratsubaMatrix.jl:133-204): K = 70000 runs as two chunks joined by the two-limb carry add."""
    m, k, n, N1, N2 = 130, 70000, 140, 8191, 8191
    M = N1 * N2
    rng = np.random.default_rng(7)
    A = rng.integers(0, M, size=(m, k), dtype=np.int64); B = rng.integers(0, M, size=(k, n), dtype=np.int64)
    A[:, :100] = M - 1; B[:100, :] = M - 1
    CK = g.KaratsubaZeros(np.float64, m, n, N1, N2)
    g.KMatMul_(CK, g.KaratsubaMatrix.from_array(A, N1, N2, M), g.KaratsubaMatrix.from_array(B, N1, N2, M))
    want = OC.matmul_mod(A, B, M, in_bound=M)
    assert np.array_equal(np.asarray(CK.Array()).astype(np.int64), want)


@pytest.mark.parametrize("N,algo_name", [(33554393, "ALGO_RNS"), (65521, "ALGO_LIMB")])
def test_gemm_more_than_65535_columns(g, N, algo_name):
    """Tensor-core paths with n > 65535 (CUDA grid.y limit): split / CRT launches are issued in column slabs."""
    m, k, n = 130, 256, 66000
    A = O.synth_matrix(41, m, k, N); B = O.synth_matrix(42, k, n, N)
    C = g.zeros(np.float32, m, n, N)
    g.mul_(C, g.CuModMatrix(A, N), g.CuModMatrix(B, N), algo=getattr(g.capi, algo_name))
    assert np.array_equal(C.to_int(), OC.matmul_mod(A, B, N))


def test_elimination_reports_zero_divisor_pivot(g):
    """Composite modulus whose largest residue is a zero divisor: the reference's rule picks it as pivot (pluq_kernels.jl:189) and has no
    inverse for it (mod_inv, :11-31); we raise the reference's CuModMatrixModulusNotPrimeException (CuModMatrix.jl:21) instead of returning
    a wrong factorisation."""
    A = g.CuModMatrix(np.array([[8, 1], [3, 5]]), 9 + 1)  # mod 10: pivot 8 = max, gcd(8, 10) = 2
    with pytest.raises(g.CuModMatrixModulusNotPrimeException):
        g.inverse(A)
    with pytest.raises(g.CuModMatrixModulusNotPrimeException):
        g.pluq_gpu_kernel(A)
    B = np.array([[9, 0], [3, 7]])  # pivots 9 and 7: both units mod 10 -> a valid factorisation over Z/10
    U, L, pr, pc = g.pluq_gpu_kernel(g.CuModMatrix(B, 10))
    assert np.array_equal((L.to_int() @ U.to_int()) % 10, O.apply_col_perm(pc, O.apply_row_perm(pr, B)))


def test_wide_moduli_container_and_elementwise(g):
    """Moduli 2^32 < N <= 2^52 (the reference container accepts N <= 2^52, CuModMatrix.jl:55-59): uint64 storage behind the same
    constructor / Array / zeros / eye / fill! / copy! / getindex / setindex! / change_modulus and the elementwise API with its mod_N
    override; exact against python integers (products need 104 bits).  Products / eliminations above 2^32 are refused (they go
    through KaratsubaMatrix, as in the reference)."""
    N = 2 ** 52 - 47
    rng = np.random.default_rng(5)
    Ah = rng.integers(-(2 ** 62), 2 ** 62, size=(70, 45), dtype=np.int64); Bh = rng.integers(0, N, size=(70, 45), dtype=np.int64)
    Ah[0, 0] = N - 1; Bh[0, 0] = N - 1
    A = g.CuModMatrix(Ah, N, elem_type=np.float64); B = g.CuModMatrix(Bh, N, elem_type=np.float64)
    Ai = np.array([[int(v) % N for v in row] for row in Ah], dtype=object); Bi = Bh.astype(object)
    assert np.array_equal(A.to_int().astype(object), Ai)
    assert A.unsafe_Array(np.int64).shape == (70 + 32, 45 + 32) and not A.unsafe_Array(np.int64)[70:, :].any()
    assert np.array_equal(A.Array(np.float64), Ai.astype(np.float64))
    assert np.array_equal((A + B).to_int().astype(object), (Ai + Bi) % N)
    assert np.array_equal((A - B).to_int().astype(object), (Ai - Bi) % N)
    C = g.zeros(np.float64, 70, 45, N); g.elementwise_multiply_(C, A, B)
    assert np.array_equal(C.to_int().astype(object), (Ai * Bi) % N)
    s = 2 ** 51 + 12345
    assert np.array_equal((A * s).to_int().astype(object), (Ai * s) % N)
    assert np.array_equal((s - A).to_int().astype(object), (s - Ai) % N)
    assert np.array_equal((A / 3).to_int().astype(object), (Ai * pow(3, -1, N)) % N)
    g.add_(C, A, B, mod_N=2 ** 40 + 15)
    assert np.array_equal(C.to_int().astype(object), (Ai % (2 ** 40 + 15) + Bi % (2 ** 40 + 15)) % (2 ** 40 + 15))
    E = g.eye(np.float64, 5, N); assert np.array_equal(E.to_int(), np.eye(5, dtype=np.int64))
    g.fill_(C, -1); assert C[3, 4] == N - 1
    C[1, 2] = N + 5; assert C[1, 2] == 5
    D = g.copy(A); assert D.equals(A) and not D.equals(B)
    g.change_modulus_no_alloc_(D, 2 ** 45 + 59)
    assert np.array_equal(D.to_int().astype(object), Ai % (2 ** 45 + 59))
    R = g.rand(np.float64, 40, 40, N, seed=3); Rh = R.to_int()
    assert Rh.min() >= 0 and Rh.max() < N and Rh.max() > 2 ** 40 and not R.unsafe_Array(np.int64)[40:, :].any()
    with pytest.raises(g.InexactError):
        A.Array(np.float32)
    with pytest.raises(g.GffmError):
        A * B.__class__(Bh.T.copy(), N, elem_type=np.float64)  # matrix product above 2^32: KaratsubaMatrix territory
    with pytest.raises(g.GffmError):
        g.pluq_gpu_kernel(g.CuModMatrix(Ah[:20, :20], N, elem_type=np.float64))
    with pytest.raises(g.CuModArrayModulusMismatchException):
        g.CuModMatrix(Ah, 2 ** 52 + 1)


# ---------------------------------------------------------------- allocations (test/CuModMatrix/allocations_test.jl:21-52) ----
def test_inplace_methods_allocate_no_device_memory(g):
    """The reference's allocation contract, restated with the library's own allocation counter (gffm_alloc_stats) in the role of
    CUDA.@timed's gpu_bytes (which likewise counts the requests made to CUDA.jl's pool, not driver-level page mappings such as lazily
    loaded kernel code): the in-place elementwise methods request ZERO device bytes (allocations_test.jl:21-46, `== 0`), a warm mul!
    (matrix and vector form) requests none either (:48-52, `< 20`)."""
    n = 3003
    rng = np.random.default_rng(5)
    Ad = rng.integers(0, 11, size=(n, n)); Bd = rng.integers(0, 11, size=(n, n)); xd = rng.integers(0, 11, size=n)
    A = g.CuModMatrix(Ad, 11); B = g.CuModMatrix(Bd, 11); x = g.CuModVector(xd, 11)
    C_ = g.zeros(np.float32, n, n, 11); z = g.zeros(np.float32, n, 1, 11)
    ctx = g.default_context()
    cases = [
        ("add!", lambda: g.add_(C_, A, B)), ("sub!", lambda: g.sub_(C_, A, B)), ("negate!", lambda: g.negate_(C_, A)),
        ("scalar_add!", lambda: g.scalar_add_(C_, A, 2)), ("scalar_sub!", lambda: g.scalar_sub_(C_, A, 2)),
        ("mul!(C,A,2)", lambda: g.mul_(C_, A, 2)), ("elementwise_multiply!", lambda: g.elementwise_multiply_(C_, A, B)),
        ("copy!", lambda: g.copy_(C_, A)), ("mod_elements!", lambda: g.mod_elements_(C_, 3)),
    ]
    for name, fn in cases:  # no warm-up: these never allocate
        ctx.sync()
        b0, c0 = ctx.alloc_stats()
        fn()
        ctx.sync()
        b1, c1 = ctx.alloc_stats()
        assert (b1 - b0, c1 - c0) == (0, 0), name
    g.mod_elements_(C_, 11)
    for name, fn in [("mul!(C,A,B)", lambda: g.mul_(C_, A, B)), ("mul!(z,A,x)", lambda: g.mul_(z, A, x))]:
        fn()  # the reference primes CUDA.@timed the same way: workspaces and operand-plane caches exist after the first call
        ctx.sync()
        b0, c0 = ctx.alloc_stats()
        fn()
        ctx.sync()
        b1, _ = ctx.alloc_stats()
        assert b1 - b0 < 20, (name, b1 - b0)
    assert np.array_equal(C_.to_int(), np.mod(Ad.astype(np.float64) @ Bd.astype(np.float64), 11).astype(np.int64))  # exact: sums < 2^53
    assert np.array_equal(z.to_int().ravel(), (Ad.dot(xd) % 11))
