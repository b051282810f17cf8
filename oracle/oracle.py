"""CPU oracle for the GPUFiniteFieldMatrices.jl hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, in numpy, the algorithms of the reference package so that the
CUDA path can be checked bit-for-bit.  It is never imported by the product
(`gpufinitefieldmatrices.jl_b200/`); only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may use it.

Parity status: the reference (Julia + CUDA.jl) cannot be executed in this image (no
`julia`), so the oracle is pinned against the literal known-answer fixtures of the
reference's own test-suite (tests/golden/reference_fixtures.json, transcribed from
/root/reference/test/CuModMatrix/*.jl -- see tests/test_oracle_golden.py) and against an
independent exact big-integer evaluation (python ints).  PLUQ on non-identity input,
rank-deficient behaviour, rref/lu and Karatsuba mat*mat are NOT pinned by any reference
test ("parity unpinned by reference tests"); for those the oracle is the
literal restatement of the reference's loop plus algebraic invariants.

All reference citations are relative to /root/reference/.
"""
from __future__ import annotations

import numpy as np

PAD = 32  # TILE_WIDTH, src/CuModMatrix/CuModMatrix.jl:2 ; padded = size + 32 (:62)


# --------------------------------------------------------------------------------------
# Synthetic inputs: val(seed,i,j) = splitmix64(seed ^ (j*rows+i)) mod N
# --------------------------------------------------------------------------------------
def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def synth_matrix(seed: int, rows: int, cols: int, N: int) -> np.ndarray:
    """Column-major counter-based generator; returns int64 (rows, cols) in [0, N)."""
    j = np.arange(cols, dtype=np.uint64)[None, :]
    i = np.arange(rows, dtype=np.uint64)[:, None]
    with np.errstate(over="ignore"):
        idx = j * np.uint64(rows) + i
    v = splitmix64(np.uint64(seed) ^ idx)
    return (v % np.uint64(N)).astype(np.int64)


# --------------------------------------------------------------------------------------
# Container semantics: CuModMatrix.jl:53-99 (ctor), :256-261 (Array)
# --------------------------------------------------------------------------------------
def construct(A: np.ndarray, N: int, do_mod: bool = True) -> np.ndarray:
    """Padded (rows+32, cols+32) integer image of CuModMatrix(A, N).

    Floored mod on construction (mod_ops.jl:8, CuModMatrix.jl:90-92); padding is zero
    (CuModMatrix.jl:67).  N > 2^52 is rejected like CuModMatrix.jl:55-59.
    """
    if N > 2 ** 52:
        raise ValueError("CuModArrayModulusMismatchException: modulus exceeds 2^52")
    A = np.asarray(A)
    if A.ndim == 1:
        A = A[:, None]
    Ai = np.asarray(np.rint(A), dtype=np.int64)
    if not np.array_equal(Ai, A):
        raise ValueError("InexactError: non-integer entry")  # convert.(T,A), :70-86
    out = np.zeros((A.shape[0] + PAD, A.shape[1] + PAD), dtype=np.int64)
    out[: A.shape[0], : A.shape[1]] = np.mod(Ai, N) if do_mod else Ai
    return out


# --------------------------------------------------------------------------------------
# Modular GEMM: kernel_mul/stripe_mul.jl:13-27 (stripe width), :175-244 (loop)
# --------------------------------------------------------------------------------------
def find_max_stripe_ops(bits: int, N: int) -> int:
    """stripe_mul.jl:13-27: M = floor((2^bits - 1)/(N-1)^2) - 1 (bits = 53/24/11)."""
    if N <= 1:
        return 1 << 30
    return ((1 << bits) - 1) // ((N - 1) ** 2) - 1


def stripe_mul(A: np.ndarray, B: np.ndarray, N: int, in_bound: int | None = None) -> np.ndarray:
    """C = A*B mod N, the reference algorithm on float64 (stripe_mul.jl:175-244).

    K is cut into stripes narrow enough that every partial sum stays below 2^53, a
    floored mod follows each stripe.  `in_bound` (R of the vector form, :96-114) bounds the
    inputs when it differs from N (Karatsuba sub-products).  Result entries in [0, N).
    """
    A = np.asarray(A, dtype=np.int64)
    B = np.asarray(B, dtype=np.int64)
    assert A.shape[1] == B.shape[0]
    R = N if in_bound is None else in_bound
    K = A.shape[1]
    # width such that (w+1)*(R-1)^2 + (N-1) < 2^53
    w = max(1, min(K if K > 0 else 1, find_max_stripe_ops(53, R)))
    if N > 2 ** 26 or w < 1:
        # beyond the reference's float64 domain: exact python-int fallback (Karatsuba P up to 2^52)
        return exact_matmul_mod(A, B, N)
    Af = A.astype(np.float64)
    Bf = B.astype(np.float64)
    C = np.zeros((A.shape[0], B.shape[1]), dtype=np.float64)
    for k0 in range(0, K, w):
        k1 = min(K, k0 + w)
        C += Af[:, k0:k1] @ Bf[k0:k1, :]
        np.mod(C, float(N), out=C)
    return C.astype(np.int64)


def exact_matmul_mod(A: np.ndarray, B: np.ndarray, N: int) -> np.ndarray:
    """mod.(A*B, N) on exact integers -- the reference TESTS' own ground truth
    (test/CuModMatrix/stripe_mul_test.jl:11,26,48).  int64 with K-chunks that cannot
    overflow; python ints when a single product can exceed 2^63."""
    A = np.asarray(A, dtype=np.int64)
    B = np.asarray(B, dtype=np.int64)
    amax = int(np.abs(A).max()) if A.size else 0
    bmax = int(np.abs(B).max()) if B.size else 0
    K = A.shape[1]
    if amax * bmax >= 2 ** 62 and amax < 2 ** 32 and bmax < 2 ** 32 and N <= 2 ** 63 and A.size and B.size and int(A.min()) >= 0 and int(B.min()) >= 0:
        # 32-bit residues: every single product fits uint64, so one exact column-times-row step per k (reduced at once); the same
        # numbers as the python-integer product below, without its per-element interpreter cost
        Au, Bu = A.astype(np.uint64), B.astype(np.uint64)
        C = np.zeros((A.shape[0], B.shape[1]), dtype=np.uint64)
        for k in range(K):
            C = (C + (Au[:, k:k + 1] * Bu[k:k + 1, :]) % N) % N  # C < N <= 2^63 and the reduced product < N: the sum fits
        return C.astype(np.int64)
    if amax * bmax >= 2 ** 62:
        Ao = A.astype(object)
        Bo = B.astype(object)
        return np.array((Ao @ Bo) % N, dtype=np.int64) if N < 2 ** 63 else (Ao @ Bo) % N
    if amax * bmax < 2 ** 47 and N < 2 ** 62:  # at least 64 terms per BLAS call; larger entries take the int64 chunks below
        # float64 BLAS on K-chunks whose partial sums stay below 2^53: every product and every partial sum is an exactly representable
        # integer, so the result is the exact integer product whatever the summation order (same numbers as the int64 path, ~50x faster)
        chunk = max(1, min(max(K, 1), (2 ** 53 - 1) // max(1, amax * bmax)))
        Af, Bf = A.astype(np.float64), B.astype(np.float64)
        C = np.zeros((A.shape[0], B.shape[1]), dtype=np.int64)
        for k0 in range(0, K, chunk):
            k1 = min(K, k0 + chunk)
            C = (C + np.mod((Af[:, k0:k1] @ Bf[k0:k1, :]).astype(np.int64), N)) % N
        return np.mod(C, N)
    chunk = max(1, min(max(K, 1), (2 ** 62) // max(1, amax * bmax)))
    C = np.zeros((A.shape[0], B.shape[1]), dtype=np.int64)
    for k0 in range(0, K, chunk):
        k1 = min(K, k0 + chunk)
        C = (C + (A[:, k0:k1] @ B[k0:k1, :]) % N) % N
    return np.mod(C, N)


def matmul_mod(A, B, N):
    """mul!(C,A,B) / A*B: CuModMatrix.jl:767-787, kernel_ops/mul_ops.jl:54-58."""
    if N <= 2 ** 26:
        return stripe_mul(A, B, N)
    return exact_matmul_mod(A, B, N)


def matvec_mod(A, x, N, in_bound=None):
    """mul!(z,A,x;R,P): CuModMatrix.jl:816-836, stripe_mul.jl:82-168."""
    x = np.asarray(x, dtype=np.int64).reshape(-1, 1)
    return stripe_mul(A, x, N, in_bound).reshape(-1)


# --------------------------------------------------------------------------------------
# Elementwise ops: kernel_ops/{mod,add,sub,mul,div}_ops.jl  (all floored mod)
# --------------------------------------------------------------------------------------
def mod_inv(p: int, P: int) -> int:
    """rref_lu_pluq/pluq_kernels.jl:11-31 -- extended Euclid, result in [0,P)."""
    inv, new_inv = 0, 1
    rem, new_rem = P, p
    while new_rem != 0:
        q = rem // new_rem
        inv, new_inv = new_inv, inv - q * new_inv
        rem, new_rem = new_rem, rem - q * new_rem
    if inv < 0:
        inv += P
    return inv


def ew_add(A, B, N):  # add_ops.jl:23-30
    return np.mod(np.asarray(A, np.int64) + np.asarray(B, np.int64), N)


def ew_sub(A, B, N):  # sub_ops.jl:33-40
    return np.mod(np.asarray(A, np.int64) - np.asarray(B, np.int64), N)


def ew_mul(A, B, N):  # mul_ops.jl:23-30 (elementwise_multiply!)
    return np.array((np.asarray(A).astype(object) * np.asarray(B).astype(object)) % N, dtype=np.int64)


def ew_scalar_add(A, s, N):  # add_ops.jl:33-40
    return np.mod(np.asarray(A, np.int64) + int(s), N)


def ew_scalar_sub(A, s, N):  # sub_ops.jl:43-50
    return np.mod(np.asarray(A, np.int64) - int(s), N)


def ew_rscalar_sub(A, s, N):  # sub_ops.jl:53-69  (s - A)
    return np.mod(int(s) - np.asarray(A, np.int64), N)


def ew_scalar_mul(A, s, N):  # mul_ops.jl:33-40
    return np.array((np.asarray(A).astype(object) * int(s)) % N, dtype=np.int64)


def ew_scalar_div(A, s, N):  # div_ops.jl:13-22: multiply by mod_inv(s, N)
    return ew_scalar_mul(A, mod_inv(int(s) % N, N), N)


def ew_negate(A, N):  # CuModMatrix negate! == rscalar_sub(0)
    return np.mod(-np.asarray(A, np.int64), N)


# --------------------------------------------------------------------------------------
# Permutations: rref_lu_pluq/permutations.jl:49-62 (cols), :112-125 (rows)
# transposition lists are 1-based tuples applied in order; *_inv = reversed list.
# --------------------------------------------------------------------------------------
def apply_col_perm(P, A, inverse=False):
    A = np.array(A, copy=True)
    for (c1, c2) in (reversed(P) if inverse else P):
        A[:, [c1 - 1, c2 - 1]] = A[:, [c2 - 1, c1 - 1]]
    return A


def apply_row_perm(P, A, inverse=False):
    A = np.array(A, copy=True)
    for (r1, r2) in (reversed(P) if inverse else P):
        A[[r1 - 1, r2 - 1], :] = A[[r2 - 1, r1 - 1], :]
    return A


def perm_array_to_matrix(perm, n=None, perm_stack=False):
    """permutations.jl:141-157."""
    if perm_stack:
        n = n if n is not None else len(perm)
        P = np.eye(n, dtype=np.int64)
        for (i, j) in perm:
            P[[i - 1, j - 1], :] = P[[j - 1, i - 1], :]
        return P
    n = len(perm)
    P = np.zeros((n, n), dtype=np.int64)
    for i in range(n):
        P[perm[i] - 1, i] = 1
    return P


# --------------------------------------------------------------------------------------
# PLUQ: rref_lu_pluq/pluq_kernels.jl:46-157 (host loop), :179-202 (find_pivot),
#       :307-320 (swap+scale), :385-393 (move col), :411-440 (rank-1 update)
# --------------------------------------------------------------------------------------
def _elim_dtype(N: int, python_ints: bool):
    """integer type in which x + m*y (x, m, y < N) is exact: int64 below 2^31, uint64 up to 2^32 ((N-1)*N < 2^64), python integers above"""
    if python_ints or N > 2 ** 32:
        return object
    return np.int64 if N < 2 ** 31 else np.uint64


def pluq_reference(A: np.ndarray, N: int, python_ints: bool = False):
    """Literal restatement of pluq_gpu_kernel, INCLUDING its rank-deficient quirk
    (an all-zero pivot column is swapped with the fixed last column `cols` and the swapped-in
    column is then skipped: pluq_kernels.jl:88,103,193).  Returns (U, L, Perm_rows, Perm_cols)
    with 1-based transposition tuples.  Pivot = maximum residue at/below `row`, first index on
    ties (findmax, :189).  L is written at column `col` (:343,389), as in the reference.
    """
    dt = _elim_dtype(N, python_ints)
    dA = np.mod(np.asarray(A, dtype=np.int64), N).astype(dt)
    rows, cols = dA.shape
    ldim = max(rows, cols)
    dL = np.zeros((rows, ldim), dtype=dt)  # reference allocates prow x prow (padded); we keep enough cols
    perm_rows, perm_cols = [], []
    row = col = 0
    last = cols - 1  # Perm_col_idx = cols, never changes (:88)
    while row < rows and col < cols:
        piv_val, piv_idx = -1, -1
        while True:
            colv = dA[row:, col]
            idx = int(np.argmax(colv))  # first maximal index
            val = int(colv[idx])
            if val == 0:
                dA[:, [col, last]] = dA[:, [last, col]]  # swap_cols over all rows (:217-226)
                perm_cols.append((col + 1, last + 1))
                col += 1  # swapped-in column is skipped (:103)
                if col >= cols:
                    break
            else:
                piv_val, piv_idx = val, idx
                break
        if col >= cols:
            break
        prow = row + piv_idx
        pinv = mod_inv(piv_val, N)
        # swap_rows_and_mod (:307-320): whole rows over 1..cols, new pivot row scaled by p^-1
        tmp = dA[prow, :].copy()
        dA[prow, :] = dA[row, :]
        dA[row, :] = (tmp * pinv) % N
        # swap_rows on d_L (:280-289)
        dL[[row, prow], :] = dL[[prow, row], :]
        if row != prow:
            perm_rows.append((row + 1, prow + 1))
        # move_and_zero_out (:340-351,366-393)
        dL[row, col] = piv_val
        dL[row + 1:, col] = dA[row + 1:, col]
        dA[row + 1:, col] = 0
        # update_sub_matrix_kernel (:411-440): A[r,c] = mod(A[r,c] + (N - L[r,col]) * A[row,c], N)
        if row + 1 < rows and col + 1 < cols:
            mult = (N - dL[row + 1:, col]) % N
            dA[row + 1:, col + 1:] = (dA[row + 1:, col + 1:] + np.outer(mult, dA[row, col + 1:])) % N
        row += 1
        col += 1
    U = np.array(dA, dtype=np.int64)
    L = np.array(dL[:, :rows], dtype=np.int64)
    return U, L, perm_rows, perm_cols


def echelon(A: np.ndarray, N: int, python_ints: bool = False):
    """Well-defined rank-revealing elimination used by the new `pluq(correct)`, `lu`, `rref`:
    same pivot rule (max residue, first index; pluq_kernels.jl:189), same scaling conventions
    (unit pivots in U, pivot values on diag(L), un-normalised sub-column in L; :314,:343,:389),
    but a column without pivot is simply skipped (no column swap) and L's column index is the
    pivot NUMBER.  Returns (E, L, Perm_rows, pivcols): E = row echelon form with unit pivots
    (rows >= rank are zero), L rows x rows lower-triangular (columns >= rank zero),
    P*A = L*E with P the product of the transpositions.
    """
    key = (int(N), np.asarray(A).shape, hash(np.ascontiguousarray(np.asarray(A, dtype=np.int64)).tobytes()))
    if not python_ints and _ECHELON_CACHE.get("key") == key:  # pluq / lu / rref / inverse of the same input: one elimination (the callers get copies)
        E_, L_, pr_, pc_ = _ECHELON_CACHE["val"]
        return E_.copy(), L_.copy(), list(pr_), list(pc_)
    dt = _elim_dtype(N, python_ints)
    E = np.mod(np.asarray(A, dtype=np.int64), N).astype(dt)
    rows, cols = E.shape
    L = np.zeros((rows, rows), dtype=dt)
    perm_rows, pivcols = [], []
    row = 0
    for col in range(cols):
        if row >= rows:
            break
        colv = E[row:, col]
        idx = int(np.argmax(colv))
        val = int(colv[idx])
        if val == 0:
            continue
        prow = row + idx
        pinv = mod_inv(val, N)
        tmp = E[prow, :].copy()
        E[prow, :] = E[row, :]
        E[row, :] = (tmp * pinv) % N
        L[[row, prow], :] = L[[prow, row], :]
        if row != prow:
            perm_rows.append((row + 1, prow + 1))
        L[row, row] = val
        L[row + 1:, row] = E[row + 1:, col]
        E[row + 1:, col] = 0
        if row + 1 < rows and col + 1 < cols:
            mult = (N - L[row + 1:, row]) % N
            E[row + 1:, col + 1:] = (E[row + 1:, col + 1:] + np.outer(mult, E[row, col + 1:])) % N
        pivcols.append(col)
        row += 1
    out = (np.array(E, dtype=np.int64), np.array(L, dtype=np.int64), perm_rows, pivcols)
    _ECHELON_CACHE["key"], _ECHELON_CACHE["val"] = key, (out[0].copy(), out[1].copy(), list(perm_rows), list(pivcols))
    return out


_ECHELON_CACHE = {}


def pivcols_to_perm(pivcols, cols):
    """Column order that moves pivot columns to the front (stable), and the equivalent ordered
    1-based transposition list in the reference's (col, other) tuple format."""
    piv = list(pivcols)
    rest = [c for c in range(cols) if c not in set(piv)]
    order = piv + rest  # new column j holds old column order[j]
    cur = list(range(cols))
    pos = {c: c for c in range(cols)}
    swaps = []
    for j in range(cols):
        want = order[j]
        pj = pos[want]
        if pj != j:
            swaps.append((j + 1, pj + 1))
            cj = cur[j]
            cur[j], cur[pj] = want, cj
            pos[want], pos[cj] = j, pj
    return order, swaps


def pluq(A: np.ndarray, N: int):
    """`pluq` in `correct` mode: P*A*Q = L*U, U = [unit-upper-trapezoidal; 0] with the pivot columns
    moved to the front.  For full-column-rank leading columns (e.g. any invertible square matrix)
    this coincides exactly with pluq_reference (no column swaps occur there)."""
    E, L, perm_rows, pivcols = echelon(A, N)
    order, perm_cols = pivcols_to_perm(pivcols, E.shape[1])
    U = E[:, order]
    return U, L, perm_rows, perm_cols, len(pivcols)


def lu(A: np.ndarray, N: int):
    """`lu(A) -> (U, L, Perm)` after the intended lu_gpu_type signature
    (test/Experiments/rref_gpu_type.jl:60-103): row pivoting only; U is the echelon form."""
    E, L, perm_rows, pivcols = echelon(A, N)
    return E, L, perm_rows


def rref(A: np.ndarray, N: int):
    """Unique reduced row echelon form (intended rref_gpu_type, rref_gpu_type.jl:8-51)."""
    E, _, _, pivcols = echelon(A, N)
    R = E.astype(_elim_dtype(N, False))
    for t in range(len(pivcols) - 1, -1, -1):
        c = pivcols[t]
        if t > 0:
            f = (N - R[:t, c]) % N  # x - f*y == x + (N - f)*y (mod N): no negative intermediate, exact in the unsigned type too
            R[:t, :] = (R[:t, :] + np.outer(f, R[t, :])) % N
    return np.array(R, dtype=np.int64), pivcols


def rank(A, N):
    return len(echelon(A, N)[3])


# --------------------------------------------------------------------------------------
# Triangular inverse: triangular/triangular_inverse_no_copy.jl:197-228 (upper), :450-478 (lower)
# base case triangular/substitution_inplace.jl:6-30, :63-89
# --------------------------------------------------------------------------------------
def upper_triangular_inverse(A: np.ndarray, N: int) -> np.ndarray:
    """Inverse of the leading square block of an upper-triangular (possibly wide) matrix.
    For a wide rows x cols input the reference returns [T^-1; 0] of shape cols x rows
    (triangular_inverse_no_copy.jl:197-228) so that A * A_inv == I_rows."""
    A = np.mod(np.asarray(A, dtype=np.int64), N)
    rows, cols = A.shape
    n = rows
    T = A[:n, :n].astype(object)
    X = np.zeros((n, n), dtype=object)
    for j in range(n):  # backward substitution per inverse column (substitution_inplace.jl:63-89)
        for i in range(j, -1, -1):
            s = (1 if i == j else 0) - sum(int(T[i, k]) * int(X[k, j]) for k in range(i + 1, j + 1))
            X[i, j] = (s * mod_inv(int(T[i, i]) % N, N)) % N
    out = np.zeros((cols, rows), dtype=np.int64)
    out[:n, :n] = np.array(X, dtype=np.int64)
    return out


def lower_triangular_inverse(A: np.ndarray, N: int) -> np.ndarray:
    """Square lower-triangular inverse (forward substitution, substitution_inplace.jl:6-30).
    A tall input raises like InverseNotDefinedException (triangular_inverse_no_copy.jl:463)."""
    A = np.mod(np.asarray(A, dtype=np.int64), N)
    rows, cols = A.shape
    if rows > cols:
        raise ValueError("InverseNotDefinedException")
    return upper_triangular_inverse(A[:rows, :rows].T, N)[:rows, :rows].T.copy()


def _fast_tri_inverse_upper(T: np.ndarray, N: int, python_ints: bool = False) -> np.ndarray:
    """Vectorised unit/non-unit upper-triangular inverse (same result as above, O(n) numpy steps).  N < 2^31: int64 rows with the
    row-times-matrix product through exact_matmul_mod (exact K-chunks); larger N: python integers."""
    n = T.shape[0]
    if N < 2 ** 31 and not python_ints:
        T = np.mod(np.asarray(T, dtype=np.int64), N)
        X = np.zeros((n, n), dtype=np.int64)
        for i in range(n - 1, -1, -1):
            rhs = (N - exact_matmul_mod(T[i, i + 1:].reshape(1, -1), X[i + 1:, :], N).reshape(-1)) % N if i + 1 < n else np.zeros(n, dtype=np.int64)
            rhs[i] = (rhs[i] + 1) % N
            X[i, :] = (rhs * mod_inv(int(T[i, i]), N)) % N  # < N * N < 2^62
        return X
    if N <= 2 ** 32 and not python_ints:  # 2^31 <= N <= 2^32: uint64, products reduced one by one, sums of n residues < 2^63
        T = np.mod(np.asarray(T, dtype=np.int64), N).astype(np.uint64)
        X = np.zeros((n, n), dtype=np.uint64)
        for i in range(n - 1, -1, -1):
            dot = ((T[i, i + 1:, None] * X[i + 1:, :]) % N).sum(axis=0, dtype=np.uint64) % N if i + 1 < n else np.zeros(n, dtype=np.uint64)
            rhs = (N - dot) % N
            rhs[i] = (rhs[i] + 1) % N
            X[i, :] = (rhs * mod_inv(int(T[i, i]), N)) % N
        return X.astype(np.int64)
    T = np.mod(np.asarray(T, dtype=np.int64), N).astype(object)
    X = np.zeros((n, n), dtype=object)
    dinv = [mod_inv(int(T[i, i]), N) for i in range(n)]
    for i in range(n - 1, -1, -1):
        rhs = -(T[i, i + 1:].reshape(1, -1) @ X[i + 1:, :]).reshape(-1) if i + 1 < n else np.zeros(n, dtype=object)
        rhs[i] += 1
        X[i, :] = (rhs * dinv[i]) % N
    return np.array(X, dtype=np.int64)


# --------------------------------------------------------------------------------------
# inverse: CuModMatrix.jl:480-502 (inverse), :356-422 (is_invertible_with_inverse)
# --------------------------------------------------------------------------------------
def is_invertible_with_inverse(A: np.ndarray, N: int):
    A = np.mod(np.asarray(A, dtype=np.int64), N)
    rows, cols = A.shape
    if rows != cols:
        return False, None
    U, L, perm_rows, perm_cols = pluq_reference(A, N)
    if any(int(L[i, i]) == 0 for i in range(rows)) or len(perm_cols) > 0:
        # reference: rank(Array(U)) != min(size) -> (false, nothing)  (CuModMatrix.jl:340-347,368-370)
        return False, None
    Uinv = _fast_tri_inverse_upper(U, N)
    Linv = _fast_tri_inverse_upper(L.T, N).T
    Linv = apply_col_perm(perm_rows, Linv, inverse=True)  # apply_col_inv_perm!(P, L_inv)  (:494)
    Uinv = apply_row_perm(perm_cols, Uinv, inverse=True)  # apply_row_inv_perm!(Q, U_inv)  (:495)
    return True, exact_matmul_mod(Uinv, Linv, N)          # U_inv * L_inv                  (:496)


def inverse(A, N):
    ok, inv = is_invertible_with_inverse(A, N)
    if not ok:
        raise ValueError("MatrixNotInvertibleException")
    return inv


# --------------------------------------------------------------------------------------
# Karatsuba two-limb matrices: KaratsubaMatrix/KaratsubaMatrix.jl:133-204, :318-336, :372-397
#                              KaratsubaMatrix/KaratsubaKernels.jl:129-158
# --------------------------------------------------------------------------------------
def karatsuba_split(A: np.ndarray, N1: int, N2: int):
    """KaratsubaMatrix(T,A,N1,N2,M): data1 = A mod N1, data2 = A div N1 (KaratsubaMatrix.jl:372-397)."""
    A = np.mod(np.asarray(A).astype(object), N1 * N2)
    return np.array(A % N1, dtype=np.int64), np.array(A // N1, dtype=np.int64)


def karatsuba_join(d1, d2, N1):
    """Array(K) = data1 + N1*data2 (KaratsubaMatrix.jl:318-336)."""
    return np.asarray(d1).astype(object) + N1 * np.asarray(d2).astype(object)


def karatsuba_matmul(A1, A2, B1, B2, N1: int, N2: int):
    """KMatMul!: three sub-products + recombination, exactly as the reference's kernels.
    kernel_1 (KaratsubaKernels.jl:129-139): plan = (d1+d2) % (2*N1)
    products  (KaratsubaMatrix.jl:181-184): P1 = A1*B1 mod N1^2, P2 = Ap*Bp mod (4N1)^2, P3 = A2*B2 mod N1^2
    kernel_2 (KaratsubaKernels.jl:141-158): carry / recombine (uses truncated % for the differences).
    Valid iff N2 divides N1."""
    A1 = np.asarray(A1, dtype=np.int64); A2 = np.asarray(A2, dtype=np.int64)
    B1 = np.asarray(B1, dtype=np.int64); B2 = np.asarray(B2, dtype=np.int64)
    if B1.ndim == 1:
        B1 = B1[:, None]; B2 = B2[:, None]
    Ap = (A1 + A2) % (2 * N1)
    Bp = (B1 + B2) % (2 * N1)
    P1 = exact_matmul_mod(A1, B1, N1 * N1).astype(object)
    P2 = exact_matmul_mod(Ap, Bp, (4 * N1) ** 2).astype(object)
    P3 = exact_matmul_mod(A2, B2, N1 * N1).astype(object)
    q = (4 * N1) ** 2
    cc1 = P1 % (N1 * N1)
    cc2 = P2 % q
    bp = P3 % N1
    cp = cc1 // N1
    trunc_rem = np.frompyfunc(lambda a, b: int(np.sign(a)) * (abs(a) % b), 2, 1)  # Julia % (rem)
    cc2 = trunc_rem(cc2 - cc1, q)
    cc2 = trunc_rem(cc2 - bp, q)
    cc2 = cc2 + cp
    C1 = np.array(cc1 % N1, dtype=np.int64)
    C2 = np.array(cc2 % N2, dtype=np.int64)
    return C1, C2


def karatsuba_matmul_direct(A1, A2, B1, B2, N1, N2):
    """Ground truth the reference's Karatsuba test uses (basic_operations_test.jl:127):
    mod.(A_full*B_full, N1*N2) split back into limbs."""
    Af = karatsuba_join(A1, A2, N1)
    Bf = karatsuba_join(B1, B2, N1)
    if Bf.ndim == 1:
        Bf = Bf[:, None]
    C = (Af @ Bf) % (N1 * N2)
    return np.array(C % N1, dtype=np.int64), np.array(C // N1, dtype=np.int64)


# ----------------------------------------------------------------------------------------------------------
# Hensel lifting of an inverse -- src/CuModMatrix/triangular/hensel.jl:13-21 (not loaded by the reference package:
# it needs Nemo for the final residue-ring matrix; the arithmetic is the three-line Newton loop below)
# ----------------------------------------------------------------------------------------------------------
def hensel_pseudoinverse(N: int, precision: int, A, T):
    """T with A*T == I (mod N)  ->  T with A*T == I (mod N^precision); exact python integers (object arrays)."""
    M = N ** precision
    A = np.array(A, dtype=object) % M
    T = np.array(T, dtype=object) % M
    i = 1
    while i < precision:  # hensel.jl:15-18
        T = (2 * T - T.dot(A.dot(T) % M)) % M
        i *= 2
    return T
