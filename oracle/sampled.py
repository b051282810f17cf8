"""Sampled-row parity checks at sizes where the full CPU product is out of reach -- TEST INFRASTRUCTURE ONLY.

The reference's tests compare the GPU result with `mod.(A*B, N)` computed on the host with `==`
(/root/reference/test/CuModMatrix/stripe_mul_test.jl:31-50).  At n = 16384 / 32768 a full host product takes hours, but any
subset of ROWS of C = A*B depends only on the same rows of A and on all of B: `rows` x n x n mul-adds (seconds for 64 rows with
oracle_c.matmul_mod).  These helpers draw the rows, rebuild them with the oracle's own generator (so the device generator is
checked too) and return the oracle's rows for a bit-exact comparison.  Used by tests/ and by bench.py's `parity_check`.
"""
from __future__ import annotations

import numpy as np

from . import oracle as O
from . import oracle_c as OC


def pick_rows(m: int, count: int, seed: int) -> np.ndarray:
    """`count` distinct row indices of an m-row matrix (sorted), always including the first and the last row."""
    if m <= count or m <= 2:
        return np.arange(m, dtype=np.int64)
    rng = np.random.default_rng(seed)
    inner = rng.choice(m - 2, size=max(count - 2, 0), replace=False) + 1
    return np.sort(np.concatenate([np.array([0, m - 1]), inner])).astype(np.int64)


def synth_rows(seed: int, rows_idx, rows: int, cols: int, N: int) -> np.ndarray:
    """Rows `rows_idx` of O.synth_matrix(seed, rows, cols, N) without building the matrix (same counter-based generator)."""
    i = np.asarray(rows_idx, dtype=np.uint64)[:, None]
    j = np.arange(cols, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        idx = j * np.uint64(rows) + i
    return (O.splitmix64(np.uint64(seed) ^ idx) % np.uint64(N)).astype(np.int64)


def synth_cols(seed: int, cols_idx, rows: int, N: int) -> np.ndarray:
    """Columns `cols_idx` of O.synth_matrix(seed, rows, cols, N), shape (rows, len(cols_idx))."""
    i = np.arange(rows, dtype=np.uint64)[:, None]
    j = np.asarray(cols_idx, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        idx = j * np.uint64(rows) + i
    return (O.splitmix64(np.uint64(seed) ^ idx) % np.uint64(N)).astype(np.int64)


def product_rows(A_rows: np.ndarray, B: np.ndarray, N: int, in_bound: int = 0) -> np.ndarray:
    """Rows of (A*B) mod N from the same rows of A (r x k) and all of B (k x n): the C oracle's exact uint64 product."""
    return OC.matmul_mod(np.asarray(A_rows), B, N, in_bound=in_bound)


def check_product_rows(C_rows: np.ndarray, A_rows: np.ndarray, B: np.ndarray, N: int, in_bound: int = 0) -> dict:
    """Bit-exact comparison of downloaded rows of C with the oracle's rows; returns a small report."""
    want = product_rows(A_rows, B, N, in_bound)
    got = np.asarray(C_rows).astype(np.int64)
    bad = int(np.count_nonzero(want != got))
    return {"rows": int(want.shape[0]), "cols": int(want.shape[1]), "mismatches": bad, "match": bad == 0}


def perm_to_map(pairs, n: int) -> np.ndarray:
    """Index map of an ordered 1-based transposition list (permutations.jl:49-62, :112-125): applying the list to the rows of A gives
    A[map] (and to the columns A[:, map]) -- O.apply_row_perm / O.apply_col_perm without touching the matrix n times."""
    m = np.arange(n, dtype=np.int64)
    for (a, b) in pairs:
        m[a - 1], m[b - 1] = m[b - 1], m[a - 1]
    return m
