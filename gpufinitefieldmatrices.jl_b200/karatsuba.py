"""Host-side mirror of `KaratsubaMatrix` (reference src/KaratsubaMatrix/KaratsubaMatrix.jl): two-limb values
x = data1 + N1*data2 modulo M = N1*N2.  All arithmetic runs in libgffm.so (gffm_kmat_mul / gffm_kmat_ewise)."""
from __future__ import annotations

import numpy as np

from . import capi
from .cumodmatrix import CuModMatrix, DEFAULT_TYPE


class KaratsubaMatrix:
    """struct KaratsubaArray (KaratsubaMatrix.jl:1-52): data1 < N1, data2 < N2, plan = scratch of the same shape."""

    def __init__(self, data1: CuModMatrix, data2: CuModMatrix, N1: int, N2: int, M: int = None):
        M = N1 * N2 if M is None else M
        if M != N1 * N2:
            raise ValueError("M must equal N1*N2")
        if (data1.rows, data1.cols) != (data2.rows, data2.cols):
            raise capi.CuModArraySizeMismatchException(capi.ERR_SIZE_MISMATCH, "limb shapes differ")
        self.data1, self.data2 = data1, data2
        self.N1, self.N2, self.M = int(N1), int(N2), int(M)
        self.plan = None

    @classmethod
    def from_array(cls, A, N1, N2, M=None, elem_type=np.float64, ctx=None):
        """`KaratsubaMatrix(T, A, N1, N2, M)` split (KaratsubaMatrix.jl:372-397): data1 = A mod N1, data2 = A div N1."""
        A = np.asarray(A).astype(object) % (N1 * N2)
        d1 = np.array(A % N1, dtype=np.int64)
        d2 = np.array(A // N1, dtype=np.int64)
        return cls(CuModMatrix(d1, N1, elem_type=elem_type, ctx=ctx), CuModMatrix(d2, N1, elem_type=elem_type, ctx=ctx), N1, N2, M)

    @property
    def shape(self):
        return self.data1.shape

    def Array(self):
        """`Array(K)` = data1 + N1*data2 (KaratsubaMatrix.jl:318-336) as python ints (exact up to 2^52)."""
        return self.data1.Array(np.int64).astype(object) + self.N1 * self.data2.Array(np.int64).astype(object)


KaratsubaVector = KaratsubaMatrix
MatToKMat = KaratsubaMatrix.from_array


def KaratsubaZeros(T, rows, cols, N1, N2, M=None, ctx=None):
    """KaratsubaMatrix.jl:404-420."""
    from .cumodmatrix import zeros
    return KaratsubaMatrix(zeros(T, rows, cols, N1, ctx=ctx), zeros(T, rows, cols, N1, ctx=ctx), N1, N2, M)


def initialize_plan_(K: KaratsubaMatrix):
    """`initialize_plan!` (KaratsubaMatrix.jl:422-424).  The B200 build fuses the limb add into the GEMM prologue,
    so no plan buffer is needed; kept for API compatibility."""
    K.plan = True
    return K


def _same(C, A, B=None):
    ks = [C, A] + ([B] if B is not None else [])
    if len({(k.N1, k.N2) for k in ks}) != 1:
        raise capi.CuModArrayModulusMismatchException(capi.ERR_MODULUS_MISMATCH, "Karatsuba operands have different moduli")


def KMatMul_(C: KaratsubaMatrix, A: KaratsubaMatrix, B: KaratsubaMatrix):
    """`KMatMul!(C, A, B)` (KaratsubaMatrix.jl:133-204); B may be a Karatsuba vector (n x 1)."""
    _same(C, A, B)
    lib = A.data1.lib
    capi.check(lib.gffm_kmat_mul(C.data1.h, C.data2.h, A.data1.h, A.data2.h, B.data1.h, B.data2.h, A.N1, A.N2))
    return C


KMatMul_gemv_ = KMatMul_  # KaratsubaMatrix.jl:238-300


def _kew(op, C, A, B=None, scalar=0):
    lib = A.data1.lib
    capi.check(lib.gffm_kmat_ewise(op, C.data1.h, C.data2.h, A.data1.h, A.data2.h, B.data1.h if B else None,
                                   B.data2.h if B else None, int(scalar), A.N1, A.N2))
    return C


def add_(C, A, B):
    """KaratsubaMatrix.jl:505-536."""
    _same(C, A, B)
    return _kew(capi.EW_ADD, C, A, B)


def sub_(C, A, B):
    """KaratsubaMatrix.jl:583-629."""
    _same(C, A, B)
    return _kew(capi.EW_SUB, C, A, B)


def scalar_multiply_(C, A, s):
    """KaratsubaMatrix.jl:631-666."""
    _same(C, A)
    return _kew(capi.EW_SMUL, C, A, scalar=s)


def negate_(C, A):
    """KaratsubaMatrix.jl:691-731."""
    _same(C, A)
    return _kew(capi.EW_RSSUB, C, A)


def hensel_pseudoinverse(precision_steps: int, A: KaratsubaMatrix, T: KaratsubaMatrix) -> KaratsubaMatrix:
    """Newton/Hensel lift `T = 2*T - T*(A*T)` (src/CuModMatrix/triangular/hensel.jl:15-18) on two-limb matrices modulo
    M = N1*N2 (up to 2^52): each of the `precision_steps` iterations doubles the p-adic precision of A*T = I.  T is
    updated in place and returned."""
    _same(T, A)
    from .cumodmatrix import zeros
    r, c = A.data1.rows, A.data1.cols
    ctx = A.data1.ctx
    W = KaratsubaZeros(A.data1.elem_type, r, c, A.N1, A.N2, ctx=ctx)
    V = KaratsubaZeros(A.data1.elem_type, r, c, A.N1, A.N2, ctx=ctx)
    for _ in range(int(precision_steps)):
        KMatMul_(W, A, T)
        KMatMul_(V, T, W)
        scalar_multiply_(W, T, 2)
        sub_(T, W, V)
    return T
