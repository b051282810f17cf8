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

    def size(self):
        """`Base.size(A)` (KaratsubaMatrix.jl:302)."""
        return self.data1.size()

    def __getitem__(self, idx):
        """`getindex` (KaratsubaMatrix.jl:304-305): data1[i,j] + N1*data2[i,j] (0-based here)."""
        return int(self.data1[idx]) + self.N1 * int(self.data2[idx])

    def __setitem__(self, idx, v):
        """`setindex!` (KaratsubaMatrix.jl:307-316): data1 = rem(v, N1), data2 = div(v, N1)."""
        v = int(v)
        self.data1[idx] = v % self.N1
        self.data2[idx] = v // self.N1

    # operators (KaratsubaMatrix.jl:428-503)
    def _like(self, rows=None, cols=None):
        from .cumodmatrix import zeros
        r = self.data1.rows if rows is None else rows
        c = self.data1.cols if cols is None else cols
        T, ctx = self.data1.elem_type, self.data1.ctx
        return KaratsubaMatrix(zeros(T, r, c, self.N1, ctx=ctx), zeros(T, r, c, self.N1, ctx=ctx), self.N1, self.N2, self.M)

    def __add__(self, o):
        return add_(self._like(), self, o)

    def __sub__(self, o):
        return sub_(self._like(), self, o)

    def __neg__(self):
        return negate_(self._like(), self)

    def __mul__(self, o):
        if isinstance(o, KaratsubaMatrix):  # `*(A,B)`: Karatsuba product (left unfinished in the reference, :467-503)
            return KMatMul_(self._like(cols=o.data1.cols), self, o)
        return scalar_multiply_(self._like(), self, o)

    __rmul__ = __mul__
    __matmul__ = __mul__

    def Array(self):
        """`Array(K)` = data1 + N1*data2 (KaratsubaMatrix.jl:318-336) as python ints (exact up to 2^52)."""
        return self.data1.Array(np.int64).astype(object) + self.N1 * self.data2.Array(np.int64).astype(object)


KaratsubaVector = KaratsubaMatrix
KaratsubaArray = KaratsubaMatrix


def MatToKMat(A, N1=None, N2=None, M=None, elem_type=np.float64, ctx=None):
    """`MatToKMat` (KaratsubaMatrix.jl:358-401): with one modulus M the limbs are both taken modulo M
    (`KaratsubaMatrix(T, A, M, M, M)`, :367-370); with N1, N2 it is the split constructor (:372-397)."""
    if N1 is None:
        raise TypeError("MatToKMat needs a modulus (the reference's modulus-free form derives one from find_max_ops)")
    if N2 is None:
        N2 = N1
    return KaratsubaMatrix.from_array(A, N1, N2, M, elem_type=elem_type, ctx=ctx)


def KMatToMat(K: "KaratsubaMatrix"):
    """`KMatToMat` (KaratsubaMatrix.jl:352-356): data1 + N1*data2 on the host (exact integers)."""
    return K.Array()


def copy_(B: "KaratsubaMatrix", A: "KaratsubaMatrix"):
    """`Base.copy!(B, A)` (KaratsubaMatrix.jl:338-345)."""
    from .cumodmatrix import copy_ as _mcopy
    _mcopy(B.data1, A.data1)
    _mcopy(B.data2, A.data2)
    if A.plan is not None and B.plan is None:
        initialize_plan_(B)
    return B


def zero_(A: "KaratsubaMatrix"):
    """`zero!` (KaratsubaMatrix.jl:347-350)."""
    from .cumodmatrix import zero_ as _mzero
    _mzero(A.data1)
    _mzero(A.data2)
    return A


def Karatsubacopy(A: "KaratsubaMatrix"):
    """`Karatsubacopy` (KaratsubaMatrix.jl:738-744)."""
    from .cumodmatrix import copy as _mcopy
    B = KaratsubaMatrix(_mcopy(A.data1), _mcopy(A.data2), A.N1, A.N2, A.N1 * A.N2)
    if A.plan is not None:
        initialize_plan_(B)
    return B


def KaratsubaZeros(T, rows, cols, N1, N2, M=None, use_gpu=True, ctx=None):
    """KaratsubaMatrix.jl:404-420 (`use_gpu` is accepted for signature parity; there is no CPU variant here)."""
    if not use_gpu:
        raise ValueError("KaratsubaZeros(use_gpu=false): this build has no CPU path")
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
