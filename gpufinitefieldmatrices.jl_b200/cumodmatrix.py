"""Host-side mirror of the reference's `CuModMatrix` interface (src/CuModMatrix/CuModMatrix.jl and friends),
written against the C ABI exactly like the Julia shim (julia/GPUFiniteFieldMatricesB200.jl).  Julia's `f!` is
spelled `f_` here; everything else keeps the reference's names, argument meaning and error behaviour so that the
parity tests read like the reference's own tests.  Indices in this Python mirror are 0-based; permutation tuples
keep the reference's 1-based convention.

All computation happens in libgffm.so on the GPU; numpy is only the host container for upload/download.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import capi

TILE_WIDTH = 32           # CuModMatrix.jl:2
DEFAULT_TYPE = np.float32  # CuModMatrix.jl:3

_DT = {np.dtype(np.float32): capi.F32, np.dtype(np.float64): capi.F64, np.dtype(np.int64): capi.I64,
       np.dtype(np.uint32): capi.U32, np.dtype(np.int32): capi.I32}


class Context:
    """One per thread/task (CUDA.jl task-local state in the reference).  Owns the stream and workspaces."""

    def __init__(self, device: int = 0):
        self.lib = capi.load()
        h = C.c_void_p()
        capi.check(self.lib.gffm_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def sync(self):
        capi.check(self.lib.gffm_sync(self.h))

    def set_stream(self, cuda_stream_ptr: int):
        capi.check(self.lib.gffm_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def get_stream(self) -> int:
        """Raw cudaStream_t of the context (e.g. for torch.cuda.ExternalStream)."""
        p = C.c_void_p()
        capi.check(self.lib.gffm_get_stream(self.h, C.byref(p)))
        return p.value or 0

    def set_profiling(self, on: bool):
        capi.check(self.lib.gffm_set_profiling(self.h, 1 if on else 0))

    def last_timings(self):
        buf = (C.c_double * 16)()
        n = C.c_int32(0)
        capi.check(self.lib.gffm_last_timings(self.h, buf, 16, C.byref(n)))
        return [buf[i] for i in range(n.value)]

    def set_gemm_ctas(self, ctas: int):
        """Cap the persistent GEMM grid (0 = one CTA per SM) so that concurrent NCCL kernels find free SMs."""
        capi.check(self.lib.gffm_set_gemm_ctas(self.h, int(ctas)))

    def launch_count(self) -> int:
        n = C.c_int64(0)
        capi.check(self.lib.gffm_launch_count(self.h, C.byref(n)))
        return n.value

    def alloc_stats(self):
        """(bytes, calls): device memory requested by the library on this context since creation -- the counterpart of
        CUDA.@timed's gpu_bytes in the reference's allocation tests (test/CuModMatrix/allocations_test.jl:21-52)"""
        b, c = C.c_int64(0), C.c_int64(0)
        capi.check(self.lib.gffm_alloc_stats(self.h, C.byref(b), C.byref(c)))
        return b.value, c.value

    def modinv_batch(self, values, N: int):
        v = np.ascontiguousarray(values, dtype=np.uint64)
        out = np.zeros_like(v)
        capi.check(self.lib.gffm_modinv_batch(self.h, v.ctypes.data_as(C.POINTER(C.c_uint64)),
                                              out.ctypes.data_as(C.POINTER(C.c_uint64)), v.size, N))
        return out

    def close(self):
        if self.h:
            self.lib.gffm_destroy(self.h)
            self.h = None


_tls = threading.local()


_default_device = 0


def set_default_device(device: int):
    """Device of the context that `default_context()` (and every constructor called without `ctx=`) uses -- one rank per GPU sets it to
    its LOCAL_RANK once (CUDA.device!(...) in the reference's host runtime)."""
    global _default_device
    _default_device = int(device)


def default_context(device: int = None) -> Context:
    if device is None:
        device = _default_device
    ctxs = getattr(_tls, "ctxs", None)
    if ctxs is None:
        ctxs = _tls.ctxs = {}
    if device not in ctxs:
        ctxs[device] = Context(device)
    return ctxs[device]


def _host_fortran(A, dtype=None):
    A = np.asarray(A)
    if A.ndim == 1:
        A = A.reshape(-1, 1)
    if A.ndim != 2:
        raise ValueError("matrix or vector expected")
    if dtype is not None:
        A = A.astype(dtype, copy=False)
    elif A.dtype not in _DT:
        if np.issubdtype(A.dtype, np.integer) or A.dtype == bool:
            A = A.astype(np.int64)
        else:
            A = A.astype(np.float64)
    return np.asfortranarray(A)


class CuModMatrix:
    """`CuModMatrix(A, N; mod=true, new_size=nothing, elem_type=Float32)` (CuModMatrix.jl:53-99,130-181).

    Storage is uint32 residues on the device with the reference's +32 zero padding; `elem_type` only selects the
    host element type `Array()` returns (the reference stores float-encoded integers)."""

    def __init__(self, A=None, N: int = None, *, mod: bool = True, new_size=None, elem_type=DEFAULT_TYPE, ctx: Context = None,
                 _handle=None, _vector=False):
        self.ctx = ctx or default_context()
        self.lib = self.ctx.lib
        self.elem_type = np.dtype(elem_type)
        self._vector = _vector
        if _handle is not None:
            self.h = _handle
            return
        if N is None:
            raise TypeError("modulus N is required")
        if A is None:
            raise TypeError("A is required")
        is_vec = np.asarray(A).ndim == 1
        self._vector = is_vec
        Ah = _host_fortran(A)
        rows, cols = Ah.shape
        if new_size is not None:
            rows, cols = (int(new_size[0]), int(new_size[1])) if len(new_size) == 2 else (int(new_size[0]), 1)
        h = C.c_void_p()
        capi.check(self.lib.gffm_mat_create(self.ctx.h, rows, cols, int(N), -1, C.byref(h)))
        self.h = h
        sub = Ah[: min(rows, Ah.shape[0]), : min(cols, Ah.shape[1])]
        if sub.size:
            if (sub.shape[0], sub.shape[1]) != (rows, cols):
                # new_size larger than the data: upload the data into the top-left corner
                tmp = np.zeros((rows, cols), dtype=Ah.dtype, order="F")
                tmp[: sub.shape[0], : sub.shape[1]] = sub
                sub = tmp
            sub = np.asfortranarray(sub)
            capi.check(self.lib.gffm_mat_upload(self.h, sub.ctypes.data_as(C.c_void_p), _DT[sub.dtype], sub.shape[0], 1 if mod else 0))

    # ---- construction helpers ----
    @classmethod
    def _new(cls, rows, cols, N, like: "CuModMatrix" = None, ctx=None, elem_type=None, vector=False):
        ctx = ctx or (like.ctx if like is not None else default_context())
        h = C.c_void_p()
        capi.check(ctx.lib.gffm_mat_create(ctx.h, int(rows), int(cols), int(N), -1, C.byref(h)))
        et = elem_type if elem_type is not None else (like.elem_type if like is not None else DEFAULT_TYPE)
        return cls(ctx=ctx, elem_type=et, _handle=h, _vector=vector)

    @classmethod
    def wrap_device(cls, device_ptr: int, rows: int, cols: int, ld: int, N: int, ctx=None, elem_type=DEFAULT_TYPE):
        """Device-wrapper ctor `CuModMatrix(::CuArray, N)` (CuModMatrix.jl:113-121): adopts uint32 column-major memory."""
        ctx = ctx or default_context()
        h = C.c_void_p()
        capi.check(ctx.lib.gffm_mat_wrap(ctx.h, C.c_void_p(device_ptr), rows, cols, ld, int(N), C.byref(h)))
        return cls(ctx=ctx, elem_type=elem_type, _handle=h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.gffm_mat_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- properties ----
    def _geti64(self, fn):
        v = C.c_int64(0)
        capi.check(fn(self.h, C.byref(v)))
        return v.value

    @property
    def rows(self):
        return self._geti64(self.lib.gffm_mat_rows)

    @property
    def cols(self):
        return self._geti64(self.lib.gffm_mat_cols)

    @property
    def N(self):
        v = C.c_uint64(0)
        capi.check(self.lib.gffm_mat_modulus(self.h, C.byref(v)))
        return v.value

    @property
    def shape(self):
        return (self.rows,) if self._vector else (self.rows, self.cols)

    def size(self, dim=None):
        s = self.shape
        return s if dim is None else s[dim]

    def __len__(self):
        return self.rows * self.cols

    @property
    def eltype(self):
        return self.elem_type

    @property
    def device_ptr(self):
        p = C.c_void_p()
        capi.check(self.lib.gffm_mat_device_ptr(self.h, C.byref(p)))
        return p.value

    @property
    def ld(self):
        return self._geti64(self.lib.gffm_mat_ld)

    # ---- host transfer ----
    def Array(self, dtype=None):
        """`Array(A)` (CuModMatrix.jl:256-261)."""
        dt = np.dtype(dtype) if dtype is not None else self.elem_type
        r, c = self.rows, self.cols
        out = np.zeros((r, c), dtype=dt, order="F")
        if out.size:
            capi.check(self.lib.gffm_mat_download(self.h, out.ctypes.data_as(C.c_void_p), _DT[dt], max(r, 1), 0))
        return out.reshape(-1) if self._vector else out

    def unsafe_Array(self, dtype=None):
        """`unsafe_Array(A)` (CuModMatrix.jl:251-253): including the +32 padding."""
        dt = np.dtype(dtype) if dtype is not None else self.elem_type
        r, c = self.rows + TILE_WIDTH, self.cols + TILE_WIDTH
        out = np.zeros((r, c), dtype=dt, order="F")
        capi.check(self.lib.gffm_mat_download(self.h, out.ctypes.data_as(C.c_void_p), _DT[dt], r, 1))
        return out

    def to_int(self):
        return self.Array(np.int64)

    def to_u32(self):
        """Residues as a column-major uint32 host array (half the host memory of `to_int`; used at BASELINE sizes)."""
        return self.Array(np.uint32)

    def gather_rows(self, rows_idx):
        """Host copy (int64, len(rows_idx) x cols) of the given 0-based rows: `A[rows_idx, :]` without downloading the matrix
        (one block copy per row into a small staging matrix, one download)."""
        idx = [int(i) for i in rows_idx]
        stage = CuModMatrix._new(len(idx), self.cols, self.N, like=self)
        for t, i in enumerate(idx):
            capi.check(self.lib.gffm_mat_copy_block(stage.h, t, 0, self.h, i, 0, 1, self.cols))
        return stage.Array(np.int64).reshape(len(idx), self.cols)

    def __getitem__(self, ij):
        i, j = ij if isinstance(ij, tuple) else (ij, 0)
        v = C.c_int64(0)
        capi.check(self.lib.gffm_mat_get_elem(self.h, int(i), int(j), C.byref(v)))
        return self.elem_type.type(v.value)

    def __setitem__(self, ij, value):
        i, j = ij if isinstance(ij, tuple) else (ij, 0)
        capi.check(self.lib.gffm_mat_set_elem(self.h, int(i), int(j), int(value)))

    def __repr__(self):
        return f"{self.rows}x{self.cols} CuModMatrix{{{self.elem_type.name}}} modulo {self.N}"

    def drop_cache(self):
        """Frees the cached 8-bit operand planes (they are rebuilt by the next product that needs them)."""
        capi.check(self.lib.gffm_mat_drop_cache(self.h))

    def touch(self):
        """Declare a modification made through the raw device pointer (invalidates the cached operand planes)."""
        capi.check(self.lib.gffm_mat_touch(self.h))

    def checksum(self) -> int:
        v = C.c_uint64(0)
        capi.check(self.lib.gffm_mat_checksum(self.h, C.byref(v)))
        return v.value

    def equals(self, other: "CuModMatrix") -> bool:
        e = C.c_int32(0)
        capi.check(self.lib.gffm_mat_equal(self.h, other.h, C.byref(e)))
        return bool(e.value)

    # ---- operators (kernel_ops/*.jl) ----
    def _ew_new(self, op, other=None, scalar=0):
        out = CuModMatrix._new(self.rows, self.cols, self.N, like=self, vector=self._vector)
        capi.check(self.lib.gffm_ewise(op, out.h, self.h, other.h if other is not None else None, int(scalar), 0))
        return out

    def __add__(self, o):
        return self._ew_new(capi.EW_ADD, o) if isinstance(o, CuModMatrix) else self._ew_new(capi.EW_SADD, scalar=o)

    __radd__ = __add__

    def __sub__(self, o):
        return self._ew_new(capi.EW_SUB, o) if isinstance(o, CuModMatrix) else self._ew_new(capi.EW_SSUB, scalar=o)

    def __rsub__(self, o):
        return self._ew_new(capi.EW_RSSUB, scalar=o)

    def __neg__(self):
        return self._ew_new(capi.EW_RSSUB, scalar=0)

    def __mul__(self, o):
        if isinstance(o, CuModMatrix):  # `*(A,B)` = matrix product (mul_ops.jl:54-58)
            out = CuModMatrix._new(self.rows, o.cols, self.N, like=self, vector=o._vector)
            mul_(out, self, o)
            return out
        return self._ew_new(capi.EW_SMUL, scalar=o)

    __rmul__ = __mul__
    __matmul__ = __mul__

    def __truediv__(self, s):
        return self._ew_new(capi.EW_SDIV, scalar=s)

    def __pow__(self, n: int):
        """`^` by repeated squaring (CuModMatrix.jl:307-329)."""
        if self.rows != self.cols:
            raise capi.CuModMatrixNotSquareException(capi.ERR_NOT_SQUARE, "power of a non-square matrix")
        if n < 0:
            return inverse(self) ** (-n)
        result = eye(self.elem_type, self.rows, self.N, ctx=self.ctx)
        base = copy(self)
        while n > 0:
            if n & 1:
                result = result * base
            n >>= 1
            if n:
                base = base * base
        return result


CuModVector = CuModMatrix


# ---- constructors: CuModMatrix.jl:510-556 -------------------------------------------------------------------
def zeros(T, rows, cols, N, ctx=None):
    return CuModMatrix._new(rows, cols, N, ctx=ctx, elem_type=T)


def eye(T, n, N, ctx=None):
    m = CuModMatrix._new(n, n, N, ctx=ctx, elem_type=T)
    capi.check(m.lib.gffm_mat_eye(m.h))
    return m


def rand(T, rows, cols, N, seed=0, ctx=None):
    m = CuModMatrix._new(rows, cols, N, ctx=ctx, elem_type=T)
    capi.check(m.lib.gffm_mat_rand(m.h, int(seed)))
    return m


def synth(rows, cols, N, seed, T=DEFAULT_TYPE, ctx=None):
    """SURVEY 8(d) counter-based synthetic matrix generated on the device."""
    m = CuModMatrix._new(rows, cols, N, ctx=ctx, elem_type=T)
    capi.check(m.lib.gffm_mat_synth(m.h, int(seed)))
    return m


def copy(A: CuModMatrix) -> CuModMatrix:
    out = CuModMatrix._new(A.rows, A.cols, A.N, like=A, vector=A._vector)
    capi.check(A.lib.gffm_mat_copy(out.h, A.h))
    return out


# ---- in-place API (CuModMatrix.jl:564-760, kernel_ops) ----------------------------------------------------------
def _ew(op, C_, A, B=None, scalar=0, mod_N=None):
    capi.check(C_.lib.gffm_ewise(op, C_.h, A.h, B.h if B is not None else None, int(scalar), int(mod_N) if mod_N else 0))
    return C_


def add_(C_, A, B, mod_N=None):
    return _ew(capi.EW_ADD, C_, A, B, mod_N=mod_N)


def sub_(C_, A, B, mod_N=None):
    return _ew(capi.EW_SUB, C_, A, B, mod_N=mod_N)


def elementwise_multiply_(C_, A, B, mod_N=None):
    return _ew(capi.EW_MUL, C_, A, B, mod_N=mod_N)


def negate_(C_, A, mod_N=None):
    return _ew(capi.EW_RSSUB, C_, A, scalar=0, mod_N=mod_N)


def scalar_add_(C_, A, s, mod_N=None):
    return _ew(capi.EW_SADD, C_, A, scalar=s, mod_N=mod_N)


def scalar_sub_(C_, A, s, mod_N=None):
    return _ew(capi.EW_SSUB, C_, A, scalar=s, mod_N=mod_N)


def rscalar_sub_(C_, A, s, mod_N=None):
    return _ew(capi.EW_RSSUB, C_, A, scalar=s, mod_N=mod_N)


def div_(C_, A, s, mod_N=None):
    return _ew(capi.EW_SDIV, C_, A, scalar=s, mod_N=mod_N)


def mod_elements_(A, mod_N=None):
    return _ew(capi.EW_MOD, A, A, mod_N=mod_N)


def mul_(C_, A, B, mod_N=None, R=None, P=None, mode=capi.GEMM_STORE, algo=capi.ALGO_AUTO):
    """`LinearAlgebra.mul!`: matrix (CuModMatrix.jl:767-787), vector (:816-836; kwargs R = input bound, P = modulus)
    or scalar (kernel_ops/mul_ops.jl:33-40) form."""
    if not isinstance(B, CuModMatrix):
        return _ew(capi.EW_SMUL, C_, A, scalar=B, mod_N=mod_N)
    Pm = int(P if P is not None else (mod_N or 0))
    capi.check(C_.lib.gffm_gemm(C_.h, A.h, B.h, int(R or 0), Pm, mode, algo))
    return C_


stripe_mul_ = mul_   # stripe_mul!(C,A,B) (kernel_mul/stripe_mul.jl:175-244): same result, no stripes needed
mulN_ = mul_         # CuModMatrix.jl:789-809


def matmul_host(A, B, N, out=None, ctx: Context = None):
    """`mod.(A*B, N)` for HOST matrices in one pipelined call (gffm_gemm_host): H2D copies, 8-bit plane split, tcgen05
    GEMM tiles and D2H copies overlap on three streams.  A, B: integer arrays with entries in [0, 2^32); returns uint32."""
    ctx = ctx or default_context()
    Ah = np.asfortranarray(np.asarray(A, dtype=np.uint32))
    Bh = np.asfortranarray(np.asarray(B, dtype=np.uint32))
    m, k = Ah.shape
    k2, n = Bh.shape
    if k != k2:
        raise capi.CuModArraySizeMismatchException(capi.ERR_SIZE_MISMATCH, "inner dimensions differ")
    Ch = out if out is not None else np.zeros((m, n), dtype=np.uint32, order="F")
    capi.check(ctx.lib.gffm_gemm_host(ctx.h, Ch.ctypes.data_as(C.c_void_p), m, Ah.ctypes.data_as(C.c_void_p), m, Bh.ctypes.data_as(C.c_void_p), k,
                                      m, n, k, capi.U32, int(N)))
    return Ch


def gemv_(z, A, x, R=None, P=None):
    capi.check(z.lib.gffm_gemv(z.h, A.h, x.h, int(R or 0), int(P or 0)))
    return z


def mat_mul_gpu_type(A, B, mod_N=None):
    """Legacy wrapper kept as a thin alias (kernel_mul/mat_mul_gpu_direct.jl:8-42)."""
    N = int(mod_N) if mod_N else A.N
    out = CuModMatrix._new(A.rows, B.cols, N, like=A)
    mul_(out, A, B, mod_N=N)
    return out


def mat_mul_type_inplace_(C_, A, B, mod_N=None):
    """kernel_mul/mat_mul_gpu_direct.jl:49-85."""
    return mul_(C_, A, B, mod_N=int(mod_N) if mod_N else C_.N)


def fill_(A, value):
    capi.check(A.lib.gffm_mat_fill(A.h, int(value)))
    return A


def zero_(A):
    capi.check(A.lib.gffm_mat_zero(A.h))
    return A


def copy_(dst, src):
    capi.check(dst.lib.gffm_mat_copy(dst.h, src.h))
    return dst


copyto_ = copy_


def change_modulus(A, N):
    out = copy(A)
    capi.check(out.lib.gffm_mat_set_modulus(out.h, int(N), 1))
    return out


def change_modulus_no_alloc_(A, N):
    capi.check(A.lib.gffm_mat_set_modulus(A.h, int(N), 1))
    return A


def transpose(A):
    out = CuModMatrix._new(A.cols, A.rows, A.N, like=A)
    capi.check(A.lib.gffm_mat_transpose(out.h, A.h))
    return out


# ---- permutations (rref_lu_pluq/permutations.jl) -----------------------------------------------------------------
def _pairs(P):
    arr = np.ascontiguousarray(np.asarray(P, dtype=np.int64).reshape(-1, 2)) if len(P) else np.zeros((0, 2), dtype=np.int64)
    return arr


def _apply(A, P, on_cols, inverse_):
    arr = _pairs(P)
    capi.check(A.lib.gffm_apply_perm(A.h, arr.ctypes.data_as(C.POINTER(C.c_int64)), arr.shape[0], on_cols, inverse_))
    return A


def apply_col_perm_(P, A):
    return _apply(A, P, 1, 0)


def apply_col_inv_perm_(P, A):
    return _apply(A, P, 1, 1)


def apply_row_perm_(P, A):
    return _apply(A, P, 0, 0)


def apply_row_inv_perm_(P, A):
    return _apply(A, P, 0, 1)


def perm_array_to_matrix(perm, N, new_size, perm_stack=False, ctx=None):
    """permutations.jl:141-157 (host construction + upload, as in the reference)."""
    n = len(perm)
    if perm_stack:
        n = max(new_size[0], n)
        Pm = np.eye(n, dtype=np.int64)
        for (i, j) in perm:
            Pm[[i - 1, j - 1], :] = Pm[[j - 1, i - 1], :]
    else:
        Pm = np.zeros((n, n), dtype=np.int64)
        for i in range(n):
            Pm[perm[i] - 1, i] = 1
    return CuModMatrix(Pm, N, new_size=new_size, ctx=ctx)


def mod_inv(p, P):
    """pluq_kernels.jl:11-31, evaluated by the batched device kernel."""
    return int(default_context().modinv_batch([int(p) % int(P)], int(P))[0])


# ---- elimination -----------------------------------------------------------------------------------------------
def _tuples(buf, n):
    return [(int(buf[2 * k]), int(buf[2 * k + 1])) for k in range(n)]


def pluq_gpu_kernel(A, debug=False, col_pivot_mode=capi.PIVOT_CORRECT, return_rank=False):
    """`pluq_gpu_kernel(A) -> (U, L, Perm_rows, Perm_cols)` (rref_lu_pluq/pluq_kernels.jl:46-157)."""
    U, L = C.c_void_p(), C.c_void_p()
    cap = max(A.rows, A.cols, 1)
    pr = (C.c_int64 * (2 * cap))()
    pc = (C.c_int64 * (2 * cap))()
    npr, npc, rk = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    capi.check(A.lib.gffm_pluq(A.h, C.byref(U), C.byref(L), pr, C.byref(npr), pc, C.byref(npc), C.byref(rk), col_pivot_mode))
    Um = CuModMatrix(ctx=A.ctx, elem_type=A.elem_type, _handle=U)
    Lm = CuModMatrix(ctx=A.ctx, elem_type=A.elem_type, _handle=L)
    res = (Um, Lm, _tuples(pr, npr.value), _tuples(pc, npc.value))
    return res + (rk.value,) if return_rank else res


pluq = pluq_gpu_kernel
_setup_PLUQ = pluq_gpu_kernel  # CuModMatrix.jl:335-338


def lu(A, return_pivots=False):
    """`lu(A) -> (U, L, Perm)` after the intended lu_gpu_type (test/Experiments/rref_gpu_type.jl:60-103)."""
    U, L = C.c_void_p(), C.c_void_p()
    cap = max(min(A.rows, A.cols), 1)
    pr = (C.c_int64 * (2 * cap))()
    piv = (C.c_int64 * cap)()
    npr, rk = C.c_int64(0), C.c_int64(0)
    capi.check(A.lib.gffm_lu(A.h, C.byref(U), C.byref(L), pr, C.byref(npr), piv, C.byref(rk)))
    Um = CuModMatrix(ctx=A.ctx, elem_type=A.elem_type, _handle=U)
    Lm = CuModMatrix(ctx=A.ctx, elem_type=A.elem_type, _handle=L)
    if return_pivots:
        return Um, Lm, _tuples(pr, npr.value), [int(piv[t]) for t in range(rk.value)]
    return Um, Lm, _tuples(pr, npr.value)


def rref(A, return_pivots=False):
    """`rref(A)` after the intended rref_gpu_type (test/Experiments/rref_gpu_type.jl:8-51)."""
    R = C.c_void_p()
    cap = max(min(A.rows, A.cols), 1)
    piv = (C.c_int64 * cap)()
    rk = C.c_int64(0)
    capi.check(A.lib.gffm_rref(A.h, C.byref(R), piv, C.byref(rk)))
    Rm = CuModMatrix(ctx=A.ctx, elem_type=A.elem_type, _handle=R)
    return (Rm, [int(piv[t]) for t in range(rk.value)]) if return_pivots else Rm


def rank(A):
    rk = C.c_int64(0)
    capi.check(A.lib.gffm_rank(A.h, C.byref(rk)))
    return rk.value


def is_invertible_with_inverse(A, debug=False):
    """CuModMatrix.jl:356-422 -> (Bool, Union{Nothing,CuModMatrix})."""
    out = C.c_void_p()
    ok = C.c_int32(0)
    capi.check(A.lib.gffm_inverse(A.h, C.byref(out), C.byref(ok)))
    if not ok.value:
        return False, None
    return True, CuModMatrix(ctx=A.ctx, elem_type=A.elem_type, _handle=out)


def is_invertible(A):
    """CuModMatrix.jl:460-465."""
    return A.rows == A.cols and rank(A) == A.rows


def inverse(A, debug=False):
    """CuModMatrix.jl:480-502; raises MatrixNotInvertibleException (undefined in the reference, :485)."""
    ok, inv = is_invertible_with_inverse(A)
    if not ok:
        raise capi.MatrixNotInvertibleException(capi.ERR_NOT_INVERTIBLE, "matrix is not invertible")
    return inv


def _triinv(A, upper):
    out = C.c_void_p()
    capi.check(A.lib.gffm_triinv(A.h, 1 if upper else 0, C.byref(out)))
    return CuModMatrix(ctx=A.ctx, elem_type=A.elem_type, _handle=out)


def upper_triangular_inverse_no_copy(A, debug=False):
    """triangular/triangular_inverse_no_copy.jl:197-228."""
    return _triinv(A, True)


def lower_triangular_inverse_no_copy(A, debug=False):
    """triangular/triangular_inverse_no_copy.jl:450-478."""
    return _triinv(A, False)


backward_sub_gpu_type_32 = upper_triangular_inverse_no_copy  # substitution_inplace.jl:51-56
forward_sub_gpu_type_32 = lower_triangular_inverse_no_copy   # substitution_inplace.jl:37-43


# ---- Hensel lifting of an inverse: src/CuModMatrix/triangular/hensel.jl:13-21 ------------------------------------
def hensel_pseudoinverse(N, precision, A: CuModMatrix, T: CuModMatrix) -> CuModMatrix:
    """`hensel_pseudoinverse(N, precision, A, T)`: lifts T with A*T = I (mod N) to A*T = I (mod N^precision) by the Newton
    step `T = 2*T - T*(A*T)` (hensel.jl:15-18), all products on the tensor cores.  A and T carry the modulus N^precision
    (< 2^32 here; use karatsuba.hensel_pseudoinverse for two-limb moduli).  Returns the lifted T as a CuModMatrix (the
    reference converts to a Nemo residue-ring matrix, hensel.jl:19-20; `.to_int()` gives the host integers)."""
    M = int(N) ** int(precision)
    if A.N != M or T.N != M:
        raise capi.CuModArrayModulusMismatchException(capi.ERR_MODULUS_MISMATCH, f"A and T must carry the modulus N^precision = {M}")
    if A.rows != A.cols or (T.rows, T.cols) != (A.rows, A.cols):
        raise capi.CuModMatrixNotSquareException(capi.ERR_NOT_SQUARE, "hensel_pseudoinverse needs square A and T of the same size")
    T = copy(T)
    W = CuModMatrix._new(A.rows, A.cols, M, like=A)
    V = CuModMatrix._new(A.rows, A.cols, M, like=A)
    i = 1
    while i < precision:
        mul_(W, A, T)                                  # A*T
        mul_(V, T, W)                                  # T*(A*T)
        _ew(capi.EW_SMUL, W, T, scalar=2)              # 2*T
        _ew(capi.EW_SUB, T, W, V)                      # T = 2*T - T*(A*T)
        i *= 2
    return T
