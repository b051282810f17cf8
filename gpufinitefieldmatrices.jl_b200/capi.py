"""ctypes binding of libgffm.so -- a 1:1 mirror of include/gffm.h (and of the Julia `ccall` shim in
julia/GPUFiniteFieldMatricesB200.jl).  There is no CPU fallback: importing works everywhere (so the ABI can be
checked on a GPU-less host) but creating a context without a CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libgffm.so")

# status codes / enums (include/gffm.h)
OK, ERR_INVALID, ERR_SIZE_MISMATCH, ERR_MODULUS_MISMATCH, ERR_MODULUS_TOO_LARGE, ERR_NOT_SQUARE, ERR_NOT_INVERTIBLE, \
    ERR_INVERSE_NOT_DEFINED, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_INEXACT, ERR_OOM, ERR_MODULUS_NOT_PRIME = range(14)
F32, F64, I64, U32, I32 = range(5)
EW_MOD, EW_ADD, EW_SUB, EW_MUL, EW_SADD, EW_SSUB, EW_RSSUB, EW_SMUL, EW_SDIV = range(9)
GEMM_STORE, GEMM_ADD, GEMM_SUB = range(3)
ALGO_AUTO, ALGO_SIMT, ALGO_LIMB, ALGO_RNS = range(4)
PIVOT_CORRECT, PIVOT_REFERENCE_QUIRK = range(2)
MG_AUTO, MG_NCCL_BCAST, MG_NCCL_PLANES, MG_P2P_PLANES, MG_P2P_PUSH, MG_P2P_RAW = range(6)
MG_DISTRIBUTED = -1


class GffmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[gffm status {code}] {msg}")
        self.code = code


# exceptions mirroring src/CuModMatrix/CuModMatrix.jl:5-31
class CuModArraySizeMismatchException(GffmError):
    pass


class CuModArrayModulusMismatchException(GffmError):
    pass


class CuModMatrixNotSquareException(GffmError):
    pass


class MatrixNotInvertibleException(GffmError):
    pass


class InverseNotDefinedException(GffmError):
    pass


class InexactError(GffmError):
    pass


class CuModMatrixModulusNotPrimeException(GffmError):
    pass


_EXC = {
    ERR_SIZE_MISMATCH: CuModArraySizeMismatchException,
    ERR_MODULUS_MISMATCH: CuModArrayModulusMismatchException,
    ERR_MODULUS_TOO_LARGE: CuModArrayModulusMismatchException,
    ERR_NOT_SQUARE: CuModMatrixNotSquareException,
    ERR_NOT_INVERTIBLE: MatrixNotInvertibleException,
    ERR_INVERSE_NOT_DEFINED: InverseNotDefinedException,
    ERR_INEXACT: InexactError,
    ERR_MODULUS_NOT_PRIME: CuModMatrixModulusNotPrimeException,
}

_vp = C.c_void_p
_i32, _i64, _u64 = C.c_int32, C.c_int64, C.c_uint64
_pi32, _pi64, _pu64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint64)
_pvp = C.POINTER(C.c_void_p)

# name -> argtypes ; every function of include/gffm.h (restype int32 unless noted)
SIGNATURES = {
    "gffm_device_count": [_pi32],
    "gffm_create": [_i32, _pvp],
    "gffm_destroy": [_vp],
    "gffm_sync": [_vp],
    "gffm_set_stream": [_vp, _vp],
    "gffm_get_stream": [_vp, _pvp],
    "gffm_set_profiling": [_vp, _i32],
    "gffm_last_timings": [_vp, C.POINTER(C.c_double), _i32, _pi32],
    "gffm_set_gemm_ctas": [_vp, _i32],
    "gffm_launch_count": [_vp, _pi64],
    "gffm_alloc_stats": [_vp, _pi64, _pi64],
    "gffm_mat_create": [_vp, _i64, _i64, _u64, _i32, _pvp],
    "gffm_mat_wrap": [_vp, _vp, _i64, _i64, _i64, _u64, _pvp],
    "gffm_mat_destroy": [_vp],
    "gffm_mat_drop_cache": [_vp],
    "gffm_mat_touch": [_vp],
    "gffm_mat_upload": [_vp, _vp, _i32, _i64, _i32],
    "gffm_mat_download": [_vp, _vp, _i32, _i64, _i32],
    "gffm_mat_rows": [_vp, _pi64],
    "gffm_mat_cols": [_vp, _pi64],
    "gffm_mat_pad": [_vp, _pi32],
    "gffm_mat_modulus": [_vp, _pu64],
    "gffm_mat_ld": [_vp, _pi64],
    "gffm_mat_device_ptr": [_vp, _pvp],
    "gffm_mat_set_modulus": [_vp, _u64, _i32],
    "gffm_mat_copy": [_vp, _vp],
    "gffm_mat_fill": [_vp, _i64],
    "gffm_mat_zero": [_vp],
    "gffm_mat_eye": [_vp],
    "gffm_mat_rand": [_vp, _u64],
    "gffm_mat_synth": [_vp, _u64],
    "gffm_mat_get_elem": [_vp, _i64, _i64, _pi64],
    "gffm_mat_set_elem": [_vp, _i64, _i64, _i64],
    "gffm_mat_transpose": [_vp, _vp],
    "gffm_mat_copy_block": [_vp, _i64, _i64, _vp, _i64, _i64, _i64, _i64],
    "gffm_mat_equal": [_vp, _vp, _pi32],
    "gffm_mat_checksum": [_vp, _pu64],
    "gffm_gemm": [_vp, _vp, _vp, _u64, _u64, _i32, _i32],
    "gffm_gemm_block": [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _i64, _u64, _u64, _i32, _i32],
    "gffm_gemm_host": [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _i32, _u64],
    "gffm_gemm_panels": [_vp, _vp, _vp, _i32, _pi64, _pvp, _pvp, _u64, _u64],
    "gffm_gemv": [_vp, _vp, _vp, _u64, _u64],
    "gffm_mg_unique_id": [_vp],
    "gffm_mg_create": [_vp, _vp, _i32, _i32, _pvp],
    "gffm_mg_destroy": [_vp],
    "gffm_mg_info": [_vp, _pi32, _pi32, _pi32, _pi32],
    "gffm_mg_set_transport": [_vp, _i32],
    "gffm_mg_barrier": [_vp],
    "gffm_mg_owner_ranges": [_i64, _i32, _pi64],
    "gffm_mg_owner_ranges_root_free": [_i64, _i32, _i32, _pi64],
    "gffm_mg_gemm": [_vp, _vp, _vp, _vp, _i32, _vp, _u64, _u64],
    "gffm_mg_kmat_mul": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _u64, _i32, _vp],
    "gffm_mg_gemv": [_vp, _vp, _vp, _vp, _i32, _u64, _u64],
    "gffm_ewise": [_i32, _vp, _vp, _vp, _i64, _u64],
    "gffm_pluq": [_vp, _pvp, _pvp, _pi64, _pi64, _pi64, _pi64, _pi64, _i32],
    "gffm_lu": [_vp, _pvp, _pvp, _pi64, _pi64, _pi64, _pi64],
    "gffm_rref": [_vp, _pvp, _pi64, _pi64],
    "gffm_rank": [_vp, _pi64],
    "gffm_inverse": [_vp, _pvp, _pi32],
    "gffm_triinv": [_vp, _i32, _pvp],
    "gffm_apply_perm": [_vp, _pi64, _i64, _i32, _i32],
    "gffm_modinv_batch": [_vp, _pu64, _pu64, _i64, _u64],
    "gffm_kmat_mul": [_vp, _vp, _vp, _vp, _vp, _vp, _u64, _u64],
    "gffm_kmat_ewise": [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _u64, _u64],
}
STRING_FUNCS = ["gffm_version", "gffm_last_error"]

_lib = None


def load():
    """dlopen libgffm.so (built in-tree by build.py / __graft_entry__.build()).  Fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                          "There is no CPU fallback for this package.")
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int32
    for name in STRING_FUNCS:
        fn = getattr(lib, name)
        fn.argtypes = []
        fn.restype = C.c_char_p
    _lib = lib
    return lib


def check(status):
    if status != OK:
        msg = load().gffm_last_error().decode("utf-8", "replace")
        raise _EXC.get(status, GffmError)(status, msg)


def version():
    return load().gffm_version().decode()


def device_count():
    n = C.c_int32(0)
    check(load().gffm_device_count(C.byref(n)))
    return n.value
