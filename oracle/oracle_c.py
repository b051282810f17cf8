"""ctypes loader for oracle/liboracle_c.so (C restatement of the CPU ground truth).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "liboracle_c.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "oracle_c.c")):
            subprocess.run(["make", "-C", HERE, "-s"], check=True)
        _lib = C.CDLL(so)
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_echelon.restype = C.c_int64
    return _lib


def num_threads():
    return load().oracle_num_threads()


def _u32f(A):
    return np.asfortranarray(np.asarray(A, dtype=np.uint32))


def matmul_mod(A, B, N, in_bound=0):
    A = _u32f(A); B = _u32f(B)
    m, k = A.shape; n = B.shape[1]
    Cm = np.zeros((m, n), dtype=np.uint32, order="F")
    p = C.POINTER(C.c_uint32)
    load().oracle_matmul_mod(A.ctypes.data_as(p), C.c_int64(max(m, 1)), B.ctypes.data_as(p), C.c_int64(max(k, 1)), Cm.ctypes.data_as(p),
                             C.c_int64(max(m, 1)), C.c_int64(m), C.c_int64(k), C.c_int64(n), C.c_uint64(N), C.c_uint64(in_bound))
    return Cm.astype(np.int64)


def echelon(A, N):
    W = _u32f(np.mod(np.asarray(A, dtype=np.int64), N))
    m, n = W.shape
    L = np.zeros((m, m), dtype=np.uint32, order="F")
    piv = np.zeros(max(min(m, n), 1), dtype=np.int64); swp = np.zeros(max(min(m, n), 1), dtype=np.int64)
    p = C.POINTER(C.c_uint32); q = C.POINTER(C.c_int64)
    r = load().oracle_echelon(W.ctypes.data_as(p), C.c_int64(max(m, 1)), L.ctypes.data_as(p), C.c_int64(max(m, 1)), C.c_int64(m), C.c_int64(n),
                              C.c_uint64(N), piv.ctypes.data_as(q), swp.ctypes.data_as(q))
    perm_rows = [(t + 1, int(swp[t]) + 1) for t in range(r) if swp[t] != t]
    return W.astype(np.int64), L.astype(np.int64), perm_rows, [int(c) for c in piv[:r]]
