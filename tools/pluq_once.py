"""Times the blocked PLUQ / RREF of an n x n synthetic matrix mod N through the C ABI (warm: best and median of 3 after one warm-up).
    python tools/pluq_once.py 16384 65521"""
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gffm_b200 as g  # noqa: E402

n, N = int(sys.argv[1]), int(sys.argv[2])
ctx = g.default_context()
A = g.synth(n, n, N, 9)
ts = []
for rep in range(4):
    ctx.sync(); t0 = time.perf_counter()
    U, L, pr, pc, rk = g.pluq_gpu_kernel(A, return_rank=True)
    ctx.sync(); ts.append(time.perf_counter() - t0)
    del U, L
tr = []
for rep in range(3):
    ctx.sync(); t0 = time.perf_counter()
    R = g.rref(A)
    ctx.sync(); tr.append(time.perf_counter() - t0)
    del R
print(f"rank {rk}  pluq best {min(ts[1:])*1e3:.1f} ms median {statistics.median(ts[1:])*1e3:.1f} ms  rref best {min(tr[1:])*1e3:.1f} ms")
