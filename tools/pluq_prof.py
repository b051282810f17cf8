"""PLUQ phase timing: python tools/pluq_prof.py n N [n N ...]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gffm_b200 as g
ctx = g.default_context()
args = [int(x) for x in sys.argv[1:]] or [4096, 65521]
for n, N in zip(args[0::2], args[1::2]):
    A = g.synth(n, n, N, 9)
    for rep in range(3):
        ctx.set_profiling(rep < 2)
        ctx.sync(); t0 = time.perf_counter()
        U, L, pr, pc, rk = g.pluq_gpu_kernel(A, return_rank=True)
        ctx.sync(); dt = time.perf_counter() - t0
        ph = ctx.last_timings() if rep < 2 else []
        ctx.set_profiling(False)
        t1 = time.perf_counter()
        del U, L
        ctx.sync(); dfree = time.perf_counter() - t1
        print(f"[pluq] n={n} N={N} rep={rep} rank={rk} total={dt*1e3:.1f} ms phases(panel,u12,trailing)={[round(x,1) for x in ph]} free={dfree*1e3:.1f} ms", flush=True)
