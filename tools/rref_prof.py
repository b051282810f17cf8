"""RREF / inverse timing (warm): python tools/rref_prof.py n N [n N ...]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gffm_b200 as g
ctx = g.default_context()
args = [int(x) for x in sys.argv[1:]] or [4096, 65521]
for n, N in zip(args[0::2], args[1::2]):
    A = g.synth(n, n, N, 9)
    for rep in range(3):
        ctx.sync(); t0 = time.perf_counter()
        R = g.rref(A)
        ctx.sync(); dt = time.perf_counter() - t0
        del R
        print(f"[rref] n={n} N={N} rep={rep} {dt*1e3:.1f} ms", flush=True)
    for rep in range(2):
        ctx.sync(); t0 = time.perf_counter()
        ok, inv = g.is_invertible_with_inverse(A)
        ctx.sync(); dt = time.perf_counter() - t0
        del inv
        print(f"[inverse] n={n} N={N} rep={rep} invertible={ok} {dt*1e3:.1f} ms", flush=True)
