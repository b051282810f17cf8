"""Timeline of the pipelined host GEMM: GFFM_HOST_PROF=1 python tools/host_pipe_prof.py [n] [N]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gffm_b200 as g
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 33554393
ctx = g.default_context()
hA = torch.randint(0, N, (n, n), dtype=torch.int32).pin_memory()
hB = torch.randint(0, N, (n, n), dtype=torch.int32).pin_memory()
hC = torch.empty((n, n), dtype=torch.int32).pin_memory()
for rep in range(3):
    t0 = time.perf_counter()
    g.capi.check(ctx.lib.gffm_gemm_host(ctx.h, hC.data_ptr(), n, hA.data_ptr(), n, hB.data_ptr(), n, n, n, n, g.capi.U32, N))
    print(f"rep {rep}: {1e3*(time.perf_counter()-t0):.2f} ms", flush=True)
# raw copy rates for reference
d = torch.empty((n, n), dtype=torch.int32, device="cuda")
for name, fn in (("H2D 1GiB contiguous", lambda: d.copy_(hA, non_blocking=True)), ("D2H 1GiB contiguous", lambda: hC.copy_(d, non_blocking=True))):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name}: {dt*1e3:.2f} ms = {4*n*n/dt/1e9:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(hA, non_blocking=True)
d2 = torch.empty_like(d)
with torch.cuda.stream(s2): hC.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D + D2H concurrently (1 GiB each): {dt*1e3:.2f} ms")
