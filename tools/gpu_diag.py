"""Stage-by-stage GPU bring-up diagnostics (each stage runs in its own process under a timeout so that a hung
kernel cannot take the whole gpurun call down).  Usage: python tools/gpu_diag.py [stage ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["basic", "simt", "l1", "l2", "rns", "big", "elim", "kara"]  # + "elimq" (quick elimination subset, on request)


def report_mismatch(name, got, want):
    import numpy as np
    bad = np.argwhere(got != want)
    print(f"[{name}] shape={got.shape} mismatches={len(bad)}/{got.size}")
    if len(bad):
        rows = np.unique(bad[:, 0]); cols = np.unique(bad[:, 1])
        print(f"   bad rows: n={len(rows)} first={rows[:16].tolist()} last={rows[-4:].tolist()}")
        print(f"   bad cols: n={len(cols)} first={cols[:16].tolist()} last={cols[-4:].tolist()}")
        for (i, j) in bad[:8]:
            print(f"   ({i},{j}): got {got[i, j]} want {want[i, j]}")
    return len(bad) == 0


def gemm_case(g, O, m, k, n, N, algo, seed=1, mode=0, allmax=False):
    import numpy as np
    A = O.synth_matrix(seed, m, k, N); B = O.synth_matrix(seed + 1, k, n, N)
    if allmax:
        A[:] = N - 1; B[:] = N - 1
    Ag = g.CuModMatrix(A, N); Bg = g.CuModMatrix(B, N)
    C0 = O.synth_matrix(seed + 2, m, n, N)
    Cg = g.CuModMatrix(C0, N)
    t0 = time.time()
    g.mul_(Cg, Ag, Bg, mode=mode, algo=algo)
    Cg.ctx.sync()
    dt = time.time() - t0
    got = Cg.to_int()
    AB = O.matmul_mod(A, B, N) if m * k * n <= 2 ** 31 else None
    if AB is None:
        Cs = g.CuModMatrix(C0, N)
        g.mul_(Cs, Ag, Bg, mode=mode, algo=g.capi.ALGO_SIMT)
        want = Cs.to_int()
    else:
        want = AB if mode == 0 else (np.mod(C0 + AB, N) if mode == 1 else np.mod(C0 - AB, N))
    ok = report_mismatch(f"gemm m={m} k={k} n={n} N={N} algo={algo} mode={mode} allmax={allmax} t={dt*1e3:.1f}ms", got, want)
    return ok


def stage(name):
    import numpy as np
    import gffm_b200 as g
    from oracle import oracle as O
    cap = g.capi
    ok = True
    if name == "basic":
        A = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]]); B = np.array([[9, 8, 7], [6, 5, 4], [3, 2, 1]])
        Ag = g.CuModMatrix(A, 11); Bg = g.CuModMatrix(B, 11)
        print("roundtrip", np.array_equal(Ag.to_int(), A), "padded shape", Ag.unsafe_Array().shape)
        print("add", (Ag + Bg).to_int().tolist()); print("sub", (Ag - Bg).to_int().tolist())
        print("neg ctor", g.CuModMatrix(np.array([[-3, 12.0]]), 7).to_int().tolist())
        print("mul3x3", (Ag * Bg).to_int().tolist(), O.matmul_mod(A, B, 11).tolist())
    elif name == "simt":
        for (m, k, n, N) in [(5, 7, 3, 11), (100, 100, 100, 2 ** 11), (130, 70, 90, 33554393), (64, 64, 64, 4294967291)]:
            ok &= gemm_case(g, O, m, k, n, N, cap.ALGO_SIMT)
    elif name == "l1":
        for (m, k, n, N) in [(128, 128, 256, 11), (128, 256, 256, 11), (256, 128, 512, 251), (100, 100, 100, 7), (300, 500, 700, 11)]:
            ok &= gemm_case(g, O, m, k, n, N, cap.ALGO_LIMB)
        ok &= gemm_case(g, O, 200, 300, 260, 256, cap.ALGO_LIMB, allmax=True)
        ok &= gemm_case(g, O, 200, 300, 260, 11, cap.ALGO_LIMB, mode=1)
        ok &= gemm_case(g, O, 200, 300, 260, 11, cap.ALGO_LIMB, mode=2)
    elif name == "l2":
        for (m, k, n, N) in [(128, 128, 128, 65521), (128, 512, 128, 65521), (100, 100, 100, 2 ** 11), (300, 500, 700, 11 ** 3), (257, 1000, 129, 65521)]:
            ok &= gemm_case(g, O, m, k, n, N, cap.ALGO_LIMB)
        ok &= gemm_case(g, O, 200, 1024, 260, 65536, cap.ALGO_LIMB, allmax=True)
        ok &= gemm_case(g, O, 200, 300, 260, 65521, cap.ALGO_LIMB, mode=2)
    elif name == "rns":
        for (m, k, n, N) in [(128, 128, 256, 33554393), (128, 512, 256, 33554393), (100, 100, 100, 16777213), (300, 500, 700, 33554393), (257, 1000, 129, 2 ** 26)]:
            ok &= gemm_case(g, O, m, k, n, N, cap.ALGO_RNS)
        ok &= gemm_case(g, O, 200, 1024, 260, 33554393, cap.ALGO_RNS, allmax=True)
        ok &= gemm_case(g, O, 200, 300, 260, 33554393, cap.ALGO_RNS, mode=1)
        ok &= gemm_case(g, O, 200, 300, 260, 33554393, cap.ALGO_RNS, mode=2)
        ok &= gemm_case(g, O, 200, 300, 260, 11, cap.ALGO_RNS)
    elif name == "big":
        for (n, N, algo) in [(2048, 11, cap.ALGO_LIMB), (2048, 65521, cap.ALGO_LIMB), (2048, 33554393, cap.ALGO_RNS)]:
            ok &= gemm_case(g, O, n, n, n, N, algo)
        for (n, N) in [(4096, 11), (4096, 65521), (4096, 33554393), (8192, 33554393), (16384, 11), (16384, 65521), (16384, 33554393)]:
            A = g.synth(n, n, N, 5); B = g.synth(n, n, N, 6); Cm = g.zeros(np.float32, n, n, N)
            g.mul_(Cm, A, B); Cm.ctx.sync()
            t0 = time.time(); g.mul_(Cm, A, B); Cm.ctx.sync(); dt = time.time() - t0
            print(f"[perf] n={n} N={N}: {dt*1e3:.2f} ms  {2*n**3/dt/1e12:.1f} effective TOPS  checksum={Cm.checksum():016x}")
    elif name in ("elim", "elimq"):
        cases = [(10, 10, 7, 1), (40, 40, 7, 2), (100, 100, 65521, 3), (300, 300, 33554393, 4), (700, 500, 65521, 5), (500, 700, 11, 6), (1000, 1000, 65521, 7)]
        if name == "elimq":  # quick variant for kernel iteration (the oracle dominates the full stage)
            cases = [(10, 10, 7, 1), (100, 100, 65521, 3), (200, 200, 33554393, 4), (330, 250, 65521, 5), (250, 330, 11, 6)]
        for (m, n, N, seed) in cases:
            A = O.synth_matrix(seed, m, n, N)
            if seed in (5, 6):
                A[:, 3] = 0; A[:, 11] = (3 * A[:, 1] + A[:, 2]) % N
            Ag = g.CuModMatrix(A, N)
            t0 = time.time()
            U, L, pr, pc, rk = g.pluq_gpu_kernel(Ag, return_rank=True)
            dt = time.time() - t0
            Uo, Lo, pro, pco, rko = O.pluq(A, N)
            okU = np.array_equal(U.to_int(), Uo); okL = np.array_equal(L.to_int(), Lo)
            print(f"[pluq {m}x{n} mod {N}] rank {rk}/{rko} U {okU} L {okL} prow {pr == pro} pcol {pc == pco} t={dt*1e3:.1f}ms")
            if not okU:
                report_mismatch("U", U.to_int(), Uo)
            if not okL:
                report_mismatch("L", L.to_int(), Lo)
            ok &= okU and okL and pr == pro and pc == pco and rk == rko
            R, piv = g.rref(Ag, return_pivots=True)
            Ro, pivo = O.rref(A, N)
            okR = np.array_equal(R.to_int(), Ro) and piv == pivo
            print(f"   rref {okR}")
            ok &= okR
            if m == n:
                okf, inv = g.is_invertible_with_inverse(Ag)
                oko, invo = O.is_invertible_with_inverse(A, N)
                oki = okf == oko and (not okf or np.array_equal(inv.to_int(), invo))
                print(f"   inverse invertible={okf} match={oki}")
                ok &= oki
        n = 4096
        for N in ((65521, 7, 33554393) if name == "elim" else ()):
            Ag = g.synth(n, n, N, 9)
            t0 = time.time(); U, L, pr, pc, rk = g.pluq_gpu_kernel(Ag, return_rank=True); dt = time.time() - t0
            print(f"[pluq perf] n={n} N={N} rank={rk} {dt*1e3:.1f} ms")
    elif name == "kara":
        for (n, N1, N2) in [(64, 13 ** 4, 13 ** 3), (500, 13 ** 4, 13 ** 3), (300, 8191, 8191), (200, 11, 11)]:
            rng = np.random.default_rng(n)
            A1 = rng.integers(0, N1, size=(n, n)); A2 = rng.integers(0, N2, size=(n, n))
            B1 = rng.integers(0, N1, size=(n, n)); B2 = rng.integers(0, N2, size=(n, n))
            AK = g.KaratsubaMatrix(g.CuModMatrix(A1, N1), g.CuModMatrix(A2, N1), N1, N2)
            BK = g.KaratsubaMatrix(g.CuModMatrix(B1, N1), g.CuModMatrix(B2, N1), N1, N2)
            CK = g.KaratsubaZeros(np.float64, n, n, N1, N2)
            g.KMatMul_(CK, AK, BK)
            D1, D2 = O.karatsuba_matmul_direct(A1, A2, B1, B2, N1, N2)
            ok1 = report_mismatch(f"kara n={n} N1={N1} N2={N2} C1", CK.data1.to_int(), D1)
            ok2 = report_mismatch(f"kara n={n} N1={N1} N2={N2} C2", CK.data2.to_int(), D2)
            ok &= ok1 and ok2
    print(f"STAGE {name}: {'OK' if ok else 'FAIL'}")
    return 0 if ok else 1


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--stage":
        sys.exit(stage(sys.argv[2]))
    stages = sys.argv[1:] or STAGES
    rc = 0
    for s in stages:
        print(f"===== stage {s} =====", flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", s], timeout=420)
            code = r.returncode
        except subprocess.TimeoutExpired:
            code = -999
            print(f"STAGE {s}: TIMEOUT")
        print(f"===== stage {s} rc={code} {time.time()-t0:.1f}s =====", flush=True)
        rc |= (code != 0)
    sys.exit(1 if rc else 0)
