#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: exact mod-p matrix multiplication, n = 16384, 25-bit prime.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

metric (BASELINE.json): mod-p matmul effective GOPS = 2 n^3 / t.
  value      : device-timed (CUDA events), inputs already resident in HBM; max over ranks; whole job.
  e2e        : same product through the public API with HOST (pinned) buffers: H2D of A and B, GEMM, D2H of C per step.
  roofline   : int8 tensor-pipe ops of the dominant kernel (tcgen05 RNS GEMM) / its CUDA-event duration, vs peak.
  cpu_baseline: C restatement of the reference tests' `mod.(A*B, N)` ground truth on the host cores (bounded sample).
  --impl reference: that CPU arm alone (the Julia reference cannot run in this image).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_DEFAULT = 16384
MOD_DEFAULT = 33554393  # 25-bit prime near the 2^26 limit (BASELINE config 2)
SEED_A, SEED_B = 5, 6   # SURVEY 8(d) "metric GEMM" seeds


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16": d.get("bf16_tflops", 1590.0), "bf16_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Start of the timed region: only samples taken from here on are reported (the sampler itself is started before the
        warm-up so that nvidia-smi is already polling when a short timed region begins)."""
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        first = getattr(self, "first", 0)
        lines = self.lines[first:] if len(self.lines) - first >= 2 else self.lines[max(0, first - 3):]  # timed region shorter than the poll period
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


CPU_SAMPLE_N = 4096  # fixed CPU sample (the same sub-product on every host, so ratios are comparable across runs and GPU counts)


def cpu_arm(n_full, N, ns=CPU_SAMPLE_N, threads=None):
    """Times the C oracle (oracle/oracle_c.c, pthreads over all host cores) on a FIXED bounded sample of the workload: the leading
    ns x ns x ns sub-product of the same synthetic matrices (ns = 4096: ~2-5 s on 16-32 cores).  Returns (GOPS, cores, description,
    seconds).  Any time for the full n is an n^3 extrapolation and labelled as such by the callers."""
    import numpy as np
    from oracle import oracle as O
    from oracle import oracle_c as OC
    cores = OC.num_threads() if threads is None else threads
    ns = min(ns, n_full)
    A = O.synth_matrix(SEED_A, ns, ns, N); B = O.synth_matrix(SEED_B, ns, ns, N)
    t0 = time.perf_counter(); C = OC.matmul_mod(A, B, N); dt = time.perf_counter() - t0
    chk = int(np.bitwise_xor.reduce(C.reshape(-1)))
    gops = 2.0 * ns ** 3 / dt / 1e9
    return gops, cores, f"{ns}x{ns}x{ns} mod {N} sub-product of the n={n_full} workload, same generator (xor checksum {chk:#x})", dt


def cpu_echelon_arm(N, sizes=(1024, 2048)):
    """CPU baseline of the elimination: the C oracle's echelon form (pluq_kernels.jl:46-157 conventions, one core, scalar) on full-rank
    synthetic matrices; field mul-adds n^3/3."""
    from oracle import oracle as O
    from oracle import oracle_c as OC
    out = []
    for n in sizes:
        A = O.synth_matrix(9, n, n, N)
        t0 = time.perf_counter(); _, _, _, piv = OC.echelon(A, N); dt = time.perf_counter() - t0
        out.append({"n": n, "seconds": dt, "rank": len(piv), "field_muladds_per_s": (n ** 3 / 3.0) / dt})
    return out


def stripe_arm(n_full, N, ns=1024):
    """The reference's OWN algorithm (kernel_mul/stripe_mul.jl:175-244: float64 GEMM stripes of the width that keeps sums below
    2^53, a floored mod after every stripe) restated with numpy/BLAS on the host, timed on a small sub-product.  Informational:
    it shows what the stripe formulation costs for this modulus next to the integer restatement used as the CPU baseline."""
    from oracle import oracle as O
    ns = min(ns, n_full)
    A = O.synth_matrix(SEED_A, ns, ns, N); B = O.synth_matrix(SEED_B, ns, ns, N)
    t0 = time.perf_counter(); C = O.stripe_mul(A, B, N); dt = time.perf_counter() - t0
    width = max(1, min(ns, O.find_max_stripe_ops(53, N)))
    return {"value": 2.0 * ns ** 3 / dt / 1e9, "unit": "GOPS", "kind": "port", "algorithm": "float64 K-stripes + mod per stripe (stripe_mul.jl:175-244) via numpy BLAS",
            "stripe_width": int(width), "stripes": int((ns + width - 1) // width), "sample": f"{ns}x{ns}x{ns} mod {N}", "seconds": dt,
            "matches_integer_port": None if ns > 2048 else bool((C == O.exact_matmul_mod(A, B, N)).all())}


def nccl_log_lines(limit=12):
    """The 'nranks' / 'Init COMPLETE' lines NCCL wrote to NCCL_DEBUG_FILE (one file per process); also echoed to stderr."""
    import glob
    pat = os.environ.get("NCCL_DEBUG_FILE", "").replace("%h", "*").replace("%p", "*")
    lines = []
    for f in sorted(glob.glob(pat)):
        try:
            for ln in open(f, errors="replace"):
                if "nranks" in ln and ("Init COMPLETE" in ln or "ncclCommInitRank" in ln):
                    lines.append(ln.strip()[:300])
        except OSError:
            pass
    for ln in lines[:64]:
        print(ln, file=sys.stderr)
    return {"files": len(glob.glob(pat)), "nranks_lines": lines[:limit], "count": len(lines)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    gops_all = []
    desc = ""
    cores = 1
    for _ in range(max(1, args.warmup if args.warmup < 2 else 1)):
        pass
    for i in range(max(1, min(args.steps, 3))):
        gops, cores, desc, dt = cpu_arm(args.n, args.modulus)
        gops_all.append(gops)
    value = statistics.median(gops_all)
    out = {
        "impl": "reference", "metric": "mod-p matmul effective GOPS (2n^3/s)", "value": value, "unit": "GOPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 2.0 * args.n ** 3 / (value * 1e9) * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64 accumulate of u32 residues (host integers)", "data": "synthetic",
        "config": {"workload": f"{args.n}x{args.n} * {args.n}x{args.n} matmul mod {args.modulus} ({(args.modulus - 1).bit_length()}-bit modulus), A,B resident as uint32 residues",
                   "n": args.n, "modulus": args.modulus,
                   "note": "the Julia reference cannot run in this image; this arm is the reference tests' CPU ground truth mod.(A*B,N) restated in C "
                           "(oracle/oracle_c.c) on all host cores; every step is the FIXED 4096^3 sub-product (same on every host); "
                           "ms_per_step is EXTRAPOLATED by n^3 to the full workload", "cpu_sample_n": min(CPU_SAMPLE_N, args.n), "ms_per_step_is_extrapolated": True},
        "cpu_baseline": {"value": value, "unit": "GOPS", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "GOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        out["reference_algorithm_on_cpu"] = stripe_arm(args.n, args.modulus)
    except Exception as e:  # informational only
        out["reference_algorithm_on_cpu"] = {"error": str(e)[:200]}
    print(json.dumps(out))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--n", type=int, default=N_DEFAULT)
    ap.add_argument("--modulus", type=int, default=MOD_DEFAULT)
    ap.add_argument("--panels", default="8", help="column panels of B for the pipelined NCCL broadcast (N > 1); several values: "
                    "the fastest is picked during the untimed warm-up")
    ap.add_argument("--gemm-ctas", default="0", help="cap on the persistent GEMM grid (0 = all SMs) so that the concurrent NCCL "
                    "kernels find free SMs (N > 1); several values: the fastest is picked during the untimed warm-up (measured on 8 B200: "
                    "148 > 140 > 132 CTAs, profiles/r01_notes.md, hence the default)")
    ap.add_argument("--grid-cols", type=int, default=1, help="N > 1: column groups of a 2-D process grid (1 = row blocks of A with a full "
                    "broadcast of B; pc > 1: rank (i, j) multiplies row block i of A with column range j of B and receives only that range)")
    ap.add_argument("--bcast", default="broadcast", choices=["broadcast", "scatter_allgather"],
                    help="how B is replicated every step (N > 1): ncclBroadcast per panel, or scatter + in-place all-gather per panel")
    ap.add_argument("--e2e-mode", default="replicated", choices=["replicated", "sliced"],
                    help="N > 1 end-to-end arm: 'replicated' = every rank uploads its A shard and all of B from its host copy (default); "
                         "'sliced' additionally times: every rank uploads its A shard and 1/N of each B panel, the panels are completed by an "
                         "in-place NCCL all-gather and consumed by gffm_gemm_panels (opt-in until measured on 8 GPUs)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--extras", action="store_true", help="(default at N = 1; kept for compatibility)")
    ap.add_argument("--no-extras", action="store_true", help="skip config.extras (other moduli, PLUQ / RREF / inverse n = 16384, GEMV, the reference's "
                    "n = 5000 timing cases) and the PLUQ roofline / CPU echelon baseline")
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled-row oracle check of the timed result (parity_check)")
    ap.add_argument("--parity-rows", type=int, default=64)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import gffm_b200 as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's own log is evidence of the rank count: keep it, but away from stdout (ONE JSON line there).  With NCCL_DEBUG set and no
        # NCCL_DEBUG_FILE the log goes to gpurun_out/nccl_debug.<host>.<pid>.log; rank 0 copies the "nranks" lines to stderr and into the
        # JSON line (config.nccl_log) after the run.
        if os.environ.get("GFFM_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = os.environ["GFFM_NCCL_DEBUG"]
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            os.environ["NCCL_DEBUG_FILE"] = os.path.join(ROOT, "gpurun_out", "nccl_debug.%h.%p.log")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, N = args.n, args.modulus
    W = max(3, args.warmup)
    K = max(1, args.steps)
    ctx = g.Context(local)
    stream = torch.cuda.Stream(device=local)
    ctx.set_stream(stream.cuda_stream)
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        # ---- resident inputs: A row block of this rank, B (broadcast from rank 0 each step when world > 1) -----------
        mg = g.multigpu
        pr, pc = mg.process_grid(world, args.grid_cols if world > 1 else 1)
        gi, gj = mg.grid_coords(rank, pr, pc)
        r0, r1 = mg.row_block(n, pr, gi)
        mloc = r1 - r0
        cl0, cl1 = mg.col_range(n, pc, gj, align=mg.PANEL_ALIGN)
        ncl = cl1 - cl0  # columns of B / C this rank works on (all of them unless --grid-cols > 1)
        if world == 1:
            A = g.synth(n, n, N, SEED_A, ctx=ctx)
        else:
            Afull = g.synth(n, n, N, SEED_A, ctx=ctx)
            A = g.zeros(np.float32, mloc, n, N, ctx=ctx)
            g.capi.check(A.lib.gffm_mat_copy_block(A.h, 0, 0, Afull.h, r0, 0, mloc, n))
            ctx.sync()
            del Afull
        ldb = ((n + 31) // 32) * 32
        Bt = torch.zeros((n, ldb), dtype=torch.int32, device=f"cuda:{local}")  # column-major n x n, leading dim ldb
        Ball = g.CuModMatrix.wrap_device(Bt.data_ptr(), n, n, ldb, N, ctx=ctx)
        B = Ball if pc == 1 else g.CuModMatrix.wrap_device(Bt.data_ptr() + 4 * cl0 * ldb, n, ncl, ldb, N, ctx=ctx)  # this rank's column range
        if rank == 0:
            Bs = g.synth(n, n, N, SEED_B, ctx=ctx)
            g.copy_(Ball, Bs)
            ctx.sync()
            del Bs
        C = g.zeros(np.float32, mloc, ncl, N, ctx=ctx)
        col_groups = mg.make_column_groups(dist, world, pc, src=0) if pc > 1 else None
        # N > 1: NCCL broadcast of B in column panels on a communication stream; ONE gffm_gemm_panels call per step consumes them
        # (split of panel p+1 / CRT of panel p under the GEMM of panel p; the next step's broadcast runs under this step's GEMMs)
        pan_cands = [int(x) for x in str(args.panels).split(",")] if world > 1 else [1]
        cta_cands = [int(x) for x in str(args.gemm_ctas).split(",")] if world > 1 else [0]
        bm = None

        def make_bm(npanels):
            pans = mg.col_panels(ncl, npanels, align=mg.PANEL_ALIGN)  # relative to this rank's column range
            if world == 1:
                return pans, None
            deliver = mg.grid_deliver(dist, Bt, col_groups, rank, pc, n, src=0, align=mg.PANEL_ALIGN) if pc > 1 else None
            return pans, mg.BroadcastMatmul(torch, dist, C, A, B, Bt, pans, src=0, collective=args.bcast, deliver=deliver)

        def step():
            A.touch()  # every step is a FRESH product: the cached 8-bit planes of A are rebuilt (B is external memory, never cached)
            if world == 1:
                g.mul_(C, A, B)
            else:
                bm.step()

        tune = {}
        choice = (pan_cands[0], cta_cands[0])
        if world > 1 and len(pan_cands) * len(cta_cands) > 1:  # untimed: pick (panels, GEMM grid cap) by a short trial of each
            for npan_c in pan_cands:
                panels, bm = make_bm(npan_c)
                for cc in cta_cands:
                    ctx.set_gemm_ctas(cc)
                    for _ in range(2):
                        step()
                    bm.finish(); barrier()
                    a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
                    a0.record(stream)
                    for _ in range(6):
                        step()
                    bm.finish(); a1.record(stream); barrier()
                    tt = torch.tensor([a0.elapsed_time(a1) / 6], dtype=torch.float64, device=f"cuda:{local}")
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    tune[(npan_c, cc)] = float(tt.item())
            choice = min(tune, key=tune.get)  # identical on every rank (all-reduced times)
        panels, bm = make_bm(choice[0])
        ctx.set_gemm_ctas(choice[1])
        npan = len(panels)
        pan = panels[0][1] - panels[0][0]

        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        for _ in range(W):
            step()
        barrier()
        if rank == 0:
            sampler.mark()
        ctx.set_profiling(True)
        l0 = ctx.launch_count()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(K):
            step()
        if bm is not None:
            bm.finish()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / K
        launches = ctx.launch_count() - l0
        phase_ms = ctx.last_timings()  # phases of the LAST timed step: [split, tcgen05 gemm, crt]
        clocks = sampler.stop() if rank == 0 else None
        ctx.set_profiling(False)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            lt = torch.tensor([launches], dtype=torch.int64, device=f"cuda:{local}")
            dist.all_reduce(lt, op=dist.ReduceOp.SUM)
            launches = int(lt.item())
        checksum = C.checksum()
        value = 2.0 * n ** 3 / (ms * 1e-3) / 1e9
        shard_ok = None
        if world > 1:  # every rank: the pipelined, broadcast-fed shard == the plain product of its row block with the B it received
            Cref = g.zeros(np.float32, mloc, ncl, N, ctx=ctx)
            g.mul_(Cref, A, B)
            same_c = C.equals(Cref)
            # ... and the B it received is the source's: rank 0 publishes the checksum of every column range
            src_sums = torch.zeros(pc, dtype=torch.int64, device=f"cuda:{local}")
            if rank == 0:
                for j in range(pc):
                    a, b = mg.col_range(n, pc, j, align=mg.PANEL_ALIGN)
                    Bj = Ball if pc == 1 else g.CuModMatrix.wrap_device(Bt.data_ptr() + 4 * a * ldb, n, b - a, ldb, N, ctx=ctx)
                    src_sums[j] = Bj.checksum() & 0x7FFFFFFFFFFFFFFF
            dist.broadcast(src_sums, src=0)
            same_b = int(src_sums[gj].item()) == (B.checksum() & 0x7FFFFFFFFFFFFFFF)
            ok = torch.tensor([1 if (same_c and same_b) else 0], dtype=torch.int32, device=f"cuda:{local}")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            shard_ok = bool(ok.item() == 1)
            del Cref

        # ---- parity of the TIMED result against the CPU oracle: sampled rows of this rank's C (the output of the last timed step),
        # recomputed by oracle_c.matmul_mod from rows of A rebuilt with the oracle's generator and the B this rank multiplied with
        # (downloaded; 32 of its columns are checked against the oracle's generator).  Reference criterion: `==` against the host
        # product, /root/reference/test/CuModMatrix/stripe_mul_test.jl:31-50.
        parity = None
        if not args.no_parity:
            from oracle import sampled as S
            nrows = max(8, args.parity_rows // world) if world > 1 else args.parity_rows
            rows_loc = S.pick_rows(mloc, nrows, seed=17 + rank)
            t0p = time.perf_counter()
            A_rows = S.synth_rows(SEED_A, rows_loc + r0, n, n, N)
            gen_ok = bool(np.array_equal(A.gather_rows(rows_loc), A_rows))
            Bh = B.to_u32()
            cols_s = S.pick_rows(ncl, 32, seed=99)
            gen_ok = gen_ok and bool(np.array_equal(Bh[:, cols_s].astype(np.int64), S.synth_cols(SEED_B, cols_s + cl0, n, N)))
            rep = S.check_product_rows(C.gather_rows(rows_loc), A_rows, Bh, N)
            del Bh
            ok_all, rows_all = rep["match"] and gen_ok, rep["rows"]
            if world > 1:
                tt = torch.tensor([1 if ok_all else 0, -rep["rows"]], dtype=torch.int64, device=f"cuda:{local}")
                dist.all_reduce(tt[0:1], op=dist.ReduceOp.MIN)
                rr = torch.tensor([rep["rows"]], dtype=torch.int64, device=f"cuda:{local}")
                dist.all_reduce(rr, op=dist.ReduceOp.SUM)
                ok_all, rows_all = bool(tt[0].item() == 1), int(rr.item())
            parity = {"rows": rows_all, "cols": rep["cols"], "match": bool(ok_all), "checker": "oracle_c.matmul_mod (exact uint64 host arithmetic)",
                      "what": "rows of the C produced by the last timed step" + ("" if world == 1 else " (every rank checks rows of its own shard)"),
                      "inputs_match_oracle_generator": gen_ok, "seconds": time.perf_counter() - t0p}

        # ---- roofline of the dominant kernel (tcgen05 GEMM): a few more profiled steps, kernel-only durations --------
        gemm_ms = [phase_ms[1]] if len(phase_ms) >= 2 else []
        ctx.set_profiling(True)
        for _ in range(3):
            if world == 1:
                g.mul_(C, A, B)
            else:
                g.capi.check(C.lib.gffm_gemm_block(C.h, 0, 0, A.h, 0, 0, B.h, 0, 0, mloc, pan, n, 0, 0, g.capi.GEMM_STORE, g.capi.ALGO_AUTO))
            pm = ctx.last_timings()
            if len(pm) >= 2:
                gemm_ms.append(pm[1])
        ctx.set_profiling(False)
        bits = (N - 1).bit_length()
        if N <= 256:
            units = 1
        elif N <= 65536:
            units = 4
        else:  # RNS: number of 8-bit moduli with product > 2*K*(N/2)^2
            need = 2 * n * (N // 2) ** 2
            need += (need >> 16) + 2
            mods = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197]
            prod, units = 1, 0
            while prod <= need:
                prod *= mods[units]; units += 1
        cols_per_launch = ncl if world == 1 else pan
        int8_ops = units * 2.0 * mloc * cols_per_launch * n
        # launch duration of the dominant kernel INSIDE the timed region (CUDA events of the last timed step): one launch per step
        # at N = 1; at N > 1 the per-panel launches of the last step (phase_ms = [0, sum of the launch durations, launches]).
        # The stand-alone launches after the region are kept as a second, informational figure.
        after_region = statistics.mean(gemm_ms[1:]) if len(gemm_ms) > 1 else None
        gemm_avg = None
        try:
            if world == 1 and len(phase_ms) >= 2 and phase_ms[1] > 0:
                gemm_avg = float(phase_ms[1])
            elif world > 1 and len(phase_ms) >= 3 and phase_ms[1] > 0 and phase_ms[2] >= 1:
                gemm_avg = float(phase_ms[1]) / float(phase_ms[2])
        except Exception:
            gemm_avg = None
        if gemm_avg is None:
            gemm_avg = after_region if after_region else (statistics.mean(gemm_ms) if gemm_ms else None)
        achieved = int8_ops / (gemm_avg * 1e-3) / 1e12 if gemm_avg else None
        int8_peak = 2.0 * peaks["bf16"]
        sustained_random = None
        peak_src = f"2 x bf16 dense {peaks['bf16']} TFLOP/s, {peaks['src']} (proxy: no int8 entry in MEASURED_PEAKS.json)"
        ip = os.path.join(ROOT, "profiles", "int8_peak.json")
        if os.path.exists(ip):  # measured by tools/int8_peak.cu on this pool's B200: back-to-back tcgen05 kind::i8 MMAs from resident smem
            try:
                d8 = json.load(open(ip))
                int8_peak = float(d8["int8_tops_sustained"])
                sustained_random = d8.get("int8_tops_sustained_random_data")
                peak_src = (f"measured int8 tensor-pipe peak {int8_peak} TOP/s (tools/int8_peak.cu: back-to-back tcgen05.mma kind::i8 128x256x32, operands "
                            f"resident in smem, all {d8.get('sms')} SMs, no HBM traffic so no power throttling); MEASURED_PEAKS.json has no int8 entry "
                            f"(its bf16 {peaks['bf16']} TFLOP/s x2 = {2*peaks['bf16']:.0f})")
            except Exception:
                pass
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if world == 1 and n == N_DEFAULT and os.path.exists(tp):  # the ncu capture is of the full single-GPU launch: meaningless for panel launches
            try:
                traffic = json.load(open(tp)).get("gemm_tc_kernel_rns_dram_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel<SchemeRNS>" if N > 65536 else "gemm_tc_kernel<limb>",
                    "achieved": achieved, "peak": int8_peak, "unit": "TOP/s (int8)", "frac": (achieved / int8_peak) if achieved else None,
                    "traffic": traffic, "launch_ms": gemm_avg, "launch_ms_standalone_after_region": after_region,
                    "int8_mma_units_per_k_step": units,
                    "peak_source": peak_src,
                    # same microbenchmark with uniformly random operand bytes, run back to back for seconds: what the tensor pipe
                    # sustains under the board power cap when NOTHING but MMAs runs (informational; frac uses the higher peak)
                    "peak_sustained_random_operands": sustained_random,
                    "frac_of_sustained_random": (achieved / sustained_random) if (achieved and sustained_random) else None,
                    "phases_ms_last_step": phase_ms}

        # ---- e2e through the public API with host buffers (rank-local shard; H2D A,B + GEMM + D2H C per step) --------
        e2e = None
        if not args.no_e2e:
            hA = torch.empty((n, mloc), dtype=torch.int32).pin_memory()   # column-major mloc x n
            hB = torch.empty((ncl, n), dtype=torch.int32).pin_memory()   # column-major n x ncl
            hC = torch.empty((ncl, mloc), dtype=torch.int32).pin_memory()
            g.capi.check(A.lib.gffm_mat_download(A.h, hA.data_ptr(), g.capi.U32, mloc, 0))
            g.capi.check(B.lib.gffm_mat_download(B.h, hB.data_ptr(), g.capi.U32, n, 0))
            A2 = g.zeros(np.float32, mloc, n, N, ctx=ctx); B2 = g.zeros(np.float32, n, ncl, N, ctx=ctx); C2 = g.zeros(np.float32, mloc, ncl, N, ctx=ctx)

            def e2e_step():
                g.capi.check(A2.lib.gffm_mat_upload(A2.h, hA.data_ptr(), g.capi.U32, mloc, 1))
                g.capi.check(B2.lib.gffm_mat_upload(B2.h, hB.data_ptr(), g.capi.U32, n, 1))
                g.mul_(C2, A2, B2)
                g.capi.check(C2.lib.gffm_mat_download(C2.h, hC.data_ptr(), g.capi.U32, mloc, 0))

            e2e_step()
            ke = max(1, min(K, 3))
            barrier()
            t0 = time.perf_counter()
            for _ in range(ke):
                e2e_step()
            barrier()
            te = (time.perf_counter() - t0) / ke
            if world > 1:
                t = torch.tensor([te], dtype=torch.float64, device=f"cuda:{local}")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                te = float(t.item())
            same = bool(C2.equals(C))
            # the same product through the single pipelined host-to-host call (gffm_gemm_host): H2D, plane split, GEMM tiles and
            # D2H overlap on three streams.  Only at N == 1 GPU (one process owns the whole product).
            pipe = None
            if world == 1:
                hC2 = torch.empty((n, n), dtype=torch.int32).pin_memory()

                def pipe_step():
                    g.capi.check(ctx.lib.gffm_gemm_host(ctx.h, hC2.data_ptr(), n, hA.data_ptr(), n, hB.data_ptr(), n, n, n, n, g.capi.U32, N))

                pipe_step()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(ke):
                    pipe_step()
                torch.cuda.synchronize()
                tp = (time.perf_counter() - t0) / ke
                pipe = {"ms_per_step": tp * 1e3, "GOPS": 2.0 * n ** 3 / tp / 1e9, "matches": bool(torch.equal(hC2, hC))}
            seq = {"ms_per_step": te * 1e3, "GOPS": 2.0 * n ** 3 / te / 1e9}
            if pipe and pipe["matches"] and pipe["ms_per_step"] < te * 1e3:
                te = pipe["ms_per_step"] / 1e3
            e2e = {"value": 2.0 * n ** 3 / te / 1e9, "unit": "GOPS", "api": "gffm_gemm_host (pipelined)" if pipe and te * 1e3 == pipe["ms_per_step"] else "upload + mul! + download",
                   "sequential_api_calls": seq, "pipelined_host_call": pipe, "h2d_bytes_per_step": int(4 * (mloc * n + n * ncl)), "d2h_bytes_per_step": int(4 * mloc * ncl),
                   "ms_per_step": te * 1e3, "steps": ke, "host_dtype": "uint32 residues (pinned)", "matches_resident_result": same}
            if world > 1 and pc == 1 and args.e2e_mode == "sliced":
                try:
                    C3 = g.zeros(np.float32, mloc, n, N, ctx=ctx)
                    Bt.zero_()  # nothing of B is resident any more: every byte must come from the host slices + the all-gather

                    def deliver_h(c0, c1):
                        rows = c1 - c0
                        if rows % world == 0:
                            per = rows // world
                            lo = c0 + rank * per
                            Bt[lo:lo + per, :n].copy_(hB[lo:lo + per], non_blocking=True)   # H2D: this rank's slice of the panel
                            dist.all_gather_into_tensor(Bt[c0:c1], Bt[lo:lo + per])         # NVLink: the other N-1 slices, in place
                        else:
                            if rank == 0:
                                Bt[c0:c1, :n].copy_(hB[c0:c1], non_blocking=True)
                            dist.broadcast(Bt[c0:c1], src=0)

                    bm_h = mg.BroadcastMatmul(torch, dist, C3, A2, B, Bt, panels, deliver=deliver_h)

                    def sliced_step():
                        g.capi.check(A2.lib.gffm_mat_upload(A2.h, hA.data_ptr(), g.capi.U32, mloc, 1))
                        bm_h.step()
                        g.capi.check(C3.lib.gffm_mat_download(C3.h, hC.data_ptr(), g.capi.U32, mloc, 0))

                    sliced_step()
                    bm_h.finish(); barrier()
                    t0 = time.perf_counter()
                    for _ in range(ke):
                        sliced_step()
                    bm_h.finish(); barrier()
                    ts = (time.perf_counter() - t0) / ke
                    tt = torch.tensor([ts], dtype=torch.float64, device=f"cuda:{local}")
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    ts = float(tt.item())
                    okk = torch.tensor([1 if C3.equals(C) else 0], dtype=torch.int32, device=f"cuda:{local}")
                    dist.all_reduce(okk, op=dist.ReduceOp.MIN)
                    e2e["sliced_upload_allgather"] = {"ms_per_step": ts * 1e3, "GOPS": 2.0 * n ** 3 / ts / 1e9, "matches_resident_result": bool(okk.item() == 1),
                                                      "h2d_bytes_per_step": int(4 * (mloc * n + n * n // world)), "d2h_bytes_per_step": int(4 * mloc * n)}
                    if okk.item() == 1 and ts < te:
                        e2e.update({"value": 2.0 * n ** 3 / ts / 1e9, "ms_per_step": ts * 1e3, "api": "upload A shard + 1/N of B, NCCL all-gather, gffm_gemm_panels, download",
                                    "h2d_bytes_per_step": int(4 * (mloc * n + n * n // world))})
                    del C3
                except Exception as ex:  # opt-in arm: never lose the line
                    e2e["sliced_upload_allgather"] = {"error": str(ex)[:300]}
            del A2, B2, C2

        extras = {}
        roofline_pluq = None
        if world == 1 and not args.no_extras:
            def dev_time(fn, reps, warm=1):
                """median / min wall time (s) of fn() bracketed by device synchronisation (the elimination calls block on their own)"""
                for _ in range(warm):
                    fn()
                ts = []
                for _ in range(reps):
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    fn()
                    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
                return statistics.median(ts), min(ts)

            def ev_time(fn, reps, warm=2):
                """mean device time (ms) of fn() over `reps` back-to-back calls (CUDA events on the library stream)"""
                for _ in range(warm):
                    fn()
                a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
                a0.record(stream)
                for _ in range(reps):
                    fn()
                a1.record(stream); torch.cuda.synchronize()
                return a0.elapsed_time(a1) / reps

            # other moduli of the metric size (fresh product each time)
            for N2 in (11, 65521):
                A_, B_ = g.synth(n, n, N2, SEED_A, ctx=ctx), g.synth(n, n, N2, SEED_B, ctx=ctx)
                C_ = g.zeros(np.float32, n, n, N2, ctx=ctx)

                def prod():
                    A_.touch(); B_.touch()
                    g.mul_(C_, A_, B_)
                t_ = ev_time(prod, 5, warm=3)
                extras[f"matmul_n{n}_mod{N2}"] = {"ms": t_, "GOPS": 2.0 * n ** 3 / t_ / 1e6}
                del A_, B_, C_
            # GEMV at the metric size: HBM-bound, 4 bytes per matrix element
            z_ = g.zeros(np.float32, n, 1, N, ctx=ctx); x_ = g.synth(n, 1, N, 77, ctx=ctx)
            t_ = ev_time(lambda: g.gemv_(z_, A, x_), 20, warm=3)
            extras[f"gemv_n{n}_mod{N}"] = {"ms": t_, "GBps": 4.0 * n * n / (t_ * 1e-3) / 1e9, "hbm_peak_GBps": peaks["hbm_gbs"],
                                            "frac_of_hbm_peak": 4.0 * n * n / (t_ * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": 4 * n * n,
                                            "note": "A (1 GiB) is larger than L2; 20 back-to-back products"}
            del z_, x_
            # the reference's only published cases (test/CuModMatrix/timing_test.jl:21-52: n = 5000, N = 11; RTX 3070 comments
            # < 0.001 s add!, < 0.001 s scalar mul!, < 0.2 s mul!, < 0.001 s mat-vec) from our path
            n5 = 5000
            A5, B5 = g.synth(n5, n5, 11, 1, ctx=ctx), g.synth(n5, n5, 11, 2, ctx=ctx)
            C5 = g.zeros(np.float32, n5, n5, 11, ctx=ctx); z5 = g.zeros(np.float32, n5, 1, 11, ctx=ctx); x5 = g.synth(n5, 1, 11, 3, ctx=ctx)

            def mul5():
                A5.touch(); B5.touch()
                g.mul_(C5, A5, B5)
            extras["reference_timing_cases_n5000_mod11"] = {
                "source": "reference test/CuModMatrix/timing_test.jl:21-52 (author's RTX 3070 comments: add! < 1 ms, scalar mul! < 1 ms, mul! < 200 ms (F32) / < 1000 ms (F64), mat-vec < 1 ms)",
                "add_ms": ev_time(lambda: g.add_(C5, A5, B5), 20), "scalar_mul_ms": ev_time(lambda: g.capi.check(C5.lib.gffm_ewise(g.capi.EW_SMUL, C5.h, A5.h, None, 2, 0)), 20),
                "mul_ms": ev_time(mul5, 10), "matvec_ms": ev_time(lambda: g.gemv_(z5, A5, x5), 20)}
            del A5, B5, C5, z5, x5
            # elimination at the metric size: PLUQ, RREF, inverse; warm, 3 repetitions, median (BASELINE metric: "PLUQ n=16384 s")
            for Np in (65521, N):
                A_ = g.synth(n, n, Np, 9, ctx=ctx)
                holder = {}

                def do_pluq():
                    holder["r"] = g.pluq_gpu_kernel(A_, return_rank=True)
                l0_ = ctx.launch_count()
                med, best = dev_time(do_pluq, 3)
                launches_pluq = (ctx.launch_count() - l0_) // 4
                rk_ = holder["r"][4]
                holder.clear()
                med_r, best_r = dev_time(lambda: holder.__setitem__("r", g.rref(A_)), 3)
                holder.clear()
                med_i, best_i = dev_time(lambda: holder.__setitem__("r", g.inverse(A_)), 3)
                holder.clear()
                ctx.set_profiling(True)
                do_pluq(); ph = ctx.last_timings(); holder.clear()
                ctx.set_profiling(False)
                Lb = 1 if Np <= 256 else (2 if Np <= 65536 else None)
                extras[f"pluq_n{n}_mod{Np}"] = {"seconds": med, "seconds_best": best, "rank": rk_, "rref_seconds": med_r, "rref_seconds_best": best_r,
                                                 "inverse_seconds": med_i, "inverse_seconds_best": best_i, "us_per_pivot": med / max(rk_, 1) * 1e6,
                                                 "kernel_launches_per_factorisation": int(launches_pluq),
                                                 "phases_ms_profiled_call": {"panels_and_in_block_updates": ph[0] if len(ph) > 0 else None,
                                                                             "triangular_solves": ph[1] if len(ph) > 1 else None, "schur_gemms": ph[2] if len(ph) > 2 else None},
                                                 "reps": 3, "timing": "median wall time between device synchronisations, warm (second call onwards)"}
                if Np == 65521:
                    fm = n ** 3 / 3.0  # m n r - (m + n) r^2 / 2 + r^3 / 3 for m = n = r
                    ip8 = 4542.2
                    try:
                        ip8 = float(json.load(open(os.path.join(ROOT, "profiles", "int8_peak.json")))["int8_tops_sustained"])
                    except Exception:
                        pass
                    peak_fm = ip8 * 1e12 / (2.0 * Lb * Lb)  # field mul-adds/s if all of them ran as L^2 int8 MMAs at the tensor peak
                    roofline_pluq = {"bound": "latency (n sequential pivots; the Schur updates are tensor-bound)", "kernel": "pluq_panel_ll_kernel",
                                     "workload": f"PLUQ {n}x{n} mod {Np}, full rank", "field_muladds": fm, "achieved": fm / med / 1e12, "peak": peak_fm / 1e12,
                                     "unit": "T field mul-add/s", "frac": fm / med / peak_fm, "seconds": med, "us_per_pivot": med / max(rk_, 1) * 1e6,
                                     "note": "peak = measured int8 tensor peak / (2 L^2), L = 2 eight-bit limbs; the panel kernel (one thread-block cluster, "
                                             "argmax + row exchange per pivot) bounds the factorisation, see DESIGN.md section 6"}
                del A_

    cpu = None
    cpu_pluq = None
    if rank == 0 and not args.no_cpu:
        gops, cores, desc, dt = cpu_arm(n, N)
        cpu = {"value": gops, "unit": "GOPS", "cores": cores, "kind": "port", "sample": desc, "seconds": dt,
               "extrapolated_full_workload_seconds": 2.0 * n ** 3 / (gops * 1e9)}
        if world == 1 and not args.no_extras:
            ech = cpu_echelon_arm(65521)
            cpu_pluq = {"kind": "port", "cores": 1, "unit": "seconds", "sample": "oracle_c.echelon (reference pivot rule, scalar C) on full-rank synthetic n x n mod 65521",
                        "runs": ech, "value": ech[-1]["seconds"],
                        "extrapolated_n16384_seconds": ech[-1]["seconds"] * (16384.0 / ech[-1]["n"]) ** 3}

    if rank == 0:
        out = {
            "metric": "mod-p matmul effective GOPS (2n^3/s)", "value": value, "unit": "GOPS", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "s8 residue limbs, int32 accumulate (exact)",
            "data": "synthetic",
            "config": {"workload": f"{n}x{n} * {n}x{n} matmul mod {N} ({bits}-bit modulus), A,B resident as uint32 residues", "n": n, "modulus": N,
                       "encoding": "RNS int8 tcgen05" if N > 65536 else "positional int8 limbs tcgen05",
                       "sharding": "single GPU" if world == 1 else f"{pr} row blocks of A x {pc} column range(s) of B over {world} GPUs, B replicated from rank 0 by NCCL ({args.bcast}) every step in {npan} column panels on a communication stream, consumed by gffm_gemm_panels (split/CRT under the GEMM, next step's broadcast under this step's GEMMs)",
                       "shards_match_local_product_on_all_ranks": shard_ok,
                       "gemm_grid_cap": choice[1], "warmup_trials_ms": {f"panels={k[0]},gemm_ctas={k[1]}": round(v, 4) for k, v in tune.items()},
                       "l2_policy": f"inputs larger than L2: A and B are {4 * n * n / 2**20:.0f} MiB each vs 126 MB L2", "checksum_rank0": f"{checksum:016x}",
                       "extras": extras},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "parity_check": parity,
        }
        if roofline_pluq is not None:
            out["roofline_pluq"] = roofline_pluq
        if cpu_pluq is not None:
            out["cpu_baseline_pluq"] = cpu_pluq
        if world > 1 and os.environ.get("NCCL_DEBUG_FILE"):
            out["config"]["nccl_log"] = nccl_log_lines()
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
