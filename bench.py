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


def cpu_arm(n_full, N, budget_s=20.0, threads=None):
    """Times the C oracle (oracle/oracle_c.c, pthreads over all host cores) on a bounded sample of the workload:
    the leading n_s x n_s x n_s sub-product of the same synthetic matrices.  Returns (GOPS, cores, description, seconds)."""
    import numpy as np
    from oracle import oracle as O
    from oracle import oracle_c as OC
    cores = OC.num_threads() if threads is None else threads
    ns = 512
    A = O.synth_matrix(SEED_A, ns, ns, N); B = O.synth_matrix(SEED_B, ns, ns, N)
    t0 = time.perf_counter(); OC.matmul_mod(A, B, N); t_small = time.perf_counter() - t0
    rate = 2.0 * ns ** 3 / max(t_small, 1e-6)
    ns = 1024
    while ns * 2 <= min(n_full, 8192) and 2.0 * (2 * ns) ** 3 / rate < budget_s:
        ns *= 2
    A = O.synth_matrix(SEED_A, ns, ns, N); B = O.synth_matrix(SEED_B, ns, ns, N)
    t0 = time.perf_counter(); C = OC.matmul_mod(A, B, N); dt = time.perf_counter() - t0
    chk = int(np.bitwise_xor.reduce(C.reshape(-1)))
    gops = 2.0 * ns ** 3 / dt / 1e9
    return gops, cores, f"{ns}x{ns}x{ns} mod {N} sub-product of the n={n_full} workload, same generator (xor checksum {chk:#x})", dt


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
        gops, cores, desc, dt = cpu_arm(args.n, args.modulus, budget_s=15.0)
        gops_all.append(gops)
    value = statistics.median(gops_all)
    out = {
        "impl": "reference", "metric": "mod-p matmul effective GOPS (2n^3/s)", "value": value, "unit": "GOPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 2.0 * args.n ** 3 / (value * 1e9) * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64 accumulate of u32 residues (host integers)", "data": "synthetic",
        "config": {"workload": f"{args.n}x{args.n} * {args.n}x{args.n} matmul mod {args.modulus} ({(args.modulus - 1).bit_length()}-bit modulus), A,B resident as uint32 residues",
                   "n": args.n, "modulus": args.modulus,
                   "note": "the Julia reference cannot run in this image; this arm is the reference tests' CPU ground truth mod.(A*B,N) restated in C "
                           "(oracle/oracle_c.c) on all host cores; ms_per_step is the n^3-extrapolated time of the full workload"},
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
    ap.add_argument("--extras", action="store_true", help="also time N=11 and N=65521 and PLUQ (reported under config.extras)")
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
        os.environ.pop("NCCL_DEBUG", None)  # no "NCCL version" banner on stdout: ONE JSON line
        if os.environ.get("GFFM_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = os.environ["GFFM_NCCL_DEBUG"]
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
            for pc in pan_cands:
                panels, bm = make_bm(pc)
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
                    tune[(pc, cc)] = float(tt.item())
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
        if os.path.exists(tp):
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
        if args.extras and world == 1:
            for N2 in (11, 65521):
                A_, B_ = g.synth(n, n, N2, SEED_A, ctx=ctx), g.synth(n, n, N2, SEED_B, ctx=ctx)
                C_ = g.zeros(np.float32, n, n, N2, ctx=ctx)
                for _ in range(3):
                    g.mul_(C_, A_, B_)
                a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
                a0.record(stream)
                for _ in range(5):
                    A_.touch(); B_.touch()  # fresh product each time (no cached planes)
                    g.mul_(C_, A_, B_)
                a1.record(stream); torch.cuda.synchronize()
                t_ = a0.elapsed_time(a1) / 5
                extras[f"matmul_n{n}_mod{N2}"] = {"ms": t_, "GOPS": 2.0 * n ** 3 / t_ / 1e6}
                del A_, B_, C_
            for (np_, Np) in ((n, 65521), (n, N)):
                A_ = g.synth(np_, np_, Np, 9, ctx=ctx)
                tp_ = None
                for rep in range(2):  # first call pays one-off costs (workspaces, inverse table, kernel attributes)
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    U_, L_, pr_, pc_, rk_ = g.pluq_gpu_kernel(A_, return_rank=True)
                    torch.cuda.synchronize(); tp_ = time.perf_counter() - t0
                    del U_, L_
                for rep in range(2):
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    R_ = g.rref(A_)
                    torch.cuda.synchronize(); tr_ = time.perf_counter() - t0
                extras[f"pluq_n{np_}_mod{Np}"] = {"seconds": tp_, "rank": rk_, "rref_seconds": tr_}
                del A_, R_

    cpu = None
    if rank == 0 and not args.no_cpu:
        gops, cores, desc, dt = cpu_arm(n, N)
        cpu = {"value": gops, "unit": "GOPS", "cores": cores, "kind": "port", "sample": desc, "seconds": dt}

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
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
