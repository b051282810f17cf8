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


CPU_SAMPLE_N = 8192  # fixed CPU sample (the same sub-product on every host, so ratios are comparable across runs and GPU counts)


def cpu_arm(n_full, N, ns=CPU_SAMPLE_N, threads=None):
    """Times the C oracle (oracle/oracle_c.c, pthreads over all host cores) on a FIXED bounded sample of the workload: the leading
    ns x ns x ns sub-product of the same synthetic matrices (ns = 8192: ~5-15 s on 16-32 cores with the blocked kernel).  Returns (GOPS, cores, description,
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
                           "(oracle/oracle_c.c) on all host cores; every step is the FIXED 8192^3 sub-product (same on every host); "
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


def run_extras(g, ctx, torch, np, stream, A, n, N, peaks):
    """config.extras of the single-GPU line: other moduli, GEMV, the reference's n = 5000 timing cases, PLUQ / RREF / inverse at the metric
    size with the PLUQ roofline entry (BASELINE metric: "...; PLUQ n=16384 s")."""
    extras = {}
    roofline_pluq = None
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

    return extras, roofline_pluq


def int8_units(n, R, P, kara=False):
    """int8 MMA units per inner-dimension step of one product with inputs < R (1 / 4 for one / two positional limbs, else the RNS modulus count)."""
    if not kara and R <= 256:
        return 1
    if not kara and R <= 65536:
        return 4
    need = (n * (R - 1) ** 2) if kara else 2 * n * (R // 2) ** 2
    need += (need >> 6) + 2
    mods = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197]
    prod, units = 1, 0
    while prod <= need:
        prod *= mods[units]; units += 1
    return units


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--n", type=int, default=N_DEFAULT)
    ap.add_argument("--modulus", type=int, default=MOD_DEFAULT)
    ap.add_argument("--workload", default="matmul", choices=["matmul", "karatsuba"],
                    help="matmul: C = A*B mod N (the BASELINE metric); karatsuba: two-limb product mod N1*N2 (BASELINE config 5 with --n 32768)")
    ap.add_argument("--n1", type=int, default=8191, help="Karatsuba limb modulus N1 (SURVEY 8d: N1 = N2 = 8191)")
    ap.add_argument("--n2", type=int, default=8191)
    ap.add_argument("--mg", default="cabi", choices=["cabi", "python"],
                    help="N > 1: 'cabi' = the library's multi-GPU layer (gffm_mg_gemm / gffm_mg_kmat_mul, csrc/mg.cu); 'python' = the round-1 driver "
                         "(torch.distributed broadcast of B in panels + gffm_gemm_panels), kept for A/B comparisons")
    ap.add_argument("--transport", default="tune",
                    help="N > 1, --mg cabi: p2p_push | p2p_planes | nccl_planes | nccl_bcast | auto, or 'tune' (default): each available transport is tried for a few "
                         "untimed steps and the fastest is used for the timed region (trial times in config.warmup_trials_ms)")
    ap.add_argument("--panels", type=int, default=8, help="--mg python: column panels of B per broadcast")
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

    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # the multi-GPU layer's streams must not share hardware queues (flag waits)
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
            # one file set per run (world size + rendezvous port in the name): runs at several N on one box must not read each other's logs
            tag = f"w{world}.p{os.environ.get('MASTER_PORT', '0')}"
            os.environ["NCCL_DEBUG_FILE"] = os.path.join(ROOT, "gpurun_out", f"nccl_debug.{tag}.%h.%p.log")
            if rank == 0:  # leftovers of an earlier run with the same tag; NCCL is initialised only after rank 0 has handed out its id
                import glob
                for f_ in glob.glob(os.environ["NCCL_DEBUG_FILE"].replace("%h", "*").replace("%p", "*")):
                    try:
                        os.remove(f_)
                    except OSError:
                        pass
        # torch.distributed is plumbing only (id exchange, barriers, max over ranks): the gloo (CPU) backend.  The data path's NCCL
        # communicator lives inside the library (gffm_mg_create).  Not using torch's NCCL backend / stream pool also keeps the process below
        # CUDA_DEVICE_MAX_CONNECTIONS streams, so the multi-GPU layer's streams never share a hardware queue (profiles/r02_notes.md).
        dist.init_process_group("nccl" if args.mg == "python" else "gloo", **({"device_id": torch.device("cuda", local)} if args.mg == "python" else {}))
    n, N = args.n, args.modulus
    kara = args.workload == "karatsuba"
    N1, N2 = args.n1, args.n2
    W = max(3, args.warmup)
    K = max(1, args.steps)
    g.set_default_device(local)
    ctx = g.Context(local)
    if world > 1 and args.mg == "python":
        stream = torch.cuda.Stream(device=local)
        ctx.set_stream(stream.cuda_stream)
    else:
        stream = torch.cuda.ExternalStream(ctx.get_stream(), device=local)  # the library's own stream (torch only records events on it)
    peaks = load_peaks()
    dev = f"cuda:{local}"
    rdev = dev if (world > 1 and args.mg == "python") else "cpu"  # where the reduction tensors of the plumbing live

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=rdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allmin_flag(ok):
        if world == 1:
            return bool(ok)
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=rdev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() == 1)

    with torch.cuda.stream(stream):
        mg = g.multigpu
        r0, r1 = mg.row_block(n, world, rank)
        mloc = r1 - r0
        ldb = ((n + 31) // 32) * 32

        def shard_of(seed, modulus):
            """this rank's row block of the synthetic n x n matrix (device generator, checked against the oracle's by parity_check)"""
            if world == 1:
                return g.synth(n, n, modulus, seed, ctx=ctx)
            full = g.synth(n, n, modulus, seed, ctx=ctx)
            sh = g.zeros(np.float32, mloc, n, modulus, ctx=ctx)
            g.capi.check(sh.lib.gffm_mat_copy_block(sh.h, 0, 0, full.h, r0, 0, mloc, n))
            ctx.sync()
            del full
            return sh

        def b_matrix(seed, modulus):
            """B: column-major n x n in a torch buffer (so that torch.distributed can replicate it for the verification); data on rank 0 only"""
            t = torch.zeros((n, ldb), dtype=torch.int32, device=dev)
            m_ = g.CuModMatrix.wrap_device(t.data_ptr(), n, n, ldb, modulus, ctx=ctx)
            if rank == 0:
                src = g.synth(n, n, modulus, seed, ctx=ctx)
                g.copy_(m_, src)
                ctx.sync()
                del src
            return t, m_

        if not kara:
            A = shard_of(SEED_A, N)
            Bt, B = b_matrix(SEED_B, N)
            C = g.zeros(np.float32, mloc, n, N, ctx=ctx)
            touch_inputs = lambda: A.touch()  # noqa: E731  every step is a FRESH product: the cached 8-bit planes of A are rebuilt; B is external memory (never cached)
        else:
            A1 = shard_of(13, N1); A2 = shard_of(14, N2)
            B1t, B1 = b_matrix(15, N1); B2t, B2 = b_matrix(16, N2)
            AK = g.KaratsubaMatrix(A1, A2, N1, N2); BK = g.KaratsubaMatrix(B1, B2, N1, N2)
            CK = g.KaratsubaZeros(np.float64, mloc, n, N1, N2, ctx=ctx)
            C = CK.data1

            def touch_inputs():
                A1.touch(); A2.touch()
        torch.cuda.synchronize()
        b_ready = torch.cuda.Event()
        b_ready.record(stream)  # B is complete from here on: the distribution of step t+1 may run under the GEMMs of step t
        torch.cuda.synchronize()

        mgpu = None
        bm = None
        tune = {}
        transport_used = None
        if world > 1 and args.mg == "cabi":
            mgpu = mg.MultiGpu.from_torch_distributed(dist, ctx)

            def step():
                touch_inputs()
                if kara:
                    mgpu.kmat_mul(CK, AK, BK, root=0, b_ready=b_ready.cuda_event)
                else:
                    mgpu.gemm(C, A, B, root=0, b_ready=b_ready.cuda_event)

            names = {"p2p_raw": g.capi.MG_P2P_RAW, "p2p_push": g.capi.MG_P2P_PUSH, "p2p_planes": g.capi.MG_P2P_PLANES, "nccl_planes": g.capi.MG_NCCL_PLANES, "nccl_bcast": g.capi.MG_NCCL_BCAST,
                     "auto": g.capi.MG_AUTO}
            cands = ["p2p_raw", "p2p_push", "p2p_planes", "nccl_planes", "nccl_bcast"] if args.transport == "tune" else [t.strip() for t in args.transport.split(",")]
            best = None
            # reference result for the trials: the product through the NCCL broadcast transport (every rank splits all of B itself)
            outs = [CK.data1, CK.data2] if kara else [C]
            refs = None
            if len(cands) > 1:
                mgpu.set_transport(g.capi.MG_NCCL_BCAST)
                step()
                mgpu.barrier()
                refs = [g.zeros(np.float32, mloc, n, o.N, ctx=ctx) for o in outs]
                for rf, o in zip(refs, outs):
                    g.copy_(rf, o)
                ctx.sync()
            for tname in cands:
                ok_t = True
                try:
                    mgpu.set_transport(names[tname])
                    for _ in range(2):
                        step()
                    mgpu.barrier()
                except g.GffmError as ex:
                    ok_t = False
                    tune[tname] = f"unavailable: {str(ex)[:160]}"
                if not allmin_flag(ok_t):  # a transport is used only if every rank can use it
                    tune.setdefault(tname, "unavailable on another rank")
                    continue
                if refs is not None and not allmin_flag(all(o.equals(rf) for o, rf in zip(outs, refs))):  # ... and only if its product is bit-equal
                    tune[tname] = "rejected: result differs from the NCCL broadcast transport's"
                    continue
                if len(cands) == 1:
                    best = tname
                    break
                # Trial = 10 untimed + 12 timed steps.  Measured on 8 B200 (profiles/r02_notes.md): the first ~10 steps after an idle period
                # (set_transport drains everything and rebuilds the arena) run up to 10 % faster than the settled rate of the NCCL
                # broadcast transport, so an 8-step trial picked it at 4.75 ms and the 20 timed steps then ran at 5.07-5.18 ms.
                for _ in range(10):
                    step()
                barrier()
                a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
                a0.record(stream)
                for _ in range(12):
                    step()
                a1.record(stream); barrier()
                tune[tname] = allmax(a0.elapsed_time(a1) / 12)
                if best is None or tune[tname] < tune[best]:
                    best = tname
            del refs
            if best is None:
                raise SystemExit("no multi-GPU transport is usable: " + json.dumps(tune))
            # ties (within 3 %) go to the library's default transport: its trial times have been the reproducible ones (see above)
            pref = "p2p_push"
            if best != pref and isinstance(tune.get(pref), float) and tune[pref] <= 1.03 * tune[best]:
                best = pref
            mgpu.set_transport(names[best])
            transport_used = best
        elif world > 1:
            if kara:
                def step():
                    touch_inputs()
                    mg.sharded_kmat_mul(dist, B1t, B2t, lambda: g.KMatMul_(CK, AK, BK), src=0)
                transport_used = "python: dist.broadcast(B1, B2) then gffm_kmat_mul"
            else:
                panels = mg.col_panels(n, args.panels, align=mg.PANEL_ALIGN)
                bm = mg.BroadcastMatmul(torch, dist, C, A, B, Bt, panels, src=0)

                def step():
                    touch_inputs()
                    bm.step()
                transport_used = f"python: dist.broadcast in {len(panels)} panels + gffm_gemm_panels"
        else:
            def step():
                touch_inputs()
                if kara:
                    g.KMatMul_(CK, AK, BK)
                else:
                    g.mul_(C, A, B)

        def drain():
            if bm is not None:
                bm.finish()

        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        for _ in range(W):
            step()
        drain(); barrier()
        if rank == 0:
            sampler.mark()
        ctx.set_profiling(not kara)  # CUDA events around every tensor-core GEMM launch of the timed steps (the Karatsuba product has three kinds)
        l0 = ctx.launch_count()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(K):
            step()
        drain()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / K
        launches = ctx.launch_count() - l0
        phase_ms = ctx.last_timings() if not kara else []  # phases of the LAST timed step: [split, tcgen05 gemm, crt] or [0, sum of GEMM launches, launches]
        clocks = sampler.stop() if rank == 0 else None
        ctx.set_profiling(False)
        mg_error = None
        if mgpu is not None:
            try:
                mgpu.barrier()
            except g.GffmError as ex:
                mg_error = str(ex)[:300]
        ms = allmax(ms)
        if world > 1:
            lt = torch.tensor([launches], dtype=torch.int64, device=rdev)
            dist.all_reduce(lt, op=dist.ReduceOp.SUM)
            launches = int(lt.item())
        checksum = C.checksum()
        value = 2.0 * n ** 3 / (ms * 1e-3) / 1e9

        # ---- every rank: the sharded result == the single-GPU product of its row block with the whole B (replicated here for the check) ----
        shard_ok = None
        if world > 1:
            def replicate(*ts):
                """B on every rank (for the checks below)"""
                if mgpu is None:
                    for t_ in ts:
                        dist.broadcast(t_, src=0)
            if kara:
                CR = g.KaratsubaZeros(np.float64, mloc, n, N1, N2, ctx=ctx)
                same_t = True
                if mgpu is not None:  # the same product through the OTHER data path (NCCL broadcast of the uint32 limbs; replicates B1, B2 as a side effect)
                    mgpu.set_transport(g.capi.MG_NCCL_BCAST)
                    mgpu.kmat_mul(CR, AK, BK, root=0)
                    mgpu.barrier()
                    same_t = CK.data1.equals(CR.data1) and CK.data2.equals(CR.data2)
                replicate(B1t, B2t)
                torch.cuda.synchronize()
                B1.touch(); B2.touch()
                g.KMatMul_(CR, AK, BK)  # ... and through the single-GPU entry point on this rank's row block
                same_c = same_t and CK.data1.equals(CR.data1) and CK.data2.equals(CR.data2)
                del CR
            else:
                Cref = g.zeros(np.float32, mloc, n, N, ctx=ctx)
                same_t = True
                if mgpu is not None:
                    mgpu.set_transport(g.capi.MG_NCCL_BCAST)
                    mgpu.gemm(Cref, A, B, root=0)
                    mgpu.barrier()
                    same_t = C.equals(Cref)
                replicate(Bt)
                torch.cuda.synchronize()
                B.touch()
                g.mul_(Cref, A, B)
                same_c = same_t and C.equals(Cref)
                del Cref
            shard_ok = allmin_flag(same_c and mg_error is None)
            if mgpu is not None:
                mgpu.set_transport(names[transport_used])  # back to the transport of the timed region (the end-to-end arm below uses it)

        # ---- parity of the TIMED result against the CPU oracle: sampled rows of this rank's C (the output of the last timed step),
        # recomputed by oracle_c.matmul_mod from rows of A rebuilt with the oracle's generator and the B held by rank 0 (downloaded;
        # 32 of its columns are checked against the oracle's generator).  Reference criterion: `==` against the host product,
        # /root/reference/test/CuModMatrix/stripe_mul_test.jl:31-50.
        parity = None
        if not args.no_parity:
            from oracle import sampled as S
            nrows = max(8, args.parity_rows // world) if world > 1 else args.parity_rows
            rows_loc = S.pick_rows(mloc, nrows, seed=17 + rank)
            t0p = time.perf_counter()
            cols_s = S.pick_rows(n, 32, seed=99)
            if not kara:
                A_rows = S.synth_rows(SEED_A, rows_loc + r0, n, n, N)
                gen_ok = bool(np.array_equal(A.gather_rows(rows_loc), A_rows))
                Bh = B.to_u32()  # every rank holds B after the shard check (world > 1) / owns it (world == 1)
                gen_ok = gen_ok and bool(np.array_equal(Bh[:, cols_s].astype(np.int64), S.synth_cols(SEED_B, cols_s, n, N)))
                rep = S.check_product_rows(C.gather_rows(rows_loc), A_rows, Bh, N)
            else:
                M_ = N1 * N2
                A_rows = S.synth_rows(13, rows_loc + r0, n, n, N1) + N1 * S.synth_rows(14, rows_loc + r0, n, n, N2)
                gen_ok = bool(np.array_equal(A1.gather_rows(rows_loc) + N1 * A2.gather_rows(rows_loc), A_rows))
                B1h = B1.to_u32(); B2h = B2.to_u32()
                gen_ok = gen_ok and bool(np.array_equal(B1h[:, cols_s].astype(np.int64), S.synth_cols(15, cols_s, n, N1)))
                gen_ok = gen_ok and bool(np.array_equal(B2h[:, cols_s].astype(np.int64), S.synth_cols(16, cols_s, n, N2)))
                Bh = B1h + np.uint32(N1) * B2h
                del B1h, B2h
                rep = S.check_product_rows(CK.data1.gather_rows(rows_loc) + N1 * CK.data2.gather_rows(rows_loc), A_rows, Bh, M_, in_bound=M_)
            del Bh
            ok_all, rows_all = rep["match"] and gen_ok, rep["rows"]
            if world > 1:
                ok_all = allmin_flag(ok_all)
                rr = torch.tensor([rep["rows"]], dtype=torch.int64, device=rdev)
                dist.all_reduce(rr, op=dist.ReduceOp.SUM)
                rows_all = int(rr.item())
            parity = {"rows": rows_all, "cols": rep["cols"], "match": bool(ok_all), "checker": "oracle_c.matmul_mod (exact uint64 host arithmetic)",
                      "what": "rows of the C produced by the last timed step" + ("" if world == 1 else " (every rank checks rows of its own shard)"),
                      "inputs_match_oracle_generator": gen_ok, "seconds": time.perf_counter() - t0p}

        # ---- roofline of the dominant kernel (tcgen05 GEMM) -------------------------------------------------------------------------
        if kara:
            units = int8_units(n, N1, N1 * N2, kara=True) + int8_units(n, N1 + N2, N2) + int8_units(n, N2, N2)
        else:
            units = int8_units(n, N, N)
        per_rank_ops = units * 2.0 * mloc * n * n
        gemm_ms_step = None   # device time of the tensor-core GEMM launches of ONE step on this rank
        n_gemm_launches = None
        try:
            if not kara and world == 1 and len(phase_ms) >= 2 and phase_ms[1] > 0:
                gemm_ms_step, n_gemm_launches = float(phase_ms[1]), 1
            elif not kara and len(phase_ms) >= 3 and phase_ms[1] > 0 and phase_ms[2] >= 1:
                gemm_ms_step, n_gemm_launches = float(phase_ms[1]), int(phase_ms[2])
        except Exception:
            gemm_ms_step = None
        after_region = None
        if not kara and world == 1:  # stand-alone launches after the region: second, informational figure
            ctx.set_profiling(True)
            ts_ = []
            for _ in range(3):
                g.mul_(C, A, B)
                pm = ctx.last_timings()
                if len(pm) >= 2:
                    ts_.append(pm[1])
            ctx.set_profiling(False)
            after_region = statistics.mean(ts_) if ts_ else None
        if gemm_ms_step is None and kara:
            gemm_ms_step = ms  # the three sub-product GEMMs are ~all of the step; per-kernel split in profiles/
        achieved = per_rank_ops / (gemm_ms_step * 1e-3) / 1e12 if gemm_ms_step else None
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
        if world == 1 and n == N_DEFAULT and not kara and os.path.exists(tp):  # the ncu capture is of the full single-GPU launch
            try:
                traffic = json.load(open(tp)).get("gemm_tc_kernel_rns_dram_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel<SchemeRNS>" if (kara or N > 65536) else "gemm_tc_kernel<limb>",
                    "achieved": achieved, "peak": int8_peak, "unit": "TOP/s (int8)", "frac": (achieved / int8_peak) if achieved else None,
                    "traffic": traffic, "launch_ms": (gemm_ms_step / n_gemm_launches) if (gemm_ms_step and n_gemm_launches) else gemm_ms_step,
                    "gemm_launches_per_step_per_rank": n_gemm_launches, "gemm_ms_per_step_per_rank": gemm_ms_step,
                    "launch_ms_standalone_after_region": after_region, "int8_mma_units_per_k_step": units,
                    "achieved_is": "per GPU (rank 0): executed int8 ops of this rank's GEMM launches of the last timed step / their CUDA-event durations",
                    "peak_source": peak_src,
                    "peak_sustained_random_operands": sustained_random,
                    "frac_of_sustained_random": (achieved / sustained_random) if (achieved and sustained_random) else None,
                    "phases_ms_last_step": phase_ms}

        # ---- e2e through the public API with HOST buffers (pinned): per step H2D of the inputs, the product, D2H of the result ----------
        e2e = None
        if not args.no_e2e and not kara:
            ke = max(1, min(K, 3))
            if world == 1:
                hA = torch.empty((n, n), dtype=torch.int32).pin_memory()
                hB = torch.empty((n, n), dtype=torch.int32).pin_memory()
                hC = torch.empty((n, n), dtype=torch.int32).pin_memory()
                g.capi.check(A.lib.gffm_mat_download(A.h, hA.data_ptr(), g.capi.U32, n, 0))
                g.capi.check(B.lib.gffm_mat_download(B.h, hB.data_ptr(), g.capi.U32, n, 0))
                A2_ = g.zeros(np.float32, n, n, N, ctx=ctx); B2_ = g.zeros(np.float32, n, n, N, ctx=ctx); C2_ = g.zeros(np.float32, n, n, N, ctx=ctx)

                def e2e_step():
                    g.capi.check(A2_.lib.gffm_mat_upload(A2_.h, hA.data_ptr(), g.capi.U32, n, 1))
                    g.capi.check(B2_.lib.gffm_mat_upload(B2_.h, hB.data_ptr(), g.capi.U32, n, 1))
                    g.mul_(C2_, A2_, B2_)
                    g.capi.check(C2_.lib.gffm_mat_download(C2_.h, hC.data_ptr(), g.capi.U32, n, 0))

                e2e_step()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(ke):
                    e2e_step()
                torch.cuda.synchronize()
                te = (time.perf_counter() - t0) / ke
                same = bool(C2_.equals(C))
                hC2 = torch.empty((n, n), dtype=torch.int32).pin_memory()

                def pipe_step():  # the same product through the single pipelined host-to-host call: H2D, plane split, GEMM tiles and D2H overlap
                    g.capi.check(ctx.lib.gffm_gemm_host(ctx.h, hC2.data_ptr(), n, hA.data_ptr(), n, hB.data_ptr(), n, n, n, n, g.capi.U32, N))

                pipe_step()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(ke):
                    pipe_step()
                torch.cuda.synchronize()
                tp_ = (time.perf_counter() - t0) / ke
                pipe = {"ms_per_step": tp_ * 1e3, "GOPS": 2.0 * n ** 3 / tp_ / 1e9, "matches": bool(torch.equal(hC2, hC))}
                seq = {"ms_per_step": te * 1e3, "GOPS": 2.0 * n ** 3 / te / 1e9}
                use_pipe = pipe["matches"] and pipe["ms_per_step"] < te * 1e3
                tbest = tp_ if use_pipe else te
                e2e = {"value": 2.0 * n ** 3 / tbest / 1e9, "unit": "GOPS", "api": "gffm_gemm_host (pipelined)" if use_pipe else "upload + mul! + download",
                       "sequential_api_calls": seq, "pipelined_host_call": pipe, "h2d_bytes_per_step": int(8 * n * n), "d2h_bytes_per_step": int(4 * n * n),
                       "ms_per_step": tbest * 1e3, "steps": ke, "host_dtype": "uint32 residues (pinned)", "matches_resident_result": same}
                del A2_, B2_, C2_
            elif mgpu is not None:
                # every rank uploads its row block of A and ITS OWN column range of B over its own PCIe link (B distributed: root =
                # GFFM_MG_DISTRIBUTED), the operand planes are exchanged over NVLink, every rank downloads its row block of C
                off = mg.owner_ranges(n, world)
                c0, c1 = off[rank], off[rank + 1]
                hA = torch.empty((n, mloc), dtype=torch.int32).pin_memory()            # column-major mloc x n
                hB = torch.empty((max(c1 - c0, 1), n), dtype=torch.int32).pin_memory()  # column-major n x (c1 - c0)
                hC = torch.empty((n, mloc), dtype=torch.int32).pin_memory()
                g.capi.check(A.lib.gffm_mat_download(A.h, hA.data_ptr(), g.capi.U32, mloc, 0))
                Bq = g.zeros(np.float32, n, c1 - c0, N, ctx=ctx)
                if c1 > c0:
                    g.capi.check(Bq.lib.gffm_mat_copy_block(Bq.h, 0, 0, B.h, 0, c0, n, c1 - c0))  # B was replicated for the shard check above
                    g.capi.check(Bq.lib.gffm_mat_download(Bq.h, hB.data_ptr(), g.capi.U32, n, 0))
                A2_ = g.zeros(np.float32, mloc, n, N, ctx=ctx); C2_ = g.zeros(np.float32, mloc, n, N, ctx=ctx)
                g.fill_(Bq, 0)

                def e2e_step():
                    g.capi.check(A2_.lib.gffm_mat_upload(A2_.h, hA.data_ptr(), g.capi.U32, mloc, 1))
                    if c1 > c0:
                        g.capi.check(Bq.lib.gffm_mat_upload(Bq.h, hB.data_ptr(), g.capi.U32, n, 1))
                    mgpu.gemm(C2_, A2_, Bq, root=g.capi.MG_DISTRIBUTED)
                    g.capi.check(C2_.lib.gffm_mat_download(C2_.h, hC.data_ptr(), g.capi.U32, mloc, 0))

                err = None
                try:
                    e2e_step()
                    barrier()
                    t0 = time.perf_counter()
                    for _ in range(ke):
                        e2e_step()
                    barrier()
                    te = allmax((time.perf_counter() - t0) / ke)
                    mgpu.barrier()
                    same = allmin_flag(bool(C2_.equals(C)))
                except g.GffmError as ex:
                    err = str(ex)[:300]
                    te, same = None, False
                tot = torch.tensor([4 * (mloc * n + n * (c1 - c0)), 4 * mloc * n], dtype=torch.int64, device=rdev)
                dist.all_reduce(tot, op=dist.ReduceOp.SUM)
                e2e = {"value": (2.0 * n ** 3 / te / 1e9) if te else None, "unit": "GOPS",
                       "api": "per rank: gffm_mat_upload(A row block), gffm_mat_upload(own column range of B), gffm_mg_gemm(root = GFFM_MG_DISTRIBUTED), gffm_mat_download(C row block)",
                       "h2d_bytes_per_step": int(tot[0].item()), "d2h_bytes_per_step": int(tot[1].item()),
                       "h2d_bytes_per_step_per_rank": int(4 * (mloc * n + n * (c1 - c0))), "d2h_bytes_per_step_per_rank": int(4 * mloc * n),
                       "ms_per_step": te * 1e3 if te else None, "steps": ke, "host_dtype": "uint32 residues (pinned)", "matches_resident_result": same, "error": err}
                del A2_, C2_, Bq
            else:
                e2e = {"value": None, "unit": "GOPS", "note": "--mg python has no end-to-end arm (use the default --mg cabi)", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

        extras, roofline_pluq = ({}, None)
        if world == 1 and not args.no_extras and not kara:
            extras, roofline_pluq = run_extras(g, ctx, torch, np, stream, A, n, N, peaks)

    cpu = None
    cpu_pluq = None
    if rank == 0 and not args.no_cpu:
        gops, cores, desc, dt = cpu_arm(n, N if not kara else N1 * N2)
        cpu = {"value": gops, "unit": "GOPS", "cores": cores, "kind": "port", "sample": desc, "seconds": dt,
               "extrapolated_full_workload_seconds": 2.0 * n ** 3 / (gops * 1e9)}
        if world == 1 and not args.no_extras and not kara:
            ech = cpu_echelon_arm(65521)
            cpu_pluq = {"kind": "port", "cores": 1, "unit": "seconds", "sample": "oracle_c.echelon (reference pivot rule, scalar C) on full-rank synthetic n x n mod 65521",
                        "runs": ech, "value": ech[-1]["seconds"],
                        "extrapolated_n16384_seconds": ech[-1]["seconds"] * (16384.0 / ech[-1]["n"]) ** 3}

    if rank == 0:
        if kara:
            wl = f"{n}x{n} * {n}x{n} Karatsuba two-limb product mod N1*N2, N1 = {N1}, N2 = {N2} (KMatMul!), limbs resident as uint32 residues"
            metric = "Karatsuba mod-(N1*N2) matmul effective GOPS (2n^3/s)"
        else:
            wl = f"{n}x{n} * {n}x{n} matmul mod {N} ({(N - 1).bit_length()}-bit modulus), A,B resident as uint32 residues"
            metric = "mod-p matmul effective GOPS (2n^3/s)"
        if world == 1:
            sharding = "single GPU"
        elif mgpu is not None:
            flow = {"p2p_raw": "the owners' copy engines forward their uint32 column ranges to every rank over NVLink peer memory and every rank splits every range into 8-bit planes itself",
                    "nccl_bcast": "ncclBroadcast of the uint32 column ranges, every rank splits every range into 8-bit planes itself"}.get(
                        transport_used, "each owner splits its column range of B into 8-bit planes and the planes are exchanged over NVLink")
            sharding = (f"{world} row blocks of A and C over {world} GPUs; B lives on rank 0 and is distributed EVERY step by the library's multi-GPU layer "
                        f"(gffm_mg_{'kmat_mul' if kara else 'gemm'}, transport {transport_used}): {flow}, while the GEMMs of the ranges that have arrived run; "
                        f"the distribution of step t+1 overlaps the GEMMs of step t")
        else:
            sharding = f"{world} row blocks of A over {world} GPUs, {transport_used}"
        out = {
            "metric": metric, "value": value, "unit": "GOPS", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "s8 residue limbs, int32 accumulate (exact)",
            "data": "synthetic",
            "config": {"workload": wl, "n": n, "modulus": N if not kara else N1 * N2,
                       "encoding": "RNS int8 tcgen05" if (kara or N > 65536) else "positional int8 limbs tcgen05",
                       "sharding": sharding, "multi_gpu_transport": transport_used, "multi_gpu_info": mgpu.info() if mgpu is not None else None,
                       "shards_match_local_product_on_all_ranks": shard_ok, "multi_gpu_error": mg_error,
                       "warmup_trials_ms": {k_: (round(v, 4) if isinstance(v, float) else v) for k_, v in tune.items()},
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
    if mgpu is not None:
        mgpu.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
