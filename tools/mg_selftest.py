#!/usr/bin/env python
"""Multi-rank self-test of the C-ABI multi-GPU layer (gffm_mg_*, csrc/mg.cu).  Launch under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/mg_selftest.py

Every transport (NCCL broadcast, NCCL planes, peer-memory planes) x several shapes / moduli: each rank's shard of the sharded product
is compared bit for bit with the CPU oracle's rows (small shapes: the whole shard) and with the single-GPU product of the same row
block; sequences of products exercise the double-buffered pipeline with and without the b_ready event, with B changing between
products; Karatsuba and mat-vec forms included.  Prints one JSON line on rank 0 and exits non-zero on any mismatch.
tests/test_multigpu_gpu.py runs this when at least two GPUs are visible."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    import numpy as np
    import torch
    import torch.distributed as dist
    import gffm_b200 as g
    from oracle import oracle as O
    from oracle import oracle_c as OC

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    dist.init_process_group("gloo", rank=rank, world_size=world)  # plumbing only (id exchange, verdict reduction): CPU backend
    g.set_default_device(local)  # the matrices below are created on the default context: make it this rank's device
    ctx = g.default_context()
    mgpu = g.multigpu.MultiGpu.from_torch_distributed(dist, ctx)
    mg = g.multigpu
    results = []
    fails = []

    def shard(m):
        return mg.row_block(m, world, rank)

    def agree(ok, what):
        t = torch.tensor([1 if ok else 0], dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        good = bool(t.item() == 1)
        results.append((what, good))
        if not good:
            fails.append(what)
        if rank == 0:
            print(f"[mg_selftest] {'ok  ' if good else 'FAIL'} {what}", file=sys.stderr, flush=True)
        return good

    transports = [(mg.capi.MG_NCCL_BCAST, "nccl_bcast"), (mg.capi.MG_NCCL_PLANES, "nccl_planes"), (mg.capi.MG_P2P_PLANES, "p2p_planes"),
                  (mg.capi.MG_P2P_PUSH, "p2p_push"), (mg.capi.MG_P2P_RAW, "p2p_raw")]
    if os.environ.get("MG_SELFTEST_TRANSPORTS"):
        want = os.environ["MG_SELFTEST_TRANSPORTS"].split(",")
        transports = [t for t in transports if t[1] in want]
    shapes = [(300, 200, 700, 33554393), (1024, 512, 2048, 33554393), (257, 384, 100, 65521), (513, 640, 1300, 11), (640, 256, 4096, 65521),
              (2048, 1024, 3000, 33554393), (64, 128, 256 * world + 5, 33554393)]
    for tcode, tname in transports:
        try:
            mgpu.set_transport(tcode)
        except g.GffmError as ex:
            agree(False, f"{tname}: set_transport failed: {ex}")
            continue
        for (m, k, n, N) in shapes:
            Ah = O.synth_matrix(1, m, k, N); Bh = O.synth_matrix(2, k, n, N)
            r0, r1 = shard(m)
            A = g.CuModMatrix(Ah[r0:r1], N)
            B = g.CuModMatrix(Bh if rank == 0 else np.zeros((k, n), dtype=np.int64), N)
            C = g.zeros(np.float32, r1 - r0, n, N)
            mgpu.gemm(C, A, B, root=0)
            want = OC.matmul_mod(Ah[r0:r1], Bh, N) if r1 > r0 else np.zeros((0, n), dtype=np.int64)
            agree(np.array_equal(C.to_int(), want), f"{tname}: gemm {m}x{k}x{n} mod {N} vs oracle")
        info = mgpu.info()
        agree(info["transport"] == tname, f"{tname}: resolved transport is {info['transport']}")
        # a sequence of products with B CHANGING in between (b_ready = None: context-stream order), odd and even epochs
        m, k, n, N = 1024, 768, 2048, 33554393
        Ah = O.synth_matrix(3, m, k, N); r0, r1 = shard(m)
        A = g.CuModMatrix(Ah[r0:r1], N); C = g.zeros(np.float32, r1 - r0, n, N)
        B = g.zeros(np.float32, k, n, N)
        ok = True
        for it in range(5):
            Bh = O.synth_matrix(10 + it, k, n, N)
            if rank == 0:
                g.copy_(B, g.CuModMatrix(Bh, N))
            mgpu.gemm(C, A, B, root=0)
            ok = ok and np.array_equal(C.to_int(), OC.matmul_mod(Ah[r0:r1], Bh, N))
        agree(ok, f"{tname}: 5 products with B rewritten on root between them")
        # pipelined sequence: constant B declared ready by an event, A changing; results checked at the end of each step
        Bh = O.synth_matrix(20, k, n, N)
        if rank == 0:
            g.copy_(B, g.CuModMatrix(Bh, N))
        ctx.sync()
        ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream()); torch.cuda.synchronize()
        outs = [g.zeros(np.float32, r1 - r0, n, N) for _ in range(6)]
        As = [g.CuModMatrix(O.synth_matrix(30 + it, m, k, N)[r0:r1], N) for it in range(6)]
        for it in range(6):
            mgpu.gemm(outs[it], As[it], B, root=0, b_ready=ev.cuda_event)
        ok = True
        for it in range(6):
            ok = ok and np.array_equal(outs[it].to_int(), OC.matmul_mod(O.synth_matrix(30 + it, m, k, N)[r0:r1], Bh, N))
        agree(ok, f"{tname}: 6 pipelined products (b_ready event, double-buffered planes)")
        del outs, As
        # pipelined AND changing: B is rewritten on root before every product and declared ready by its own event; any stale operand
        # plane (a race in the double-buffered distribution) shows up as a wrong product
        ok = True
        outs = [g.zeros(np.float32, r1 - r0, n, N) for _ in range(6)]
        evs = [torch.cuda.Event() for _ in range(6)]
        lib_stream = torch.cuda.ExternalStream(ctx.get_stream(), device=local)
        for it in range(6):
            if rank == 0:
                g.copy_(B, g.CuModMatrix(O.synth_matrix(60 + it, k, n, N), N))
            evs[it].record(lib_stream)
            mgpu.gemm(outs[it], A, B, root=0, b_ready=evs[it].cuda_event)
        for it in range(6):
            ok = ok and np.array_equal(outs[it].to_int(), OC.matmul_mod(Ah[r0:r1], O.synth_matrix(60 + it, k, n, N), N))
        agree(ok, f"{tname}: 6 pipelined products, B rewritten before each (per-product ready events)")
        del outs
        # root != 0
        if world > 1:
            Bh = O.synth_matrix(40, k, n, N)
            B2 = g.CuModMatrix(Bh if rank == world - 1 else np.zeros((k, n), dtype=np.int64), N)
            mgpu.gemm(C, A, B2, root=world - 1)
            agree(np.array_equal(C.to_int(), OC.matmul_mod(Ah[r0:r1], Bh, N)), f"{tname}: root = last rank")
        # B already distributed: every rank passes its own column range (root = GFFM_MG_DISTRIBUTED)
        for nn in (2048, 256 * world + 5):
            Bh = O.synth_matrix(41, k, nn, N)
            off = mg.owner_ranges(nn, world)
            Bq = g.CuModMatrix(Bh[:, off[rank]:off[rank + 1]], N)
            Cd = g.zeros(np.float32, r1 - r0, nn, N)
            mgpu.gemm(Cd, A, Bq, root=mg.capi.MG_DISTRIBUTED)
            agree(np.array_equal(Cd.to_int(), OC.matmul_mod(Ah[r0:r1], Bh, N)), f"{tname}: distributed B, n = {nn}")
        # Karatsuba product on row blocks
        for (mm, kk, nn, N1, N2) in [(600, 512, 800, 8191, 8191), (300, 256, 520, 13 ** 4, 13 ** 3)]:
            M = N1 * N2
            rng = np.random.default_rng(mm)
            Ak = rng.integers(0, M, size=(mm, kk), dtype=np.int64); Bk = rng.integers(0, M, size=(kk, nn), dtype=np.int64)
            r0, r1 = shard(mm)
            AK = g.KaratsubaMatrix.from_array(Ak[r0:r1], N1, N2, M)
            BK = g.KaratsubaMatrix.from_array(Bk if rank == 0 else np.zeros((kk, nn), dtype=np.int64), N1, N2, M)
            CK = g.KaratsubaZeros(np.float64, r1 - r0, nn, N1, N2)
            mgpu.kmat_mul(CK, AK, BK, root=0)
            want = OC.matmul_mod(Ak[r0:r1], Bk, M, in_bound=M) if r1 > r0 else np.zeros((0, nn), dtype=np.int64)
            agree(np.array_equal(np.asarray(CK.Array()).astype(np.int64).reshape(r1 - r0, nn), want), f"{tname}: Karatsuba {mm}x{kk}x{nn} N1={N1} N2={N2}")
        # mat-vec on row blocks
        m, k, N = 1000, 900, 33554393
        Ah = O.synth_matrix(50, m, k, N); xh = O.synth_matrix(51, k, 1, N)
        r0, r1 = shard(m)
        z = g.zeros(np.float32, r1 - r0, 1, N)
        x = g.CuModMatrix(xh if rank == 0 else np.zeros((k, 1), dtype=np.int64), N)
        mgpu.gemv(z, g.CuModMatrix(Ah[r0:r1], N), x, root=0)
        agree(np.array_equal(z.to_int().reshape(-1), OC.matmul_mod(Ah[r0:r1], xh, N).reshape(-1)), f"{tname}: gemv")
        try:
            mgpu.barrier()
            agree(True, f"{tname}: barrier / no peer-flag time-out")
        except g.GffmError as ex:
            agree(False, f"{tname}: barrier reported {ex}")

    # timing sketch (informational): n = 8192 per transport, 8 pipelined steps
    timing = {}
    if os.environ.get("MG_SELFTEST_TIMING", "1") == "1":
        n, N = 8192, 33554393
        r0, r1 = shard(n)
        Afull = g.synth(n, n, N, 5, ctx=ctx)
        A = g.zeros(np.float32, r1 - r0, n, N, ctx=ctx)
        g.capi.check(A.lib.gffm_mat_copy_block(A.h, 0, 0, Afull.h, r0, 0, r1 - r0, n))
        del Afull
        B = g.synth(n, n, N, 6, ctx=ctx) if rank == 0 else g.zeros(np.float32, n, n, N, ctx=ctx)
        C = g.zeros(np.float32, r1 - r0, n, N, ctx=ctx)
        ctx.sync()
        ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream()); torch.cuda.synchronize()
        for tcode, tname in transports:
            try:
                mgpu.set_transport(tcode)
                for _ in range(3):
                    A.touch(); mgpu.gemm(C, A, B, root=0, b_ready=ev.cuda_event)
                mgpu.barrier()
                t0 = time.perf_counter()
                for _ in range(8):
                    A.touch(); mgpu.gemm(C, A, B, root=0, b_ready=ev.cuda_event)
                mgpu.barrier()
                timing[tname] = (time.perf_counter() - t0) / 8 * 1e3
            except g.GffmError as ex:
                timing[tname] = f"error: {ex}"
    mgpu.close()
    if rank == 0:
        print(json.dumps({"world": world, "checks": len(results), "failed": fails, "ms_per_step_n8192": timing}))
    dist.barrier()
    dist.destroy_process_group()
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
