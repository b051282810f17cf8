"""BASELINE config 5 on N GPUs: 32768 x 32768 plain mod-p matmul and Karatsuba product, A sharded by row blocks, B broadcast
with NCCL.  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/mg_config5.py [n]
Prints one JSON line per workload (rank 0): ms per step (max over ranks, CUDA events) and effective TOP/s."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import gffm_b200 as g

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = g.Context(local)
stream = torch.cuda.Stream(device=local)
ctx.set_stream(stream.cuda_stream)
mg = g.multigpu
r0, r1 = mg.row_block(n, world, rank)
mloc = r1 - r0
ld = ((n + 31) // 32) * 32


def shard(seed, N):
    full = g.synth(n, n, N, seed, ctx=ctx)
    s = g.zeros(np.float32, mloc, n, N, ctx=ctx)
    g.capi.check(s.lib.gffm_mat_copy_block(s.h, 0, 0, full.h, r0, 0, mloc, n))
    ctx.sync()
    return s


def bcast_buffer(seed, N):
    t = torch.zeros((n, ld), dtype=torch.int32, device=f"cuda:{local}")
    m = g.CuModMatrix.wrap_device(t.data_ptr(), n, n, ld, N, ctx=ctx)
    if rank == 0:
        src = g.synth(n, n, N, seed, ctx=ctx)
        g.copy_(m, src); ctx.sync()
    return t, m


def timed(step, steps=5, warm=2):
    for _ in range(warm):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


with torch.cuda.stream(stream):
    # ---- plain product mod a 25-bit prime
    N = 33554393
    A = shard(11, N); Bt, B = bcast_buffer(12, N); Cm = g.zeros(np.float32, mloc, n, N, ctx=ctx)
    panels = mg.col_panels(n, 8 if world > 1 else 1)

    def gemm_panel(c0, c1):
        g.capi.check(Cm.lib.gffm_gemm_block(Cm.h, 0, c0, A.h, 0, 0, B.h, 0, c0, mloc, c1 - c0, n, 0, 0, g.capi.GEMM_STORE, g.capi.ALGO_AUTO))

    def step_plain():
        A.touch()
        if world == 1:
            g.mul_(Cm, A, B)
        else:
            mg.pipelined_broadcast_matmul(dist, Bt, panels, gemm_panel, src=0)

    ms = timed(step_plain)
    # Freivalds on the shard with the independent GEMV kernel
    x = g.synth(n, 1, N, 77, ctx=ctx)
    Bx = g.zeros(np.float64, n, 1, N, ctx=ctx); g.gemv_(Bx, B, x)
    ABx = g.zeros(np.float64, mloc, 1, N, ctx=ctx); g.gemv_(ABx, A, Bx)
    Cx = g.zeros(np.float64, mloc, 1, N, ctx=ctx); g.gemv_(Cx, Cm, x)
    ok = ABx.equals(Cx)
    if rank == 0:
        print(json.dumps({"workload": f"plain {n}^3 mod {N}", "n_gpus": world, "ms_per_step": ms, "effective_TOPS": 2.0 * n ** 3 / ms / 1e9, "freivalds_rank0": ok}), flush=True)
    del A, B, Bt, Cm, x, Bx, ABx, Cx

    # ---- Karatsuba product, N1 = N2 = 8191
    N1 = N2 = 8191
    A1 = shard(13, N1); A2 = shard(14, N1)
    B1t, B1 = bcast_buffer(15, N1); B2t, B2 = bcast_buffer(16, N1)
    C1 = g.zeros(np.float64, mloc, n, N1, ctx=ctx); C2 = g.zeros(np.float64, mloc, n, N1, ctx=ctx)
    AK = g.KaratsubaMatrix(A1, A2, N1, N2); BK = g.KaratsubaMatrix(B1, B2, N1, N2); CK = g.KaratsubaMatrix(C1, C2, N1, N2)

    def step_kara():
        A1.touch(); A2.touch()
        if world == 1:
            g.KMatMul_(CK, AK, BK)
        else:
            mg.sharded_kmat_mul(dist, B1t, B2t, lambda: g.KMatMul_(CK, AK, BK), src=0)

    ms = timed(step_kara, steps=3, warm=1)
    x = g.synth(n, 1, N1, 78, ctx=ctx)
    t = g.zeros(np.float64, n, 1, N1, ctx=ctx); g.gemv_(t, B1, x, P=N1)
    l = g.zeros(np.float64, mloc, 1, N1, ctx=ctx); g.gemv_(l, A1, t, P=N1)
    r = g.zeros(np.float64, mloc, 1, N1, ctx=ctx); g.gemv_(r, C1, x, P=N1)
    ok = l.equals(r)
    if rank == 0:
        print(json.dumps({"workload": f"karatsuba {n}^3 mod {N1}*{N2}", "n_gpus": world, "ms_per_step": ms, "effective_TOPS": 2.0 * n ** 3 / ms / 1e9, "low_limb_freivalds_rank0": ok}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
