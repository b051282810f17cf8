#!/usr/bin/env python
"""One rank's share of the G-GPU sharded product on ONE GPU: A row block (n/G x n), B (n x n) delivered in column panels
by a device copy on the communication stream (stands in for the NCCL broadcast: same event protocol, same HBM writes, no
NVLink).  Prints the compute-side step time of the panel pipeline (gffm_gemm_panels via multigpu.BroadcastMatmul) next
to the previous scheme (one gffm_gemm_block per panel on a single stream) -- what 8-GPU scaling can reach at best.

    python tools/mg_emulate.py [--n 16384] [--world 8] [--panels 8] [--steps 10]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--modulus", type=int, default=33554393)
    ap.add_argument("--world", type=int, nargs="+", default=[8, 4, 2])
    ap.add_argument("--panels", type=int, nargs="+", default=[8])
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import numpy as np
    import torch
    import gffm_b200 as g

    n, N = args.n, args.modulus
    ctx = g.Context(0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    out = []
    with torch.cuda.stream(stream):
        ld = ((n + 31) // 32) * 32
        Bsrc_t = torch.zeros((n, ld), dtype=torch.int32, device="cuda")
        Bsrc = g.CuModMatrix.wrap_device(Bsrc_t.data_ptr(), n, n, ld, N, ctx=ctx)
        g.copy_(Bsrc, g.synth(n, n, N, 6, ctx=ctx))
        Bt = torch.zeros((n, ld), dtype=torch.int32, device="cuda")
        B = g.CuModMatrix.wrap_device(Bt.data_ptr(), n, n, ld, N, ctx=ctx)
        for world in args.world:
            mloc = (n + world - 1) // world
            A = g.synth(mloc, n, N, 5, ctx=ctx)
            C = g.zeros(np.float32, mloc, n, N, ctx=ctx)
            ref = g.zeros(np.float32, mloc, n, N, ctx=ctx)
            g.mul_(ref, A, Bsrc)
            for npan in args.panels:
                panels = g.multigpu.col_panels(n, npan, align=g.multigpu.PANEL_ALIGN)

                def deliver(c0, c1):
                    Bt[c0:c1].copy_(Bsrc_t[c0:c1], non_blocking=True)

                bm = g.multigpu.BroadcastMatmul(torch, None, C, A, B, Bt, panels, deliver=deliver)

                def step_new():
                    A.touch()
                    bm.step()

                def step_old():  # previous scheme: deliveries issued up front, one gemm_block per panel on the compute stream
                    A.touch()
                    evs = []
                    with torch.cuda.stream(bm.comm):
                        bm.comm.wait_stream(stream)
                        for (c0, c1) in panels:
                            deliver(c0, c1)
                            e = torch.cuda.Event(); e.record(bm.comm); evs.append(e)
                    for e, (c0, c1) in zip(evs, panels):
                        stream.wait_event(e)
                        g.capi.check(C.lib.gffm_gemm_block(C.h, 0, c0, A.h, 0, 0, B.h, 0, c0, mloc, c1 - c0, n, 0, 0, g.capi.GEMM_STORE, g.capi.ALGO_AUTO))

                res = {"world": world, "mloc": mloc, "panels": len(panels)}
                for name, fn in (("old_ms", step_old), ("new_ms", step_new)):
                    g.fill_(C, 0)
                    for _ in range(3):
                        fn()
                    bm.finish()
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record(stream)
                    for _ in range(args.steps):
                        fn()
                    bm.finish()
                    e1.record(stream)
                    torch.cuda.synchronize()
                    res[name] = e0.elapsed_time(e1) / args.steps
                    res[name.replace("_ms", "_ok")] = bool(C.equals(ref))
                out.append(res)
                print(json.dumps(res), flush=True)
            del A, C, ref
    return 0


if __name__ == "__main__":
    sys.exit(main())
