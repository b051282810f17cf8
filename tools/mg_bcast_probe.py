#!/usr/bin/env python
"""Stand-alone bandwidth of the two ways the multi-GPU layer can replicate B (1 GiB of uint32 residues at n = 16384):
ncclBroadcast per column panel vs scatter + in-place all-gather per panel (multigpu.broadcast_scatter_allgather).
Run under torchrun:  python -m torch.distributed.run --nproc-per-node G tools/mg_bcast_probe.py [--n 16384] [--panels 8]
Rank 0 prints one JSON line; device-timed (CUDA events), max over ranks, nothing else running on the GPUs."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--panels", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import gffm_b200 as g
    os.environ["NCCL_DEBUG"] = os.environ.get("GFFM_NCCL_DEBUG", "WARN")
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n
    buf = torch.zeros((n, n), dtype=torch.int32, device="cuda")
    if rank == 0:
        buf.copy_(torch.randint(0, 2 ** 25, (n, n), dtype=torch.int32, device="cuda"))
    panels = g.multigpu.col_panels(n, args.panels, align=g.multigpu.PANEL_ALIGN)
    res = {"world": world, "n": n, "panels": len(panels), "bytes": 4 * n * n}

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.reps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def bcast_panels():
        for (c0, c1) in panels:
            dist.broadcast(buf[c0:c1], src=0)

    def sag_panels():
        for (c0, c1) in panels:
            g.multigpu.broadcast_scatter_allgather(dist, buf[c0:c1], src=0)

    for name, fn in (("broadcast_whole", lambda: dist.broadcast(buf, src=0)), ("broadcast_panels", bcast_panels), ("scatter_allgather_panels", sag_panels)):
        if rank != 0:
            buf.zero_()
        ms = timed(fn)
        chk = torch.tensor([int(buf.view(-1)[::4099].to(torch.int64).sum().item())], device="cuda", dtype=torch.int64)
        lo = chk.clone(); hi = chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        res[name] = {"ms": ms, "GBps": 4 * n * n / ms / 1e6, "replicas_identical": bool(lo.item() == hi.item())}
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
