"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row-block sharding of A/C, pipelined column-panel
broadcast of B.  The per-panel compute is injected (here: the CPU oracle, test-only) -- on GPUs it is gffm_gemm_block."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, m, k, n, N, npan, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import gffm_b200 as g
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mg = g.multigpu
    r0, r1 = mg.row_block(m, world, rank)
    A = O.synth_matrix(11, m, k, N)
    A_shard = A[r0:r1]
    ld = ((k + 31) // 32) * 32
    Bt = torch.zeros((n, ld), dtype=torch.int32)          # column-major k x n, row j = column j
    if rank == 0:
        Bt[:, :k] = torch.from_numpy(O.synth_matrix(12, k, n, N).T.astype(np.int32).copy())
    C_shard = np.zeros((r1 - r0, n), dtype=np.int64)

    def gemm_panel(c0, c1):
        Bp = Bt[c0:c1, :k].numpy().T.astype(np.int64)
        C_shard[:, c0:c1] = O.matmul_mod(A_shard, Bp, N)

    mg.pipelined_broadcast_matmul(dist, Bt, mg.col_panels(n, npan), gemm_panel, src=0)
    np.save(os.path.join(out_dir, f"c_{rank}.npy"), C_shard)
    dist.barrier()
    dist.destroy_process_group()


def _worker_kmat_gemv(rank, world, port, m, k, n, N1, N2, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import gffm_b200 as g
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mg = g.multigpu
    r0, r1 = mg.row_block(m, world, rank)
    A1 = O.synth_matrix(21, m, k, N1)[r0:r1]; A2 = O.synth_matrix(22, m, k, N2)[r0:r1]
    B1t = torch.zeros((n, k), dtype=torch.int64); B2t = torch.zeros((n, k), dtype=torch.int64)
    x = torch.zeros(k, dtype=torch.int64)
    if rank == 0:
        B1t.copy_(torch.from_numpy(O.synth_matrix(23, k, n, N1).T.copy())); B2t.copy_(torch.from_numpy(O.synth_matrix(24, k, n, N2).T.copy()))
        x.copy_(torch.from_numpy(O.synth_matrix(25, k, 1, N1)[:, 0].copy()))
    res = {}

    def kmul():
        res["C1"], res["C2"] = O.karatsuba_matmul(A1, A2, B1t.numpy().T, B2t.numpy().T, N1, N2)

    def gemv():
        res["z"] = O.matvec_mod(A1, x.numpy(), N1)

    mg.sharded_kmat_mul(dist, B1t, B2t, kmul, src=0)
    mg.sharded_gemv(dist, x, gemv, src=0)
    np.savez(os.path.join(out_dir, f"k_{rank}.npz"), C1=res["C1"], C2=res["C2"], z=np.asarray(res["z"]).reshape(-1))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_karatsuba_and_gemv_world2(tmp_path):
    """Row-block sharding of the Karatsuba product (both limbs of B replicated) and of the GEMV (x replicated)."""
    world, m, k, n, N1, N2 = 2, 21, 16, 10, 13 ** 4, 13 ** 3
    port = _free_port()
    mp.spawn(_worker_kmat_gemv, args=(world, port, m, k, n, N1, N2, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"k_{r}.npz") for r in range(world)]
    A1 = O.synth_matrix(21, m, k, N1); A2 = O.synth_matrix(22, m, k, N2)
    B1 = O.synth_matrix(23, k, n, N1); B2 = O.synth_matrix(24, k, n, N2); x = O.synth_matrix(25, k, 1, N1)[:, 0]
    C1, C2 = O.karatsuba_matmul(A1, A2, B1, B2, N1, N2)
    assert np.array_equal(np.concatenate([p["C1"] for p in parts]), C1) and np.array_equal(np.concatenate([p["C2"] for p in parts]), C2)
    full = (A1.astype(object) + N1 * A2.astype(object)).dot(B1.astype(object) + N1 * B2.astype(object)) % (N1 * N2)
    assert np.array_equal(C1.astype(object) + N1 * C2.astype(object), full)
    assert np.array_equal(np.concatenate([p["z"] for p in parts]), np.asarray(O.matvec_mod(A1, x, N1)).reshape(-1))


def _worker_sag(rank, world, port, rows, cols, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import gffm_b200 as g
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    buf = torch.zeros((rows, cols), dtype=torch.int32)
    if rank == 0:
        buf.copy_(torch.arange(rows * cols, dtype=torch.int32).reshape(rows, cols) * 7 + 3)
    for (c0, c1) in g.multigpu.col_panels(rows, 3):
        g.multigpu.broadcast_scatter_allgather(dist, buf[c0:c1], src=0)
    np.save(os.path.join(out_dir, f"b_{rank}.npy"), buf.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,rows", [(2, 12), (3, 18), (2, 7)])
def test_broadcast_scatter_allgather(tmp_path, world, rows):
    """scatter + in-place all-gather == broadcast on every rank (divisible panels; indivisible ones fall back to broadcast)."""
    port = _free_port()
    mp.spawn(_worker_sag, args=(world, port, rows, 5, str(tmp_path)), nprocs=world, join=True)
    want = (np.arange(rows * 5, dtype=np.int32).reshape(rows, 5) * 7 + 3)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"b_{r}.npy"), want), f"rank {r}"


def _worker_grid(rank, world, port, pc, m, k, n, N, npan, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import gffm_b200 as g
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mg = g.multigpu
    pr, pc = mg.process_grid(world, pc)
    i, j = mg.grid_coords(rank, pr, pc)
    r0, r1 = mg.row_block(m, pr, i)
    c0, c1 = mg.col_range(n, pc, j)
    A_shard = O.synth_matrix(11, m, k, N)[r0:r1]
    ld = ((k + 31) // 32) * 32
    Bt = torch.zeros((n, ld), dtype=torch.int32)
    if rank == 0:
        Bt[:, :k] = torch.from_numpy(O.synth_matrix(12, k, n, N).T.astype(np.int32).copy())
    groups = mg.make_column_groups(dist, world, pc, src=0)
    deliver = mg.grid_deliver(dist, Bt, groups, rank, pc, n, src=0)
    C_shard = np.zeros((r1 - r0, c1 - c0), dtype=np.int64)
    width = max(c1 - c0 for (c0, c1) in [mg.col_range(n, pc, jj) for jj in range(pc)])
    for (p0, p1) in mg.col_panels(width, npan):          # panel offsets are relative to the column range
        deliver(p0, p1)
        lo, hi = min(c1, c0 + p0), min(c1, c0 + p1)
        if hi > lo:
            Bp = Bt[lo:hi, :k].numpy().T.astype(np.int64)
            C_shard[:, lo - c0:hi - c0] = O.matmul_mod(A_shard, Bp, N)
    np.save(os.path.join(out_dir, f"c_{rank}.npy"), C_shard)
    if rank != 0:  # ranks never receive columns outside their range
        other = torch.cat([Bt[:c0], Bt[c1:]])
        assert not other.any()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,pc,m,k,n,N,npan", [(4, 2, 70, 50, 90, 33554393, 3), (4, 4, 9, 20, 30, 11, 2), (2, 2, 33, 17, 5, 65521, 2), (3, 1, 20, 10, 12, 7, 2)])
def test_grid_sharded_matmul(tmp_path, world, pc, m, k, n, N, npan):
    """pr x pc process grid: row block i of A times column range j of B on rank (i, j); B travels only to the ranks that need it."""
    import gffm_b200 as g
    mg = g.multigpu
    port = _free_port()
    mp.spawn(_worker_grid, args=(world, port, pc, m, k, n, N, npan, str(tmp_path)), nprocs=world, join=True)
    pr, pc = mg.process_grid(world, pc)
    A = O.synth_matrix(11, m, k, N); B = O.synth_matrix(12, k, n, N)
    want = O.matmul_mod(A, B, N)
    for r in range(world):
        i, j = mg.grid_coords(r, pr, pc)
        r0, r1 = mg.row_block(m, pr, i); c0, c1 = mg.col_range(n, pc, j)
        assert np.array_equal(np.load(tmp_path / f"c_{r}.npy"), want[r0:r1, c0:c1]), f"rank {r}"
    with pytest.raises(ValueError):
        mg.process_grid(8, 3)


@pytest.mark.parametrize("m,k,n,N,npan", [(70, 50, 90, 33554393, 4), (5, 9, 3, 11, 8)])
def test_sharded_matmul_world2(tmp_path, m, k, n, N, npan):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, m, k, n, N, npan, str(tmp_path)), nprocs=world, join=True)
    C = np.concatenate([np.load(tmp_path / f"c_{r}.npy") for r in range(world)], axis=0)
    A = O.synth_matrix(11, m, k, N); B = O.synth_matrix(12, k, n, N)
    assert np.array_equal(C, O.matmul_mod(A, B, N))


def test_partition_helpers():
    import gffm_b200 as g
    mg = g.multigpu
    for m in (0, 1, 7, 16384, 16385):
        for world in (1, 2, 3, 8):
            blocks = [mg.row_block(m, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == m
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    for n in (1, 5, 16384):
        for p in (1, 3, 8, 100):
            pans = mg.col_panels(n, p)
            assert pans[0][0] == 0 and pans[-1][1] == n and all(a[1] == b[0] for a, b in zip(pans, pans[1:]))
    # aligned panels (gffm_gemm_panels): interior boundaries are multiples of the GEMM tile width, no empty panel
    for n in (1, 255, 256, 1000, 16384, 16385, 32768):
        for p in (1, 3, 8, 16):
            pans = mg.col_panels(n, p, align=mg.PANEL_ALIGN)
            assert pans[0][0] == 0 and pans[-1][1] == n and all(a[1] == b[0] for a, b in zip(pans, pans[1:]))
            assert all(c0 % mg.PANEL_ALIGN == 0 and c1 > c0 for c0, c1 in pans) and len(pans) <= p
    assert mg.col_panels(16384, 8, align=256) == [(2048 * i, 2048 * (i + 1)) for i in range(8)]
