"""Multi-GPU layer (new; the reference is single-GPU): one process per GPU, `torch.distributed` for the plumbing.

Matmul and Karatsuba products shard by ROW BLOCKS of A (and C): rank g owns rows [g*ceil(m/G), ...).  B lives on the
source rank and is broadcast in COLUMN PANELS; the modular GEMM of panel p (through the C ABI, `gffm_gemm_block`) is
enqueued as soon as panel p has arrived, so it overlaps the NCCL broadcast of panel p+1 over NVLink.  No reduction is
needed.  The pipeline below is backend-agnostic (NCCL on GPUs, gloo in the CPU tests) -- the compute is injected.
"""
from __future__ import annotations

from typing import Callable, List, Tuple


def row_block(m: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [r0, r1) of A/C owned by `rank` (contiguous blocks of ceil(m/world) rows; trailing ranks may be empty)."""
    per = (m + world - 1) // world
    r0 = min(m, rank * per)
    return r0, min(m, r0 + per)


def col_panels(n: int, npanels: int) -> List[Tuple[int, int]]:
    """Column panels [c0, c1) of B used to pipeline the broadcast against the GEMM."""
    npanels = max(1, min(npanels, n if n > 0 else 1))
    per = (n + npanels - 1) // npanels
    return [(c0, min(n, c0 + per)) for c0 in range(0, n, per)] if n > 0 else []


def pipelined_broadcast_matmul(dist, b_colmajor, panels, gemm_panel: Callable[[int, int], None], src: int = 0):
    """One sharded product step.

    b_colmajor : tensor of shape (n_cols, ld) whose row j is column j of B (column-major storage); on `src` it holds
                 B, on the other ranks it is the receive buffer.
    gemm_panel : gemm_panel(c0, c1) enqueues C_shard[:, c0:c1] = A_shard * B[:, c0:c1] mod N on the current stream.
    All broadcasts are issued first (async); waiting on work p only makes the compute stream depend on panel p, so the
    collective of panel p+1 runs concurrently with the GEMM of panel p.
    """
    works = [dist.broadcast(b_colmajor[c0:c1], src=src, async_op=True) for (c0, c1) in panels]
    for w, (c0, c1) in zip(works, panels):
        w.wait()
        gemm_panel(c0, c1)


def broadcast_then(dist, tensors, fn: Callable[[], None], src: int = 0):
    """Broadcast whole operands (e.g. both limbs of a Karatsuba B) and run `fn` once they have arrived.  Used where the
    product has no per-panel entry point (gffm_kmat_mul): the shard of A stays local, B1/B2 are replicated."""
    works = [dist.broadcast(t, src=src, async_op=True) for t in tensors]
    for w in works:
        w.wait()
    fn()
