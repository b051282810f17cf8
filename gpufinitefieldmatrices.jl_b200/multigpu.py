"""Multi-GPU layer (new; the reference is single-GPU): one process per GPU, `torch.distributed` for the plumbing.

Matmul and Karatsuba products shard by ROW BLOCKS of A (and C): rank g owns rows [g*ceil(m/G), ...).  B lives on the
source rank and is broadcast in COLUMN PANELS over NVLink.  Two drivers:

* `BroadcastMatmul` (GPUs): the broadcasts run on a dedicated communication stream and mark one CUDA event per panel; ONE
  ABI call (`gffm_gemm_panels`) consumes the panels -- plane split of panel p+1 and CRT of panel p on the library's
  auxiliary stream while the tensor-core GEMM of panel p owns the compute stream.  The library hands back a `consumed`
  event per panel, so the broadcast of the NEXT step's panel p only waits until this step has turned panel p into operand
  planes: in a sequence of products the collective runs entirely under the previous product's GEMMs.
* `pipelined_broadcast_matmul` (backend-agnostic, gloo in the CPU tests): same data flow with the per-panel compute
  injected, one broadcast in flight ahead of the compute.

No reduction is needed in either.
"""
from __future__ import annotations

import ctypes
from typing import Callable, List, Sequence, Tuple

from . import capi

PANEL_ALIGN = 256  # widest GEMM tile (columns): interior panel boundaries are multiples of it (gffm_gemm_panels)


def row_block(m: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [r0, r1) of A/C owned by `rank` (contiguous blocks of ceil(m/world) rows; trailing ranks may be empty)."""
    per = (m + world - 1) // world
    r0 = min(m, rank * per)
    return r0, min(m, r0 + per)


def col_panels(n: int, npanels: int, align: int = 1) -> List[Tuple[int, int]]:
    """Column panels [c0, c1) of B used to pipeline the broadcast against the GEMM.  With `align` > 1 every interior
    boundary is a multiple of it (fewer panels than asked for when n is small)."""
    if n <= 0:
        return []
    npanels = max(1, min(npanels, n))
    per = (n + npanels - 1) // npanels
    per = ((per + align - 1) // align) * align
    return [(c0, min(n, c0 + per)) for c0 in range(0, n, per)]


# ---- optional 2-D process grid ------------------------------------------------------------------------------------------
# Rank r sits at (i, j) = (r // pc, r % pc) of a pr x pc grid: it owns row block i of A and computes the column range j of
# C = A*B, so it needs only B[:, range j].  Row-block sharding with a full broadcast of B is the pc = 1 case.  With pc > 1 every
# rank splits 1/pc of B into operand planes instead of all of it (the part of the work that does not shrink with the GPU
# count) and receives 1/pc of the broadcast bytes; the price is that pc ranks repeat the split of the same A row block.
def process_grid(world: int, pc: int = 1) -> Tuple[int, int]:
    """(pr, pc) with pr*pc == world; raises when pc does not divide the world size."""
    if pc < 1 or world % pc != 0:
        raise ValueError(f"a grid with {pc} column groups does not tile {world} ranks")
    return world // pc, pc


def grid_coords(rank: int, pr: int, pc: int) -> Tuple[int, int]:
    return rank // pc, rank % pc


def col_range(n: int, pc: int, j: int, align: int = 1) -> Tuple[int, int]:
    """Columns [c0, c1) of B / C handled by column group j (equal aligned widths; trailing groups may be narrower or empty)."""
    per = (n + pc - 1) // pc
    per = ((per + align - 1) // align) * align
    c0 = min(n, j * per)
    return c0, min(n, c0 + per)


def make_column_groups(dist, world: int, pc: int, src: int = 0):
    """One communicator per column group j: the ranks of that group plus the source of B.  Collective: every rank must call
    it, and all ranks create the groups in the same order.  Returns the list of (group, ranks)."""
    pr, pc = process_grid(world, pc)
    out = []
    for j in range(pc):
        ranks = sorted(set([src] + [i * pc + j for i in range(pr)]))
        out.append((dist.new_group(ranks=ranks), ranks))
    return out


def grid_deliver(dist, b_colmajor, groups, rank: int, pc: int, n: int, src: int = 0, align: int = 1):
    """deliver(c0, c1) for `BroadcastMatmul` on a pr x pc grid; [c0, c1) is relative to the column range of the calling rank.
    The source takes part in the broadcast of every column group (it holds all of B); every other rank only in its own."""
    my_j = rank % pc

    def deliver(c0, c1):
        for j in (range(pc) if rank == src else [my_j]):
            base, end = col_range(n, pc, j, align)
            lo, hi = min(end, base + c0), min(end, base + c1)
            group, ranks = groups[j]
            if hi > lo and len(ranks) > 1:
                dist.broadcast(b_colmajor[lo:hi], src=src, group=group)

    return deliver


def pipelined_broadcast_matmul(dist, b_colmajor, panels, gemm_panel: Callable[[int, int], None], src: int = 0):
    """One sharded product step, compute injected.

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


def sharded_kmat_mul(dist, b1_colmajor, b2_colmajor, kmat_mul_local: Callable[[], None], src: int = 0):
    """Karatsuba product on row blocks: A1, A2 (and C1, C2) are sharded by rows like the plain product, both limbs of
    B are replicated from `src`, the three sub-products and the recombination run locally (`kmat_mul_local` = gffm_kmat_mul on
    this rank's shard; the `B1 + B2` planes are derived locally inside it)."""
    broadcast_then(dist, [b1_colmajor, b2_colmajor], kmat_mul_local, src=src)


def sharded_gemv(dist, x, gemv_local: Callable[[], None], src: int = 0):
    """z = A*x mod P on row blocks: x is tiny and broadcast whole, every rank computes its rows of z."""
    broadcast_then(dist, [x], gemv_local, src=src)


def broadcast_scatter_allgather(dist, panel, src: int = 0):
    """Broadcast of one column panel as scatter + all-gather: `src` sends a different 1/G slice of the panel to every
    rank, then the ranks all-gather the slices in place.  Every NVLink port carries 1/G of the panel per peer instead of
    the whole panel travelling down one ring, and the all-gather is the collective NCCL runs fastest on NVSwitch (NVLS).
    `panel` is the same region of the replicated buffer on every rank (on `src` it already holds the data); its first
    dimension must be divisible by the world size, otherwise the plain broadcast is used.  Blocking semantics are those of
    the enclosed collectives (stream-ordered on CUDA)."""
    world = dist.get_world_size()
    rank = dist.get_rank()
    rows = panel.shape[0]
    if world == 1:
        return
    if rows % world != 0 or rows == 0:
        dist.broadcast(panel, src=src)
        return
    per = rows // world
    mine = panel[rank * per:(rank + 1) * per]
    if rank == src:
        dist.scatter(mine, [panel[r * per:(r + 1) * per] for r in range(world)], src=src)
    else:
        dist.scatter(mine, None, src=src)
    dist.all_gather_into_tensor(panel, mine)


class BroadcastMatmul:
    """Repeated sharded products C_shard = A_shard * B mod N with B broadcast from `src` every step (GPU only).

    C, A     : CuModMatrix row blocks of this rank (same context).
    B        : CuModMatrix wrapping `b_colmajor` (external memory: never plane-cached).
    b_colmajor: torch tensor (n_cols, ld) int32, row j = column j of B; holds B on `src`, receive buffer elsewhere.
    collective: "broadcast" (ncclBroadcast per panel) or "scatter_allgather" (broadcast_scatter_allgather per panel).
    deliver  : deliver(c0, c1) enqueues the arrival of columns [c0, c1) in `b_colmajor` on the CURRENT torch stream
               (default: dist.broadcast of b_colmajor[c0:c1] from `src`; the 1-GPU tests inject a device copy instead).
    """

    def __init__(self, torch, dist, C, A, B, b_colmajor, panels: Sequence[Tuple[int, int]], src: int = 0, deliver=None,
                 collective: str = "broadcast"):
        self.torch, self.dist = torch, dist
        self.C, self.A, self.B, self.bt = C, A, B, b_colmajor
        self.panels = list(panels)
        self.src = src
        if deliver is None:
            if collective == "scatter_allgather":
                deliver = lambda c0, c1: broadcast_scatter_allgather(dist, b_colmajor[c0:c1], src=src)  # noqa: E731
            elif collective == "broadcast":
                deliver = lambda c0, c1: dist.broadcast(b_colmajor[c0:c1], src=src)  # noqa: E731
            else:
                raise ValueError(f"unknown collective {collective!r}")
        self.deliver = deliver
        dev = b_colmajor.device
        self.comm = torch.cuda.Stream(device=dev)
        np_ = len(self.panels)
        self.ready = [torch.cuda.Event() for _ in range(np_)]
        self.consumed = [torch.cuda.Event() for _ in range(np_)]
        self.off = (ctypes.c_int64 * (np_ + 1))(*([p[0] for p in self.panels] + [self.panels[-1][1]]))
        self._ready_h = (ctypes.c_void_p * np_)()
        self._cons_h = (ctypes.c_void_p * np_)()
        self._first = True

    def step(self):
        torch = self.torch
        cur = torch.cuda.current_stream()
        if self._first:
            # whatever filled / last read the buffers on the caller's stream precedes the first delivery
            self.comm.wait_stream(cur)
            with torch.cuda.stream(cur):
                for e in self.consumed:  # torch creates the CUDA event lazily, at the first record
                    e.record(cur)
        with torch.cuda.stream(self.comm):
            for p, (c0, c1) in enumerate(self.panels):
                self.comm.wait_event(self.consumed[p])  # previous step no longer reads panel p
                self.deliver(c0, c1)
                self.ready[p].record(self.comm)
        for p in range(len(self.panels)):
            self._ready_h[p] = self.ready[p].cuda_event
            self._cons_h[p] = self.consumed[p].cuda_event
        self._first = False
        capi.check(self.C.lib.gffm_gemm_panels(self.C.h, self.A.h, self.B.h, len(self.panels), self.off,
                                               self._ready_h, self._cons_h, 0, 0))

    def finish(self):
        """Make the current stream wait for the communication stream (before the buffers are reused or freed)."""
        self.torch.cuda.current_stream().wait_stream(self.comm)


# ---- the multi-GPU layer of the C ABI (gffm_mg_*, csrc/mg.cu) ---------------------------------------------------------------------
TRANSPORT_NAMES = {capi.MG_AUTO: "auto", capi.MG_NCCL_BCAST: "nccl_bcast", capi.MG_NCCL_PLANES: "nccl_planes", capi.MG_P2P_PLANES: "p2p_planes", capi.MG_P2P_PUSH: "p2p_push", capi.MG_P2P_RAW: "p2p_raw"}


def owner_ranges(n: int, nranks: int) -> List[int]:
    """Column offsets off[0..nranks] of the ranges of B owned by the ranks (gffm_mg_owner_ranges; no GPU needed)."""
    off = (ctypes.c_int64 * (nranks + 1))()
    capi.check(capi.load().gffm_mg_owner_ranges(int(n), int(nranks), off))
    return [int(v) for v in off]


def owner_ranges_root_free(n: int, nranks: int, root: int) -> List[int]:
    """The ranges the peer-memory transports use with many ranks (gffm_mg_owner_ranges_root_free): the root owns no columns."""
    off = (ctypes.c_int64 * (nranks + 1))()
    capi.check(capi.load().gffm_mg_owner_ranges_root_free(int(n), int(nranks), int(root), off))
    return [int(v) for v in off]


class MultiGpu:
    """One rank of the library's own multi-GPU layer: sharded products through `gffm_mg_gemm` / `gffm_mg_kmat_mul` /
    `gffm_mg_gemv`.  The 128-byte id comes from `MultiGpu.unique_id()` on one rank and reaches the others by any channel
    (`from_torch_distributed` uses torch.distributed's object broadcast)."""

    def __init__(self, ctx, rank: int, nranks: int, unique_id: bytes):
        if len(unique_id) != 128:
            raise ValueError("the NCCL unique id has 128 bytes")
        self.ctx, self.lib = ctx, ctx.lib
        self.rank, self.nranks = int(rank), int(nranks)
        h = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(unique_id, 128)
        capi.check(self.lib.gffm_mg_create(ctx.h, buf, self.nranks, self.rank, ctypes.byref(h)))
        self.h = h

    @staticmethod
    def unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        capi.check(capi.load().gffm_mg_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, dist, ctx):
        """Bootstrap over an initialised torch.distributed process group (any backend): rank 0 creates the id."""
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(ctx, rank, world, box[0])

    def info(self):
        r, n, t, p = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        capi.check(self.lib.gffm_mg_info(self.h, ctypes.byref(r), ctypes.byref(n), ctypes.byref(t), ctypes.byref(p)))
        return {"rank": r.value, "nranks": n.value, "transport": TRANSPORT_NAMES.get(t.value, str(t.value)), "peer_memory": bool(p.value & 1),
                "flag_waits": "stream memory ops" if p.value & 2 else "polling kernel", "flag_writes": "stream memory ops" if p.value & 4 else "signal kernel"}

    def set_transport(self, transport: int):
        capi.check(self.lib.gffm_mg_set_transport(self.h, int(transport)))

    def barrier(self):
        capi.check(self.lib.gffm_mg_barrier(self.h))

    def gemm(self, C, A, B, root: int = 0, b_ready=None, R: int = 0, P: int = 0):
        """C_shard = A_shard * B mod P; `b_ready` is a raw cudaEvent_t (int) or None."""
        capi.check(self.lib.gffm_mg_gemm(self.h, C.h, A.h, B.h, int(root), ctypes.c_void_p(b_ready) if b_ready else None, int(R), int(P)))
        return C

    def kmat_mul(self, CK, AK, BK, root: int = 0, b_ready=None):
        """Karatsuba product on row blocks (KMatMul!, KaratsubaMatrix.jl:133-204)."""
        capi.check(self.lib.gffm_mg_kmat_mul(self.h, CK.data1.h, CK.data2.h, AK.data1.h, AK.data2.h, BK.data1.h, BK.data2.h, int(AK.N1), int(AK.N2),
                                             int(root), ctypes.c_void_p(b_ready) if b_ready else None))
        return CK

    def gemv(self, z, A, x, root: int = 0, R: int = 0, P: int = 0):
        capi.check(self.lib.gffm_mg_gemv(self.h, z.h, A.h, x.h, int(root), int(R), int(P)))
        return z

    def close(self):
        if getattr(self, "h", None):
            self.lib.gffm_mg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
