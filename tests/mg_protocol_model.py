"""Executable model of the peer-memory protocols of the library's multi-GPU layer (csrc/mg.cu) -- TEST INFRASTRUCTURE ONLY.

The multi-GPU layer cannot run without GPUs, and its correctness is a matter of ORDER: which stream of which rank may touch which
buffer when.  This module restates the enqueue logic of `mg_distribute` / `gffm_bplan_gemm` / `mg_round_done` for the three peer-memory
transports (GFFM_MG_P2P_RAW, GFFM_MG_P2P_PUSH, GFFM_MG_P2P_PLANES) as per-stream operation lists and executes them under an adversarial
scheduler with CUDA's semantics:

  * a stream is a FIFO; an operation starts when everything before it in its stream has ended;
  * `cudaEventRecord` / `cudaStreamWaitEvent`: a wait refers to the most recent record of that event AT ENQUEUE TIME (none: no-op);
  * stream memory operations: `wait_flag` blocks until (flag - target) >= 0, `write_flag` stores into a (peer) rank's control words;
  * copies, splits and GEMMs take time: they are a start (locks taken: sources shared, destinations exclusive) and an end (locks released,
    destination content := source content), and any other stream may run in between.

What is checked, over random schedules: no deadlock; no two operations that are not ordered by the protocol touch the same buffer
region in conflicting ways (a started write meets a reader or writer, a started read meets a writer); every split reads the residues of
ITS product and every GEMM the planes of ITS product (content tags), also when the caller rewrites B before every product.

Mirrors (keep in sync): mg.cu `mg_distribute` RAW block, PLANES/PUSH block, `mg_round_done`, `gffm_mg_gemm`; gemm_tc.cu `gffm_bplan_gemm`
(per-range wait + launch).  NCCL transports are collectives ordered by NCCL itself and are not modelled."""
from __future__ import annotations

import random
from collections import defaultdict, deque

NBUF = 3   # MG_NBUF
NCOPY = 4  # gffm_mg::NCOPY (copy streams per direction)


class ProtocolError(AssertionError):
    pass


class World:
    def __init__(self, nr: int, seed: int = 0):
        self.nr = nr
        self.rng = random.Random(seed)
        self.streams = {}                      # (rank, name) -> deque of ops
        self.flags = defaultdict(int)          # (rank, word, idx) -> value
        self.latest = {}                       # (rank, event) -> token of the most recent record at enqueue time
        self.done = set()                      # completed record tokens
        self.ntok = 0
        self.readers = defaultdict(int)        # region -> number of started, unfinished readers
        self.writers = defaultdict(int)
        self.content = {}                      # region -> tag
        self.log = []

    # ---- enqueue side (what the host does) ----------------------------------------------------------------------------------
    def q(self, rank, stream):
        return self.streams.setdefault((rank, stream), deque())

    def record(self, rank, stream, event):
        self.ntok += 1
        self.latest[(rank, event)] = self.ntok
        self.q(rank, stream).append(("record", self.ntok))

    def wait_event(self, rank, stream, event):
        tok = self.latest.get((rank, event))
        if tok is not None:
            self.q(rank, stream).append(("wait_event", tok))

    def wait_flag(self, rank, stream, word, idx, target):
        self.q(rank, stream).append(("wait_flag", (rank, word, idx), target))

    def write_flag(self, rank, stream, dst_rank, word, idx, value):
        self.q(rank, stream).append(("write_flag", (dst_rank, word, idx), value))

    def work(self, rank, stream, what, srcs, dsts, expect=None, tag=None):
        """a copy / split / GEMM / caller write: reads `srcs`, writes `dsts`; `expect`: tag every source must carry; `tag`: content
        written (default: the sources' common tag)"""
        op = {"what": what, "srcs": list(srcs), "dsts": list(dsts), "expect": expect, "tag": tag, "rank": rank}
        self.q(rank, stream).append(("start", op))
        self.q(rank, stream).append(("end", op))

    # ---- execution side ---------------------------------------------------------------------------------------------------------
    def _ready(self, op):
        kind = op[0]
        if kind == "wait_event":
            return op[1] in self.done
        if kind == "wait_flag":
            return self.flags[op[1]] - op[2] >= 0
        return True

    def _exec(self, key, op):
        kind = op[0]
        if kind == "record":
            self.done.add(op[1])
        elif kind == "write_flag":
            # flags only ever grow in the protocol (epochs); a smaller value would be a protocol bug of its own
            if op[2] < self.flags[op[1]]:
                raise ProtocolError(f"flag {op[1]} written backwards: {self.flags[op[1]]} -> {op[2]} by {key}")
            self.flags[op[1]] = op[2]
        elif kind == "start":
            o = op[1]
            for r in o["srcs"]:
                if self.writers[r]:
                    raise ProtocolError(f"RACE: {o['what']} on {key} starts reading {r} while it is being written")
                self.readers[r] += 1
            for r in o["dsts"]:
                if self.writers[r] or self.readers[r]:
                    raise ProtocolError(f"RACE: {o['what']} on {key} starts writing {r} while it is in use "
                                        f"(readers {self.readers[r]}, writers {self.writers[r]})")
                self.writers[r] += 1
            tags = {self.content.get(r) for r in o["srcs"]}
            if o["expect"] is not None and tags != {o["expect"]}:
                raise ProtocolError(f"STALE: {o['what']} on {key} expected content {o['expect']} in {o['srcs']}, found {tags}")
            o["_tag"] = o["tag"] if o["tag"] is not None else (next(iter(tags)) if len(tags) == 1 else ("mixed", tuple(sorted(map(str, tags)))))
        elif kind == "end":
            o = op[1]
            for r in o["srcs"]:
                self.readers[r] -= 1
            for r in o["dsts"]:
                self.writers[r] -= 1
                self.content[r] = o["_tag"]

    def run(self, max_steps=10_000_000):
        """Adversarial schedule: every rank and every stream draws a speed (most are fast, some crawl), so that single streams or whole
        ranks fall several products behind -- the situations the buffer-reuse rules exist for.  A ready stream is chosen with
        probability proportional to its speed."""
        keys = list(self.streams.keys())
        rank_speed = {r: self.rng.choice([1.0, 1.0, 1.0, 0.05, 0.01]) for r in range(self.nr)}
        speed = {k: rank_speed[k[0]] * self.rng.choice([1.0, 1.0, 0.2, 0.02]) for k in keys}
        steps = 0
        while True:
            live = [k for k in keys if self.streams[k]]
            if not live:
                return steps
            ready = [k for k in live if self._ready(self.streams[k][0])]
            if not ready:
                heads = {k: self.streams[k][0][:3] for k in live}
                raise ProtocolError(f"DEADLOCK: no stream can make progress; heads: {heads}")
            k = self.rng.choices(ready, weights=[speed[x] for x in ready])[0]
            self._exec(k, self.streams[k].popleft())
            steps += 1
            if steps > max_steps:
                raise ProtocolError("model did not terminate")


# ---- ranges (mg.cu: uniform ranges / mg_root_free_ranges; here in units of 256-column blocks) -------------------------------------
def ranges(nr, root, root_free, blocks):
    if not root_free:
        per = -(-blocks // nr)
        off = [min(blocks, q * per) for q in range(nr + 1)]
    else:
        base, extra = divmod(blocks, nr - 1)
        off, o = [0], 0
        for q in range(nr):
            nb = 0
            if q != root:
                nb = base + (1 if o < extra else 0)
                o += 1
            off.append(min(blocks, off[-1] + nb))
    return [off[q + 1] - off[q] for q in range(nr)]


# region names: ("B", rank, q)  residues of range q in rank's own matrix;  ("S", rank, b, q, part)  staging;  ("P", rank, b, q, part)  planes
def S(rank, b, q):
    return [("S", rank, b, q, 0), ("S", rank, b, q, 1)]


def P(rank, b, q):
    return [("P", rank, b, q, 0), ("P", rank, b, q, 1)]


class Rank:
    """the enqueue logic of one rank's library calls; `broken` switches off one ordering rule at a time (the model must then fail)"""

    def __init__(self, w: World, r: int, transport: str, root_free_min: int = 6, broken: str = ""):
        self.w, self.r, self.nr, self.transport, self.root_free_min, self.broken = w, r, w.nr, transport, root_free_min, broken
        self.epoch = 0

    def product(self, root: int, bready, blocks: int, version):
        """gffm_mg_gemm (one K chunk): distribute, per-range GEMMs, round_done.  bready: event name or None (context-stream order);
        version: the content tag B carries for this product"""
        w, r, nr, tr = self.w, self.r, self.nr, self.transport
        distributed = root < 0
        root_free = (not distributed) and self.root_free_min > 0 and nr >= self.root_free_min and nr >= 2
        cnt = ranges(nr, root, root_free, blocks)
        self.epoch += 1
        e = self.epoch
        b = e % NBUF
        if bready is None:
            w.record(r, "ctx", "ev_call")
            bready = "ev_call"
        back = 0 if self.broken == "no_buffer_reuse_wait" else NBUF
        have_b = r == root or distributed

        def scatter(raw):
            for j in range(NCOPY):
                w.wait_event(r, f"push{j}", bready)
            for i in range(1, nr):
                qq = (root + i) % nr
                used = [0, 1] if cnt[qq] > 0 else [0]
                for j in used:
                    if back:
                        w.wait_flag(r, f"push{j}", "SPLIT_DONE", qq, e - back)
                    if cnt[qq] > 0:
                        w.work(r, f"push{j}", f"scatter e{e} ->r{qq}", [("B", r, qq)], [("S", qq, b, qq, j)])
                    if j > 0:
                        w.record(r, f"push{j}", f"copy_ev1_{j}")
                        w.wait_event(r, "push0", f"copy_ev1_{j}")
                w.write_flag(r, "push0", qq, "STAGED", 0, e)
            if not raw:
                w.record(r, "push0", "ev_push")

        if tr == "raw":
            if r == root:
                scatter(True)
            if not have_b:
                w.wait_flag(r, "dist", "STAGED", 0, e)
            else:
                w.wait_event(r, "dist", bready)
            w.wait_event(r, "dist", f"gemm_done{b}")
            used_push = set(range(2)) if r == root else set()
            if cnt[r] > 0:
                for i in range(1, nr):
                    p = (r + i) % nr
                    if p == root:
                        continue
                    cs = f"push{i % NCOPY}"
                    used_push.add(i % NCOPY)
                    if not have_b:
                        w.wait_flag(r, cs, "STAGED", 0, e)
                    else:
                        w.wait_event(r, cs, bready)
                    if back:
                        w.wait_flag(r, cs, "SPLIT_DONE", p, e - back)
                    src = [("B", r, r)] if have_b else S(r, b, r)
                    w.work(r, cs, f"fwd e{e} r{r}->r{p}", src, S(p, b, r))
                    w.write_flag(r, cs, p, "READY", r, e)
            for j in range(1, NCOPY):
                if j in used_push:
                    w.record(r, f"push{j}", f"copy_ev1_{j}")
                    w.wait_event(r, "push0", f"copy_ev1_{j}")
            w.record(r, "push0", "ev_push")
            for i in range(nr):
                qq = (r - i) % nr  # arrival order of the forwards (owner q sends to q+1, q+2, ...)
                local = have_b if qq == r else r == root
                if not local and qq != r and cnt[qq] > 0:
                    w.wait_flag(r, "dist", "READY", qq, e)
                if cnt[qq] > 0:
                    w.work(r, "dist", f"split e{e} range {qq}", [("B", r, qq)] if local else S(r, b, qq), P(r, b, qq), expect=version)
                w.record(r, "dist", f"ready{b}_{qq}")
            if self.broken != "no_forward_join":
                w.wait_event(r, "dist", "ev_push")
            for qq in range(nr):
                if qq != r:
                    w.write_flag(r, "dist", qq, "SPLIT_DONE", r, e)
            w.record(r, "dist", f"split_ev{b}")
        else:
            push = tr == "push"
            if r == root:
                scatter(False)
            if r != root and not distributed:
                w.wait_flag(r, "dist", "STAGED", 0, e)
            else:
                w.wait_event(r, "dist", bready)
            w.wait_event(r, "dist", f"gemm_done{b}")
            if back:
                for qq in range(nr):
                    if qq != r:
                        w.wait_flag(r, "dist", "FREE" if push else "PULLED", qq, e - back)
            from_stage = r != root and not distributed
            src = S(r, b, r) if from_stage else [("B", r, r)]
            if cnt[r] > 0:
                dsts = [reg for qq in range(nr) for reg in P(qq, b, r)] if push else P(r, b, r)
                w.work(r, "dist", f"split{'+push' if push else ''} e{e} range {r}", src, dsts, expect=version)
            w.record(r, "dist", f"ready{b}_{r}")
            for i in range(1, nr):
                w.write_flag(r, "dist", (r + i) % nr, "READY", r, e)
            for qq in range(nr):
                if qq != r:
                    w.write_flag(r, "dist", qq, "SPLIT_DONE", r, e)
            if push:
                for i in range(1, nr):
                    qq = (r + i) % nr
                    w.wait_flag(r, "pull0", "READY", qq, e)
                    w.record(r, "pull0", f"ready{b}_{qq}")
            else:
                for j in range(NCOPY):
                    w.wait_event(r, f"pull{j}", f"gemm_done{b}")
                for i in range(1, nr):
                    qq = (r + i) % nr
                    used = [0, 1] if cnt[qq] > 0 else [0]
                    for j in used:
                        w.wait_flag(r, f"pull{j}", "READY", qq, e)
                        if cnt[qq] > 0:
                            w.work(r, f"pull{j}", f"pull e{e} range {qq}", [("P", qq, b, qq, j)], [("P", r, b, qq, j)])
                        if j > 0:
                            w.record(r, f"pull{j}", f"copy_ev0_{j}")
                            w.wait_event(r, "pull0", f"copy_ev0_{j}")
                    w.record(r, "pull0", f"ready{b}_{qq}")
                    w.write_flag(r, "pull0", qq, "PULLED", r, e)

        # ---- gffm_bplan_gemm: per range, in the order own-first: wait for the planes, GEMM on the context stream ----
        for i in range(nr):
            qq = (r - i) % nr if tr == "raw" else (r + i) % nr  # out->order
            if cnt[qq] <= 0:
                continue
            w.wait_event(r, "ctx", f"ready{b}_{qq}")
            w.work(r, "ctx", f"gemm e{e} range {qq}", P(r, b, qq), [], expect=version)
        # ---- mg_round_done ----
        w.record(r, "ctx", f"gemm_done{b}")
        if tr == "push":
            w.wait_event(r, "pull1", f"gemm_done{b}")
            for qq in range(nr):
                if qq != r:
                    w.write_flag(r, "pull1", qq, "FREE", r, e)
        if tr == "raw":
            if (r == root or distributed) and self.broken != "no_caller_fence":
                w.wait_event(r, "ctx", "ev_push")
                w.wait_event(r, "ctx", f"split_ev{b}")
        elif r == root and self.broken != "no_caller_fence":
            w.wait_event(r, "ctx", "ev_push")
        w.wait_event(r, "ctx", f"ready{b}_{r}")


def simulate(nr, transport, root=0, products=7, caller="rewrite", blocks=None, root_free_min=6, seed=0, broken="", roots=None):
    """`products` sharded products back to back.  caller: "rewrite" = B is rewritten (context stream) before every product and declared
    ready by a per-product event; "const" = B written once, one ready event for all products; "inorder" = no event (b_ready = NULL).
    roots: optional per-product list of roots (any rank may be the root of a later product)."""
    w = World(nr, seed)
    blocks = blocks if blocks is not None else 2 * nr + 1
    ranks = [Rank(w, r, transport, root_free_min, broken) for r in range(nr)]
    holders = lambda rt: list(range(nr)) if rt < 0 else [rt]  # noqa: E731  ranks whose matrix carries the data
    for it in range(products):
        rt = roots[it] if roots else root
        version = it + 1 if caller != "const" else 1
        for r in range(nr):
            ev = None
            if r in holders(rt) and (caller != "const" or it == 0 or roots):
                regs = [("B", r, qq) for qq in range(nr)] if rt >= 0 else [("B", r, r)]
                w.work(r, "ctx", f"caller writes B v{version}", [], regs, tag=version)
            if caller in ("rewrite", "const"):
                if caller == "rewrite" or it == 0 or roots:
                    w.record(r, "ctx", f"bready_{it if caller == 'rewrite' or roots else 0}")
                ev = f"bready_{it if caller == 'rewrite' or roots else 0}"
            ranks[r].product(rt, ev, blocks, version)
    return w.run()
