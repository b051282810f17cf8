"""CPU unit test of multigpu.BroadcastMatmul's stream/event protocol with a recording stand-in for torch.cuda and for the C
ABI: what is enqueued on which stream, in which order, and which raw handles reach gffm_gemm_panels.  (The real thing runs
in tests/test_gpu_parity.py::test_broadcast_matmul_streamed_panels and in bench.py --gpus N.)"""
import ctypes
import types

LOG = []


class FakeStream:
    _n = 0

    def __init__(self, device=None, name=None):
        FakeStream._n += 1
        self.name = name or f"s{FakeStream._n}"

    def wait_stream(self, other):
        LOG.append(("wait_stream", self.name, other.name))

    def wait_event(self, ev):
        LOG.append(("wait_event", self.name, ev.id))


class FakeEvent:
    _n = 0

    def __init__(self, enable_timing=False):
        FakeEvent._n += 1
        self.id = FakeEvent._n
        self.recorded = False

    def record(self, stream):
        self.recorded = True
        LOG.append(("record", stream.name, self.id))

    @property
    def cuda_event(self):
        assert self.recorded, "raw handle read before the first record (torch creates the event lazily)"
        return 0x1000 + self.id


class FakeCuda:
    def __init__(self):
        self._cur = FakeStream(name="compute")
        self.Stream = FakeStream
        self.Event = FakeEvent

    def current_stream(self):
        return self._cur

    def stream(self, s):
        outer = self

        class _Ctx:
            def __enter__(self_inner):
                self_inner.prev = outer._cur
                outer._cur = s

            def __exit__(self_inner, *a):
                outer._cur = self_inner.prev

        return _Ctx()


class FakeLib:
    def __init__(self):
        self.calls = []

    def gffm_gemm_panels(self, C, A, B, npanels, off, ready, consumed, R, P):
        self.calls.append({"C": C, "A": A, "B": B, "n": npanels, "off": list(off), "ready": [ready[i] for i in range(npanels)],
                           "consumed": [consumed[i] for i in range(npanels)], "R": R, "P": P})
        LOG.append(("gemm_panels", npanels))
        return 0


class FakeMat:
    def __init__(self, h, lib):
        self.h, self.lib = h, lib


class FakeTensor:
    device = "cuda:0"

    def __getitem__(self, sl):
        return ("slice", sl.start, sl.stop)


def test_broadcast_matmul_event_protocol():
    import gffm_b200 as g
    del LOG[:]
    cuda = FakeCuda()
    torch = types.SimpleNamespace(cuda=cuda)
    lib = FakeLib()
    delivered = []

    def deliver(c0, c1):
        LOG.append(("deliver", cuda.current_stream().name, c0, c1))
        delivered.append((c0, c1))

    panels = [(0, 512), (512, 1024), (1024, 1300)]
    bm = g.multigpu.BroadcastMatmul(torch, None, FakeMat(11, lib), FakeMat(12, lib), FakeMat(13, lib), FakeTensor(), panels, deliver=deliver)
    comm = bm.comm.name
    bm.step()
    first = list(LOG)
    # the first delivery is ordered after whatever the caller's stream did to the buffers, and every consumed event exists
    assert first[0] == ("wait_stream", comm, "compute")
    cons_ids = [e.id for e in bm.consumed]; ready_ids = [e.id for e in bm.ready]
    assert first[1:4] == [("record", "compute", i) for i in cons_ids]
    # per panel, on the communication stream: wait consumed[p] -> deliver -> record ready[p]
    body = first[4:4 + 9]
    for p, (c0, c1) in enumerate(panels):
        assert body[3 * p:3 * p + 3] == [("wait_event", comm, cons_ids[p]), ("deliver", comm, c0, c1), ("record", comm, ready_ids[p])]
    assert first[-1] == ("gemm_panels", 3) and cuda.current_stream().name == "compute"
    call = lib.calls[0]
    assert (call["C"], call["A"], call["B"]) == (11, 12, 13) and call["off"] == [0, 512, 1024, 1300]
    assert call["ready"] == [0x1000 + i for i in ready_ids] and call["consumed"] == [0x1000 + i for i in cons_ids]
    assert (call["R"], call["P"]) == (0, 0)
    # second step: no stream-wide wait any more -- only the per-panel consumed events gate the next broadcast
    del LOG[:]
    bm.step()
    assert not any(op[0] == "wait_stream" for op in LOG)
    assert [op for op in LOG if op[0] == "wait_event"] == [("wait_event", comm, i) for i in cons_ids]
    assert delivered == panels * 2
    del LOG[:]
    bm.finish()
    assert LOG == [("wait_stream", "compute", comm)]


def test_broadcast_matmul_collective_choice():
    import gffm_b200 as g
    import pytest
    cuda = FakeCuda()
    torch = types.SimpleNamespace(cuda=cuda)
    sent = []
    dist = types.SimpleNamespace(broadcast=lambda t, src=0, group=None: sent.append(("bcast", t, src)))
    lib = FakeLib()
    bm = g.multigpu.BroadcastMatmul(torch, dist, FakeMat(1, lib), FakeMat(2, lib), FakeMat(3, lib), FakeTensor(), [(0, 256), (256, 300)], src=0)
    bm.step()
    assert sent == [("bcast", ("slice", 0, 256), 0), ("bcast", ("slice", 256, 300), 0)]
    with pytest.raises(ValueError):
        g.multigpu.BroadcastMatmul(torch, dist, FakeMat(1, lib), FakeMat(2, lib), FakeMat(3, lib), FakeTensor(), [(0, 256)], collective="ring")
