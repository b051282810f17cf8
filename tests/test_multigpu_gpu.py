"""GPU tests of the C-ABI multi-GPU layer (gffm_mg_*, csrc/mg.cu).  The single-rank cases run on any B200 box; the multi-rank
self-test (tools/mg_selftest.py under torchrun: every transport, bit-exact vs the CPU oracle) needs at least two GPUs and is skipped
otherwise."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from oracle import oracle as O  # noqa: E402  (checker only)
from oracle import oracle_c as OC  # noqa: E402


@pytest.fixture(scope="module")
def g():
    import gffm_b200
    gffm_b200.default_context()
    return gffm_b200


@pytest.mark.parametrize("transport", ["MG_NCCL_BCAST", "MG_NCCL_PLANES", "MG_P2P_PLANES", "MG_P2P_PUSH", "MG_P2P_RAW", "MG_AUTO"])
def test_single_rank_products_through_the_mg_layer(g, transport):
    """One rank: the whole data flow (owner split into the plane arena, external-plane GEMM per column range, CRT) with a
    1-rank NCCL communicator; results bit-exact vs the oracle for RNS / two-limb / one-limb moduli, Karatsuba and mat-vec."""
    ctx = g.default_context()
    mgpu = g.multigpu.MultiGpu(ctx, 0, 1, g.multigpu.MultiGpu.unique_id())
    mgpu.set_transport(getattr(g.capi, transport))
    for (m, k, n, N) in [(300, 200, 700, 33554393), (513, 640, 1300, 11), (640, 256, 1000, 65521), (100, 17000, 300, 65521), (128, 128, 5, 33554393)]:
        Ah = O.synth_matrix(1, m, k, N); Bh = O.synth_matrix(2, k, n, N)
        C = g.zeros(np.float32, m, n, N)
        mgpu.gemm(C, g.CuModMatrix(Ah, N), g.CuModMatrix(Bh, N), root=0)
        assert np.array_equal(C.to_int(), OC.matmul_mod(Ah, Bh, N)), (m, k, n, N)
    N1, N2 = 8191, 8191
    M = N1 * N2
    rng = np.random.default_rng(3)
    Ak = rng.integers(0, M, size=(300, 256), dtype=np.int64); Bk = rng.integers(0, M, size=(256, 520), dtype=np.int64)
    CK = g.KaratsubaZeros(np.float64, 300, 520, N1, N2)
    mgpu.kmat_mul(CK, g.KaratsubaMatrix.from_array(Ak, N1, N2, M), g.KaratsubaMatrix.from_array(Bk, N1, N2, M))
    assert np.array_equal(np.asarray(CK.Array()).astype(np.int64), OC.matmul_mod(Ak, Bk, M, in_bound=M))
    Ah = O.synth_matrix(5, 700, 333, 33554393); xh = O.synth_matrix(6, 333, 1, 33554393)
    z = g.zeros(np.float32, 700, 1, 33554393)
    mgpu.gemv(z, g.CuModMatrix(Ah, 33554393), g.CuModMatrix(xh, 33554393))
    assert np.array_equal(z.to_int().reshape(-1), OC.matmul_mod(Ah, xh, 33554393).reshape(-1))
    # a pipelined sequence on alternating plane buffers
    A = g.synth(1024, 512, 33554393, 7); B = g.synth(512, 1536, 33554393, 8)
    want = OC.matmul_mod(A.to_int(), B.to_int(), 33554393)
    outs = [g.zeros(np.float32, 1024, 1536, 33554393) for _ in range(5)]
    for Cq in outs:
        A.touch(); mgpu.gemm(Cq, A, B)
    assert all(np.array_equal(Cq.to_int(), want) for Cq in outs)
    mgpu.barrier()
    info = mgpu.info()
    assert info["nranks"] == 1 and info["rank"] == 0
    # error behaviour of mul! is kept (CuModMatrix.jl:769-783)
    with pytest.raises(g.CuModArrayModulusMismatchException):
        mgpu.gemm(g.zeros(np.float32, 4, 4, 7), g.zeros(np.float32, 4, 4, 11), g.zeros(np.float32, 4, 4, 7))
    with pytest.raises(g.CuModArraySizeMismatchException):
        mgpu.gemm(g.zeros(np.float32, 4, 4, 7), g.zeros(np.float32, 4, 5, 7), g.zeros(np.float32, 4, 4, 7))
    mgpu.close()


def test_multi_rank_selftest_under_torchrun():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs (run on a multi-GPU box: gpurun --gpus 2)")
    nproc = 2 if ngpu < 4 else 4
    env = dict(os.environ)
    env.pop("NCCL_DEBUG", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "mg_selftest.py")], capture_output=True, text=True, cwd=ROOT, env=env, timeout=1200)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["failed"] == [] and d["checks"] > 20
