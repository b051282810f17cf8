"""ncu target: a few modular GEMVs at the metric size (n = 16384, 25-bit modulus) through the C ABI."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import gffm_b200 as g  # noqa: E402

n, N = 16384, 33554393
A = g.synth(n, n, N, 5); x = g.synth(n, 1, N, 77); z = g.zeros(np.float32, n, 1, N)
for _ in range(3):
    g.gemv_(z, A, x)
g.default_context().sync()
print("gemv done", z.checksum())
