import sys, os
sys.path.insert(0, os.getcwd())
import gffm_b200 as g
n, N = int(sys.argv[1]), int(sys.argv[2])
A = g.synth(n, n, N, 9)
U, L, pr, pc, rk = g.pluq_gpu_kernel(A, return_rank=True)
g.default_context().sync()
print("rank", rk)
