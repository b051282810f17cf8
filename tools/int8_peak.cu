// int8_peak.cu -- microbenchmark of the sm_100a int8 tensor pipe: every SM issues back-to-back
// tcgen05.mma.cta_group::1.kind::i8 128x256x32 from operands that stay resident in shared memory (no TMA, no
// epilogue), so the number is the issue-rate ceiling of the pipe under the board's power limit.  It replaces the
// "2 x bf16" proxy as the roofline denominator of the modular GEMM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/int8_peak tools/int8_peak.cu && tools/int8_peak
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../gpufinitefieldmatrices.jl_b200/csrc/tc_ptx.cuh"

constexpr int BM = 128, BN = 256, KB = 128;

__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, int commit_every, int random_data) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (BM + BN) * KB);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (BM + BN) * KB / 4; i += blockDim.x) {
    uint32_t v = 0x01010101u;
    if (random_data) {  // uniformly random bytes: the operand toggling (and so the power) of real residue planes
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u + 12345u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
      v = h;
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(slot, 512);
    tc::tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA (async proxy)
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc = tc::make_idesc_i8(BM, BN, true, true);
    const uint32_t sa = tc::smem_u32(smem), sb = sa + BM * KB;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da = tc::make_smem_desc_sw128(sa + kk * 32), db = tc::make_smem_desc_sw128(sb + kk * 32);
        tc::mma_i8_ss(tmem + (it & 1) * BN, da, db, idesc, (it > 1 || kk > 0) ? 1u : 0u);
      }
      if ((it + 1) % commit_every == 0 || it == iters - 1) {
        tc::mma_commit(bar);
        tc::mbar_wait(bar, phase);
        phase ^= 1;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
  int dev = 0;
  cudaSetDevice(dev);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  const int sms = prop.multiProcessorCount;
  const int smem = (BM + BN) * KB + 1024 + 64;
  cudaFuncSetAttribute(peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = argc > 1 ? atoi(argv[1]) : 200000;  // x4 MMAs each
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0, sustained = 0, best_rnd = 0, sustained_rnd = 0;
  for (int pattern = 0; pattern < 2; ++pattern) {
    for (int rep = 0; rep < 10; ++rep) {
      const int it = rep < 3 ? iters / 10 : iters;  // short bursts first, then long (power-capped) runs back to back
      cudaEventRecord(e0);
      peak_kernel<<<sms, 128, smem>>>(it, 64, pattern);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double tops = (double)sms * it * 4 * 2.0 * BM * BN * 32 / (ms * 1e-3) / 1e12;
      double& b = pattern ? best_rnd : best;
      double& su = pattern ? sustained_rnd : sustained;
      if (tops > b) b = tops;
      if (rep >= 3) su = tops;
      fprintf(stderr, "pattern %d rep %d: %d x4 MMAs/SM in %.3f ms -> %.1f TOP/s\n", pattern, rep, it, ms, tops);
    }
  }
  cudaError_t err = cudaGetLastError();
  printf("{\"int8_tops_burst\": %.1f, \"int8_tops_sustained\": %.1f, \"int8_tops_burst_random_data\": %.1f, \"int8_tops_sustained_random_data\": %.1f, \"sms\": %d, \"shape\": \"tcgen05.mma.cta_group::1.kind::i8 128x256x32, operands resident in smem\", \"error\": \"%s\"}\n",
         best, sustained, best_rnd, sustained_rnd, sms, cudaGetErrorString(err));
  return 0;
}
