// gemm_tc.cu -- exact modular GEMM on the 5th-gen tensor cores (tcgen05 kind::i8, int32 accumulators in TMEM,
// operands staged by TMA with the 128-byte swizzle).  Replaces the reference's cuBLAS S/DGEMM K-stripe loop
// (reference src/CuModMatrix/kernel_mul/stripe_mul.jl:175-244) and its per-stripe k_mod! pass
// (kernel_ops/mod_ops.jl:3-27): one launch covers all of K, the modular reduction is fused into the epilogue.
//
// Two exact encodings of residues as 8-bit planes ("limbs"):
//   LIMB  positional base-256 digits (inputs < 2^16): L in {1,2} unsigned planes per operand, L^2 MMAs per
//         K-step accumulated by weight i+j into 2L-1 TMEM accumulators; epilogue recombines sum_w acc_w * 2^(8w)
//         in 64 bit, one Barrett reduction mod P, optional fused C +/-= (GFFM_GEMM_ADD / SUB).
//   RNS   residue limbs: balanced residues mod s pairwise-coprime 8-bit moduli m_t (256, 255, 253, ...);
//         s independent signed-int8 GEMMs (one TMEM accumulator each) whose epilogue reduces mod m_t and
//         stores one byte per element; a CRT kernel (HBM-bound, 1 pass) reconstructs the exact integer dot
//         product mod P.  For 17..26-bit moduli this needs 8-9 MMAs per K-step instead of the 9-16 of
//         positional limbs and only ONE accumulator, so the 128x256 tile + double-buffered TMEM stays available.
//
// Kernel anatomy (one CTA per SM, persistent over output tiles, 192 threads):
//   warp 0      TMA producer      cp.async.bulk.tensor.3d -> smem ring (full/empty mbarriers)
//   warp 1      MMA issuer        tcgen05.mma.cta_group::1.kind::i8, tcgen05.commit -> empty[stage] / tmem_full[buf]
//   warps 2..5  epilogue          tcgen05.ld 32x32b.x16 -> registers -> reduction -> coalesced global stores
#include <cuda.h>
#include <algorithm>
#include <vector>
#include <stdlib.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;        // UMMA M
constexpr int BK_BYTES = 128;  // one 128-byte swizzle atom along K per stage
constexpr int UMMA_K = 32;     // K per tcgen05.mma for 8-bit inputs
constexpr int MAX_MODS = 15;
constexpr int GROUP_M = 16;    // rasterisation: 16 m-blocks x (148/16) n-blocks run concurrently -> L2 reuse

enum { EPI_POS = 0, EPI_RNS = 1 };

struct RnsModDev {
  uint32_t m;    // modulus
  uint32_t mu;   // floor(2^32 / m)
  uint32_t off;  // multiple of m, >= 2^31 : makes the int32 accumulator non-negative
  uint32_t u;    // (M/m)^-1 mod m : CRT pre-scaling applied in the epilogue
};

struct GemmParams {
  int m, n;
  int num_kb;
  int num_m_blk, num_n_blk, batches;
  int fake_loads;  // DEBUG (GFFM_FAKE_LOADS=1): after the first ring fill the producer only signals the barriers -- measures the
                   // tensor pipe without L2->smem traffic (results are garbage)
  int l2_hints;  // 1: A strips (re-read by every n-block of a raster group) EVICT_LAST, B panels (streamed) EVICT_FIRST
  int mb0, nb0;  // tile offsets (in blocks) into the operand planes: sub-problems of a larger plane set (pipelined host GEMM)
  // positional epilogue
  uint32_t* C;
  int64_t ldc;
  int mode;
  ModP modP;
  // RNS epilogue
  uint8_t* E;
  int64_t lde;
  int64_t e_plane_stride;
  RnsModDev mods[MAX_MODS];
};

// ---- schemes ---------------------------------------------------------------------------------------
struct SchemeL1 {  // inputs < 2^8 : one unsigned plane, one accumulator, 128x256 tile, 2 TMEM buffers
  static constexpr int PA = 1, PB = 1, NSLOT = 1, BN = 256, NPROD = 1, STAGES = 4, EPI = EPI_POS;
  static constexpr bool SIGNED = false;
  __host__ __device__ static constexpr int pa(int) { return 0; }
  __host__ __device__ static constexpr int pb(int) { return 0; }
  __host__ __device__ static constexpr int slot(int) { return 0; }
};
struct SchemeL2 {  // inputs < 2^16: two unsigned planes, weights 0,1,2 -> 3 accumulators x 128 columns
  static constexpr int PA = 2, PB = 2, NSLOT = 3, BN = 128, NPROD = 4, STAGES = 3, EPI = EPI_POS;
  static constexpr bool SIGNED = false;
  __host__ __device__ static constexpr int pa(int i) { return i >> 1; }
  __host__ __device__ static constexpr int pb(int i) { return i & 1; }
  __host__ __device__ static constexpr int slot(int i) { return (i >> 1) + (i & 1); }
};
struct SchemeRNS {  // one signed plane per modulus, batch index = modulus
  static constexpr int PA = 1, PB = 1, NSLOT = 1, BN = 256, NPROD = 1, STAGES = 4, EPI = EPI_RNS;
  static constexpr bool SIGNED = true;
  __host__ __device__ static constexpr int pa(int) { return 0; }
  __host__ __device__ static constexpr int pb(int) { return 0; }
  __host__ __device__ static constexpr int slot(int) { return 0; }
};

template <class S>
struct Cfg {
  static constexpr int A_TILE = BM * BK_BYTES;
  static constexpr int B_TILE = S::BN * BK_BYTES;
  static constexpr int STAGE_BYTES = S::PA * A_TILE + S::PB * B_TILE;
  static constexpr int ACC_COLS = S::NSLOT * S::BN;
  static constexpr int NBUF = (512 / ACC_COLS) >= 2 ? 2 : 1;
  static constexpr int SMEM_BYTES = S::STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void decode_tile(int tile, const GemmParams& p, int& z, int& mb, int& nb) {
  const int per_batch = p.num_m_blk * p.num_n_blk;
  z = tile / per_batch;
  int t = tile - z * per_batch;
  const int per_group = GROUP_M * p.num_n_blk;
  const int g = t / per_group;
  const int first_m = g * GROUP_M;
  const int gsize = min(p.num_m_blk - first_m, GROUP_M);
  const int r = t - g * per_group;
  mb = first_m + (r % gsize);
  nb = r / gsize;
}

template <class S>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmParams p) {
  using C = Cfg<S>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tfull_bar = empty_bar + S::STAGES;
  uint64_t* tempty_bar = tfull_bar + C::NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + C::NBUF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    for (int s = 0; s < S::STAGES; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < C::NBUF; ++b) {
      tc::mbar_init(&tfull_bar[b], 1);
      tc::mbar_init(&tempty_bar[b], 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.batches * p.num_m_blk * p.num_n_blk;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int z, mb, nb;
        decode_tile(tile, p, z, mb, nb);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (p.fake_loads == 1 && (tile != (int)blockIdx.x || kb >= S::STAGES)) {
            tc::mbar_arrive(&full_bar[stage]);
            if (++stage == S::STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          tc::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          uint8_t* st = smem + stage * C::STAGE_BYTES;
          if (p.fake_loads == 2) {  // DEBUG: real TMA traffic but always the same L2-resident tiles (no DRAM traffic)
#pragma unroll
            for (int a = 0; a < S::PA; ++a) tc::tma_load_3d(st + a * C::A_TILE, &tmA, &full_bar[stage], 0, (blockIdx.x & 7) * BM, a);
#pragma unroll
            for (int b = 0; b < S::PB; ++b) tc::tma_load_3d(st + S::PA * C::A_TILE + b * C::B_TILE, &tmB, &full_bar[stage], 0, (blockIdx.x & 7) * S::BN, b);
            if (++stage == S::STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
#pragma unroll
          for (int a = 0; a < S::PA; ++a)
            if (p.l2_hints) tc::tma_load_3d_hint(st + a * C::A_TILE, &tmA, &full_bar[stage], kb * BK_BYTES, (mb + p.mb0) * BM, z * S::PA + a, tc::kEvictLast);
            else tc::tma_load_3d(st + a * C::A_TILE, &tmA, &full_bar[stage], kb * BK_BYTES, (mb + p.mb0) * BM, z * S::PA + a);
#pragma unroll
          for (int b = 0; b < S::PB; ++b)
            if (p.l2_hints) tc::tma_load_3d_hint(st + S::PA * C::A_TILE + b * C::B_TILE, &tmB, &full_bar[stage], kb * BK_BYTES, (nb + p.nb0) * S::BN,
                                                 z * S::PB + b, tc::kEvictFirst);
            else tc::tma_load_3d(st + S::PA * C::A_TILE + b * C::B_TILE, &tmB, &full_bar[stage], kb * BK_BYTES, (nb + p.nb0) * S::BN,
                                 z * S::PB + b);
          if (++stage == S::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = tc::make_idesc_i8(BM, S::BN, S::SIGNED, S::SIGNED);
    int stage = 0;
    uint32_t phase = 0;
    int buf = 0;
    uint32_t bphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      tc::mbar_wait(&tempty_bar[buf], bphase ^ 1);  // epilogue has drained this accumulator buffer
      tc::tc_fence_after();
      const uint32_t acc_base = tmem_base + buf * C::ACC_COLS;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        tc::mbar_wait(&full_bar[stage], phase);  // TMA bytes have landed
        tc::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = tc::smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + S::PA * C::A_TILE;
#pragma unroll
          for (int kk = 0; kk < BK_BYTES / UMMA_K; ++kk) {
#pragma unroll
            for (int pr = 0; pr < S::NPROD; ++pr) {
              const uint64_t da = tc::make_smem_desc_sw128(sa + S::pa(pr) * C::A_TILE + kk * UMMA_K);
              const uint64_t db = tc::make_smem_desc_sw128(sb + S::pb(pr) * C::B_TILE + kk * UMMA_K);
              // first MMA that touches a slot in this tile overwrites, everything else accumulates
              bool first_in_slot = true;
#pragma unroll
              for (int q = 0; q < pr; ++q)
                if (S::slot(q) == S::slot(pr)) first_in_slot = false;
              const uint32_t accumulate = (kb > 0 || kk > 0 || !first_in_slot) ? 1u : 0u;
              tc::mma_i8_ss(acc_base + S::slot(pr) * S::BN, da, db, idesc, accumulate);
            }
          }
          tc::mma_commit(&empty_bar[stage]);                       // smem slot reusable once these MMAs retire
          if (kb == p.num_kb - 1) tc::mma_commit(&tfull_bar[buf]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == S::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++buf == C::NBUF) {
        buf = 0;
        bphase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    int buf = 0;
    uint32_t bphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int z, mb, nb;
      decode_tile(tile, p, z, mb, nb);
      tc::mbar_wait(&tfull_bar[buf], bphase);
      tc::tc_fence_after();
      const int row = mb * BM + q * 32 + lane;
      const bool row_ok = row < p.m;
      const uint32_t acc_base = tmem_base + buf * C::ACC_COLS + lane_addr;
      const int col0 = nb * S::BN;
#pragma unroll 1
      for (int c0 = 0; c0 < S::BN; c0 += 16) {
        if (col0 + c0 >= p.n) break;  // warp-uniform
        uint32_t v[S::NSLOT][16];
#pragma unroll
        for (int s = 0; s < S::NSLOT; ++s) tc::tmem_ld16(acc_base + s * S::BN + c0, v[s]);
        tc::tmem_ld_wait();
        if constexpr (S::EPI == EPI_POS) {
          // fused C +/-= : all 16 old values are requested before any is used (16 loads in flight per thread)
          uint32_t old[16];
          uint32_t* dst0 = p.C + (int64_t)(col0 + c0) * p.ldc + row;
          if (p.mode != GFFM_GEMM_STORE) {
#pragma unroll
            for (int c = 0; c < 16; ++c) old[c] = (row_ok && col0 + c0 + c < p.n) ? dst0[(int64_t)c * p.ldc] : 0u;
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            uint64_t acc = (uint64_t)v[0][c];
            if constexpr (S::NSLOT == 3) acc += ((uint64_t)v[1][c] << 8) + ((uint64_t)v[2][c] << 16);
            uint32_t r = (uint32_t)mod_u64(acc, p.modP);
            if (p.mode == GFFM_GEMM_ADD) r = addmod_u32(old[c], r, (uint32_t)p.modP.P);
            else if (p.mode == GFFM_GEMM_SUB) r = submod_u32(old[c], r, (uint32_t)p.modP.P);
            old[c] = r;
          }
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (row_ok && col0 + c0 + c < p.n) dst0[(int64_t)c * p.ldc] = old[c];
        } else {
          const RnsModDev md = p.mods[z];
          uint8_t* eplane = p.E + (int64_t)z * p.e_plane_stride;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int col = col0 + c0 + c;
            const uint32_t u = v[0][c] + md.off;  // == acc (mod m), non-negative
            uint32_t e = u - __umulhi(u, md.mu) * md.m;
            if (e >= md.m) e -= md.m;
            e *= md.u;  // < 2^16
            e -= __umulhi(e, md.mu) * md.m;
            if (e >= md.m) e -= md.m;
            if (row_ok && col < p.n) eplane[(int64_t)col * p.lde + row] = (uint8_t)e;
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty_bar[buf]);
      if (++buf == C::NBUF) {
        buf = 0;
        bphase ^= 1;
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// Cluster variant (2 CTAs along M share every B tile): each CTA loads its own A tile and HALF of the B tile, the half is
// TMA-multicast into both CTAs.  L2->SM traffic per MMA drops by a third (48 -> 32 KB per 128x256x128 k-block), which on a
// power-capped B200 buys SM clock (profiles/r01_notes.md: the same kernel without operand traffic runs 22 % faster).
// Both CTAs walk the same (tile, k-block) sequence; a stage is reusable only when BOTH CTAs' MMAs have retired it, so the
// MMA commit is multicast to both empty barriers (count 2).
// ---------------------------------------------------------------------------------------------------
template <class S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
gemm_tc_kernel_mc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                  const __grid_constant__ GemmParams p) {
  using C = Cfg<S>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tfull_bar = empty_bar + S::STAGES;
  uint64_t* tempty_bar = tfull_bar + C::NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + C::NBUF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = tc::cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  constexpr int B_HALF = C::B_TILE / 2;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmBh);
    for (int s = 0; s < S::STAGES; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 2);  // both CTAs of the cluster retire the stage
    }
    for (int b = 0; b < C::NBUF; ++b) {
      tc::mbar_init(&tfull_bar[b], 1);
      tc::mbar_init(&tempty_bar[b], 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();  // peer barriers are initialised before any multicast can reach them
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // cluster tiles: pairs of m-blocks (2*mp, 2*mp+1) x one n-block, rasterised GROUP_M/2 pairs wide
  const int num_mp = (p.num_m_blk + 1) >> 1;
  const int total = p.batches * num_mp * p.num_n_blk;
  auto decode = [&](int t, int& z, int& mb, int& nb) {
    const int per_batch = num_mp * p.num_n_blk;
    z = t / per_batch;
    int r = t - z * per_batch;
    const int gw = GROUP_M / 2;
    const int per_group = gw * p.num_n_blk;
    const int g = r / per_group;
    const int first = g * gw;
    const int gsize = min(num_mp - first, gw);
    r -= g * per_group;
    mb = 2 * (first + (r % gsize)) + (int)crank;
    nb = r / gsize;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total; tile += num_clusters) {
        int z, mb, nb;
        decode(tile, z, mb, nb);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          tc::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          uint8_t* st = smem + stage * C::STAGE_BYTES;
#pragma unroll
          for (int a = 0; a < S::PA; ++a)
            tc::tma_load_3d(st + a * C::A_TILE, &tmA, &full_bar[stage], kb * BK_BYTES, (mb + p.mb0) * BM, z * S::PA + a);
#pragma unroll
          for (int b = 0; b < S::PB; ++b)
            tc::tma_load_3d_mcast(st + S::PA * C::A_TILE + b * C::B_TILE + crank * B_HALF, &tmBh, &full_bar[stage], kb * BK_BYTES,
                                  (nb + p.nb0) * S::BN + (int)crank * (S::BN / 2), z * S::PB + b, (uint16_t)0x3);
          if (++stage == S::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = tc::make_idesc_i8(BM, S::BN, S::SIGNED, S::SIGNED);
    int stage = 0;
    uint32_t phase = 0;
    int buf = 0;
    uint32_t bphase = 0;
    for (int tile = cluster_id; tile < total; tile += num_clusters) {
      tc::mbar_wait(&tempty_bar[buf], bphase ^ 1);
      tc::tc_fence_after();
      const uint32_t acc_base = tmem_base + buf * C::ACC_COLS;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        tc::mbar_wait(&full_bar[stage], phase);
        tc::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = tc::smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + S::PA * C::A_TILE;
#pragma unroll
          for (int kk = 0; kk < BK_BYTES / UMMA_K; ++kk) {
#pragma unroll
            for (int pr = 0; pr < S::NPROD; ++pr) {
              const uint64_t da = tc::make_smem_desc_sw128(sa + S::pa(pr) * C::A_TILE + kk * UMMA_K);
              const uint64_t db = tc::make_smem_desc_sw128(sb + S::pb(pr) * C::B_TILE + kk * UMMA_K);
              bool first_in_slot = true;
#pragma unroll
              for (int q = 0; q < pr; ++q)
                if (S::slot(q) == S::slot(pr)) first_in_slot = false;
              const uint32_t accumulate = (kb > 0 || kk > 0 || !first_in_slot) ? 1u : 0u;
              tc::mma_i8_ss(acc_base + S::slot(pr) * S::BN, da, db, idesc, accumulate);
            }
          }
          tc::mma_commit_mcast(&empty_bar[stage], (uint16_t)0x3);  // stage retired here -> tell both producers
          if (kb == p.num_kb - 1) tc::mma_commit(&tfull_bar[buf]);
        }
        __syncwarp();
        if (++stage == S::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++buf == C::NBUF) {
        buf = 0;
        bphase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    int buf = 0;
    uint32_t bphase = 0;
    for (int tile = cluster_id; tile < total; tile += num_clusters) {
      int z, mb, nb;
      decode(tile, z, mb, nb);
      tc::mbar_wait(&tfull_bar[buf], bphase);
      tc::tc_fence_after();
      const int row = mb * BM + q * 32 + lane;
      const bool row_ok = row < p.m;
      const uint32_t acc_base = tmem_base + buf * C::ACC_COLS + lane_addr;
      const int col0 = nb * S::BN;
#pragma unroll 1
      for (int c0 = 0; c0 < S::BN; c0 += 16) {
        if (col0 + c0 >= p.n) break;
        uint32_t v[S::NSLOT][16];
#pragma unroll
        for (int s = 0; s < S::NSLOT; ++s) tc::tmem_ld16(acc_base + s * S::BN + c0, v[s]);
        tc::tmem_ld_wait();
        if constexpr (S::EPI == EPI_POS) {
          uint32_t old[16];
          uint32_t* dst0 = p.C + (int64_t)(col0 + c0) * p.ldc + row;
          if (p.mode != GFFM_GEMM_STORE) {
#pragma unroll
            for (int c = 0; c < 16; ++c) old[c] = (row_ok && col0 + c0 + c < p.n) ? dst0[(int64_t)c * p.ldc] : 0u;
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            uint64_t acc = (uint64_t)v[0][c];
            if constexpr (S::NSLOT == 3) acc += ((uint64_t)v[1][c] << 8) + ((uint64_t)v[2][c] << 16);
            uint32_t r = (uint32_t)mod_u64(acc, p.modP);
            if (p.mode == GFFM_GEMM_ADD) r = addmod_u32(old[c], r, (uint32_t)p.modP.P);
            else if (p.mode == GFFM_GEMM_SUB) r = submod_u32(old[c], r, (uint32_t)p.modP.P);
            old[c] = r;
          }
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (row_ok && col0 + c0 + c < p.n) dst0[(int64_t)c * p.ldc] = old[c];
        } else {
          const RnsModDev md = p.mods[z];
          uint8_t* eplane = p.E + (int64_t)z * p.e_plane_stride;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int col = col0 + c0 + c;
            const uint32_t u = v[0][c] + md.off;
            uint32_t e = u - __umulhi(u, md.mu) * md.m;
            if (e >= md.m) e -= md.m;
            e *= md.u;
            e -= __umulhi(e, md.mu) * md.m;
            if (e >= md.m) e -= md.m;
            if (row_ok && col < p.n) eplane[(int64_t)col * p.lde + row] = (uint8_t)e;
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty_bar[buf]);
      if (++buf == C::NBUF) {
        buf = 0;
        bphase ^= 1;
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();  // no CTA may leave while its peer can still multicast into it
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): the two CTAs of a cluster compute one 256 x BN tile.  Each CTA stages its own
// 128 rows of A and only HALF of the B tile (BN/2 rows); the leader issues M = 256 MMAs that read both shared memories.
// Shared-memory fill traffic per MMA drops from 12 KB to 8 KB per SM -- the resource the traffic experiments in
// profiles/r01_notes.md identify as the limiter of the single-CTA kernel under the power cap.
//   full[s]   leader only; armed by the leader's producer for the bytes of BOTH CTAs, credited by both CTAs' TMA loads
//   empty[s]  both CTAs; released by the leader's tcgen05.commit multicast
//   tfull[b]  both CTAs; leader's commit multicast -> each CTA's epilogue drains its own 128 TMEM lanes
//   tempty[b] leader only, 8 arrivals: 4 local epilogue warps + 4 remote (peer) epilogue warps
// ---------------------------------------------------------------------------------------------------
template <class S>
struct Cfg2 {
  static constexpr int A_TILE = BM * BK_BYTES;
  static constexpr int BH_TILE = (S::BN / 2) * BK_BYTES;
  static constexpr int STAGE_BYTES = S::PA * A_TILE + S::PB * BH_TILE;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 6 ? 6 : (200 * 1024) / STAGE_BYTES;
  static constexpr int ACC_COLS = S::NSLOT * S::BN;
  static constexpr int NBUF = (512 / ACC_COLS) >= 2 ? 2 : 1;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <class S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
gemm_tc_kernel_2cta(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                    const __grid_constant__ GemmParams p) {
  using C = Cfg2<S>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tfull_bar = empty_bar + C::STAGES;
  uint64_t* tempty_bar = tfull_bar + C::NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + C::NBUF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = tc::cluster_ctarank();
  const bool leader = crank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmBh);
    for (int s = 0; s < C::STAGES; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < C::NBUF; ++b) {
      tc::mbar_init(&tfull_bar[b], 1);
      tc::mbar_init(&tempty_bar[b], 8);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc_2sm(tmem_slot, 512);
    tc::tmem_relinquish_2sm();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_mp = (p.num_m_blk + 1) >> 1;
  const int total = p.batches * num_mp * p.num_n_blk;
  auto decode = [&](int t, int& z, int& mb, int& nb) {
    const int per_batch = num_mp * p.num_n_blk;
    z = t / per_batch;
    int r = t - z * per_batch;
    const int gw = GROUP_M / 2;
    const int per_group = gw * p.num_n_blk;
    const int g = r / per_group;
    const int first = g * gw;
    const int gsize = min(num_mp - first, gw);
    r -= g * per_group;
    mb = 2 * (first + (r % gsize)) + (int)crank;
    nb = r / gsize;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total; tile += num_clusters) {
        int z, mb, nb;
        decode(tile, z, mb, nb);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) tc::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
          uint8_t* st = smem + stage * C::STAGE_BYTES;
#pragma unroll
          for (int a = 0; a < S::PA; ++a)
            tc::tma_load_3d_2sm(st + a * C::A_TILE, &tmA, &full_bar[stage], kb * BK_BYTES, (mb + p.mb0) * BM, z * S::PA + a);
#pragma unroll
          for (int b = 0; b < S::PB; ++b)
            tc::tma_load_3d_2sm(st + S::PA * C::A_TILE + b * C::BH_TILE, &tmBh, &full_bar[stage], kb * BK_BYTES,
                                (nb + p.nb0) * S::BN + (int)crank * (S::BN / 2), z * S::PB + b);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      constexpr uint32_t idesc = tc::make_idesc_i8(2 * BM, S::BN, S::SIGNED, S::SIGNED);
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t bphase = 0;
      for (int tile = cluster_id; tile < total; tile += num_clusters) {
        tc::mbar_wait(&tempty_bar[buf], bphase ^ 1);  // both CTAs' epilogues have drained this accumulator buffer
        tc::tc_fence_after();
        const uint32_t acc_base = tmem_base + buf * C::ACC_COLS;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = tc::smem_u32(smem + stage * C::STAGE_BYTES);
            const uint32_t sb = sa + S::PA * C::A_TILE;
#pragma unroll
            for (int kk = 0; kk < BK_BYTES / UMMA_K; ++kk) {
#pragma unroll
              for (int pr = 0; pr < S::NPROD; ++pr) {
                const uint64_t da = tc::make_smem_desc_sw128(sa + S::pa(pr) * C::A_TILE + kk * UMMA_K);
                const uint64_t db = tc::make_smem_desc_sw128(sb + S::pb(pr) * C::BH_TILE + kk * UMMA_K);
                bool first_in_slot = true;
#pragma unroll
                for (int q = 0; q < pr; ++q)
                  if (S::slot(q) == S::slot(pr)) first_in_slot = false;
                const uint32_t accumulate = (kb > 0 || kk > 0 || !first_in_slot) ? 1u : 0u;
                tc::mma_i8_ss_2sm(acc_base + S::slot(pr) * S::BN, da, db, idesc, accumulate);
              }
            }
            tc::mma_commit_2sm(&empty_bar[stage], (uint16_t)0x3);
            if (kb == p.num_kb - 1) tc::mma_commit_2sm(&tfull_bar[buf], (uint16_t)0x3);
          }
          __syncwarp();
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++buf == C::NBUF) {
          buf = 0;
          bphase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 of both CTAs) =====================
    const int q = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    int buf = 0;
    uint32_t bphase = 0;
    for (int tile = cluster_id; tile < total; tile += num_clusters) {
      int z, mb, nb;
      decode(tile, z, mb, nb);
      tc::mbar_wait(&tfull_bar[buf], bphase);
      tc::tc_fence_after();
      const int row = mb * BM + q * 32 + lane;
      const bool row_ok = row < p.m;
      const uint32_t acc_base = tmem_base + buf * C::ACC_COLS + lane_addr;
      const int col0 = nb * S::BN;
#pragma unroll 1
      for (int c0 = 0; c0 < S::BN; c0 += 16) {
        if (col0 + c0 >= p.n) break;
        uint32_t v[S::NSLOT][16];
#pragma unroll
        for (int s = 0; s < S::NSLOT; ++s) tc::tmem_ld16(acc_base + s * S::BN + c0, v[s]);
        tc::tmem_ld_wait();
        if constexpr (S::EPI == EPI_POS) {
          uint32_t old[16];
          uint32_t* dst0 = p.C + (int64_t)(col0 + c0) * p.ldc + row;
          if (p.mode != GFFM_GEMM_STORE) {
#pragma unroll
            for (int c = 0; c < 16; ++c) old[c] = (row_ok && col0 + c0 + c < p.n) ? dst0[(int64_t)c * p.ldc] : 0u;
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            uint64_t acc = (uint64_t)v[0][c];
            if constexpr (S::NSLOT == 3) acc += ((uint64_t)v[1][c] << 8) + ((uint64_t)v[2][c] << 16);
            uint32_t r = (uint32_t)mod_u64(acc, p.modP);
            if (p.mode == GFFM_GEMM_ADD) r = addmod_u32(old[c], r, (uint32_t)p.modP.P);
            else if (p.mode == GFFM_GEMM_SUB) r = submod_u32(old[c], r, (uint32_t)p.modP.P);
            old[c] = r;
          }
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (row_ok && col0 + c0 + c < p.n) dst0[(int64_t)c * p.ldc] = old[c];
        } else {
          const RnsModDev md = p.mods[z];
          uint8_t* eplane = p.E + (int64_t)z * p.e_plane_stride;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int col = col0 + c0 + c;
            const uint32_t u = v[0][c] + md.off;
            uint32_t e = u - __umulhi(u, md.mu) * md.m;
            if (e >= md.m) e -= md.m;
            e *= md.u;
            e -= __umulhi(e, md.mu) * md.m;
            if (e >= md.m) e -= md.m;
            if (row_ok && col < p.n) eplane[(int64_t)col * p.lde + row] = (uint8_t)e;
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) tc::mbar_arrive(&tempty_bar[buf]);
        else tc::mbar_arrive_remote(&tempty_bar[buf], 0);
      }
      if (++buf == C::NBUF) {
        buf = 0;
        bphase ^= 1;
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();
  if (warp == 1) tc::tmem_dealloc_2sm(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// prologue: residues -> 8-bit planes (HBM-bound).  Plane layout: [plane][row][Kp] bytes, K contiguous
// ("K-major"), rows = output rows of the operand (A: i, B: j), zero-filled for k >= K.
// ---------------------------------------------------------------------------------------------------
struct SplitParams {
  int nplanes;          // planes written per call
  int mode;             // 0 = positional limb p -> (x >> 8p) & 255 ; 1 = RNS residues (generic) ; 2 = RNS fast path (R <= 2^27)
  uint32_t R;           // RNS: input bound (for balancing), 0 = unbalanced
  uint32_t half;        // RNS: values > half are shifted by -R
  uint32_t m[MAX_MODS], mu[MAX_MODS];
  uint32_t cneg[MAX_MODS];  // (-R) mod m
  // fast path (fp32 pipe, all values exact integers below 2^24):  v = hi * 2^14 + lo, |hi| <= 2^13, 0 <= lo < 2^14
  //   t = hi * (2^14 mod m) + lo  (== v mod m, |t| < 2^22);  q = rint(t / m);  digit = t - q * m, |digit| <= 127 for m <= 255
  //   (the error of the rounded reciprocal moves the rounding point by at most |t| * 2^-24 < 1/4 of a unit);
  //   m = 256: the digit is simply the low byte of v
  float c14[MAX_MODS];      // 2^14 mod m
  float inv[MAX_MODS];      // 1 / m rounded to nearest
  float mneg[MAX_MODS];     // -m
};

__device__ __forceinline__ uint32_t small_mod(uint32_t a, uint32_t m, uint32_t mu) {
  uint32_t r = a - __umulhi(a, mu) * m;
  if (r >= m) r -= m;
  return r;
}

__device__ __forceinline__ uint32_t put_byte(uint32_t w, uint32_t e, int pos) {
  return __byte_perm(w, e, pos == 0 ? 0x3214 : pos == 1 ? 0x3240 : pos == 2 ? 0x3410 : 0x4210);
}

constexpr float kMagic = 12582912.f;  // 1.5 * 2^23: adding it rounds to an integer and leaves that integer in the low mantissa bits

// Encoder of one thread's NE elements (consecutive k): prepare once, then emit the NE/4 packed words of any plane.
// RNS digits are residues as two's-complement int8 in [-128, 127] (any representative of the class in that range is valid
// for the signed-int8 MMA).
template <int MODE, int NE>
struct Encoder {
  uint32_t xi[NE];              // MODE 0/1: the value; MODE 2: v + 2^27 (v = x, or x - R for the upper half when balanced)
  float hf[MODE == 2 ? NE : 1], lf[MODE == 2 ? NE : 1];
  uint32_t negmask = 0;

  __device__ __forceinline__ void prepare(const uint32_t (&x)[NE], const SplitParams& sp) {
#pragma unroll
    for (int t = 0; t < NE; ++t) {
      if constexpr (MODE == 2) {
        uint32_t v = x[t] + (1u << 27);
        if (sp.R && x[t] > sp.half) v -= sp.R;
        xi[t] = v;
        hf[t] = __uint_as_float(0x4B000000u | (v >> 14)) - 8396800.f;   // (v >> 14) - 2^13: removes the 2^27 bias
        lf[t] = __uint_as_float(0x4B000000u | (v & 16383u)) - 8388608.f;
      } else {
        xi[t] = x[t];
        if (MODE == 1 && sp.R) negmask |= (x[t] > sp.half ? 1u : 0u) << t;
      }
    }
  }

  __device__ __forceinline__ void plane(int pl, const SplitParams& sp, uint32_t (&w)[NE / 4]) const {
    if constexpr (MODE == 0) {
      const uint32_t sh = 8u * pl;
#pragma unroll
      for (int t = 0; t < NE; ++t) w[t >> 2] = put_byte(t & 3 ? w[t >> 2] : 0u, xi[t] >> sh, t & 3);
    } else if constexpr (MODE == 1) {
      const uint32_t m = sp.m[pl], mu = sp.mu[pl], cneg = sp.cneg[pl];
#pragma unroll
      for (int t = 0; t < NE; ++t) {
        uint32_t r = small_mod(xi[t], m, mu);
        if ((negmask >> t) & 1u) {  // x - R
          r += cneg;
          if (r >= m) r -= m;
        }
        w[t >> 2] = put_byte(t & 3 ? w[t >> 2] : 0u, r >= 128u ? r - m : r, t & 3);
      }
    } else {
      if (sp.m[pl] == 256u) {
#pragma unroll
        for (int t = 0; t < NE; ++t) w[t >> 2] = put_byte(t & 3 ? w[t >> 2] : 0u, xi[t], t & 3);
        return;
      }
      const float c14 = sp.c14[pl], inv = sp.inv[pl], mneg = sp.mneg[pl];
#pragma unroll
      for (int t = 0; t < NE; ++t) {
        const float tt = fmaf(hf[t], c14, lf[t]);
        const float q = fmaf(tt, inv, kMagic) - kMagic;
        const float r = fmaf(q, mneg, tt);
        w[t >> 2] = put_byte(t & 3 ? w[t >> 2] : 0u, __float_as_uint(r + kMagic), t & 3);
      }
    }
  }
};

// B operand (k x n column-major, K contiguous already): thread = 16 consecutive k of one column j; a warp reads 2 KiB
// and writes 512 contiguous bytes per plane.
// Optional second source (Karatsuba prologue fusion, reference KaratsubaKernels.jl:129-139): x = src + src2.
// Fused split + push (multi-GPU layer, mg.cu): the same split, but every 16-byte chunk of every plane is stored to `nd` plane
// buffers at once -- the local one and the peers' (NVLink peer memory, posted stores).  The owner of a column range of B thereby
// delivers its operand planes to all ranks in the pass that creates them: no staging copy, no separate collective.
struct PlaneDests {
  uint8_t* p[32];
  int n;
  int sms;  // host side: SMs (= CTAs) the push may occupy
};

template <int MODE>
__global__ void __launch_bounds__(1024)
split_b_push_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ src2, int64_t ld, int64_t ld2, int K, int ncols, int64_t Kp,
                    int64_t rowsP, const __grid_constant__ SplitParams sp, const __grid_constant__ PlaneDests dests) {
  // Grid-stride over a FEW DEDICATED SMs (one 1024-thread CTA each, with a dynamic shared-memory request large enough that no
  // tensor-core GEMM CTA fits beside it).  Measured on 8 B200 (profiles/r02_notes.md): when push CTAs share SMs with the persistent
  // GEMM, the slow peer stores clog those SMs' path to the crossbar and the GEMM's TMA loads queue behind them (a 0.4 ms GEMM
  // launch took 2.5 ms); on its own SMs the push is bound by NVLink, and the GEMM runs undisturbed on the remaining SMs
  // (gffm_ctx::gemm_ctas).  Work item = 16 consecutive k of one column; consecutive threads take consecutive items.
  const int64_t per_col = Kp / 16, items = per_col * ncols;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < items; w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = w / per_col, k16 = (w - j * per_col) * 16;
    uint32_t x[16];
    const uint32_t* col = src + j * ld + k16;
    if (k16 + 15 < K && ((reinterpret_cast<uintptr_t>(col) & 15) == 0)) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 v = reinterpret_cast<const uint4*>(col)[g];
        x[4 * g] = v.x; x[4 * g + 1] = v.y; x[4 * g + 2] = v.z; x[4 * g + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int t = 0; t < 16; ++t) x[t] = (k16 + t < K) ? col[t] : 0u;
    }
    if (src2) {
      const uint32_t* col2 = src2 + j * ld2 + k16;
#pragma unroll
      for (int t = 0; t < 16; ++t)
        if (k16 + t < K) x[t] += col2[t];
    }
    Encoder<MODE, 16> enc;
    enc.prepare(x, sp);
    const int64_t off = j * Kp + k16;
    const int64_t pstride = rowsP * Kp;
#pragma unroll 1
    for (int pl = 0; pl < sp.nplanes; ++pl) {
      uint32_t wv[4];
      enc.plane(pl, sp, wv);
      const uint4 v = make_uint4(wv[0], wv[1], wv[2], wv[3]);
#pragma unroll 1
      for (int d = 0; d < dests.n; ++d) *reinterpret_cast<uint4*>(dests.p[d] + off + pl * pstride) = v;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(128)
split_b_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ src2, int64_t ld, int64_t ld2, int K, int ncols,
               uint8_t* __restrict__ planes, int64_t Kp, int64_t rowsP, const __grid_constant__ SplitParams sp) {
  const int64_t k16 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  const int j = blockIdx.y;
  if (k16 >= Kp || j >= ncols) return;
  uint32_t x[16];
  const uint32_t* col = src + (int64_t)j * ld + k16;
  if (k16 + 15 < K && ((reinterpret_cast<uintptr_t>(col) & 15) == 0)) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint4 v = reinterpret_cast<const uint4*>(col)[g];
      x[4 * g] = v.x; x[4 * g + 1] = v.y; x[4 * g + 2] = v.z; x[4 * g + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int t = 0; t < 16; ++t) x[t] = (k16 + t < K) ? col[t] : 0u;
  }
  if (src2) {
    const uint32_t* col2 = src2 + (int64_t)j * ld2 + k16;
#pragma unroll
    for (int t = 0; t < 16; ++t)
      if (k16 + t < K) x[t] += col2[t];
  }
  Encoder<MODE, 16> enc;
  enc.prepare(x, sp);
  uint8_t* out = planes + (int64_t)j * Kp + k16;
  const int64_t pstride = rowsP * Kp;
#pragma unroll 1
  for (int pl = 0; pl < sp.nplanes; ++pl) {
    uint32_t w[4];
    enc.plane(pl, sp, w);
    *reinterpret_cast<uint4*>(out + pl * pstride) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// A operand (m x k column-major, M contiguous): block tile = 32 rows x 256 k, lane = row i.  A thread's 32 loads (two
// passes of 16 consecutive k) are each coalesced across the warp (32 rows x 4 B) and all in flight together.  The digits
// of 16 consecutive k are packed in registers into one 16-byte chunk per plane; chunks are exchanged through a swizzled
// 4 KiB shared tile (conflict-free 128-bit stores and loads, double-buffered, one barrier per plane) so that every warp
// store instruction writes four full 128-byte lines of the K-major planes.
constexpr int SPLIT_A_KT = 256;

template <int MODE>
__global__ void __launch_bounds__(256)
split_a_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ src2, int64_t ld, int64_t ld2, int M, int K,
               uint8_t* __restrict__ planes, int64_t Kp, int64_t rowsP, const __grid_constant__ SplitParams sp) {
  __shared__ uint4 tile[2][32 * 8];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  const int i0 = blockIdx.x * 32;
  const int i = i0 + lane;
  const int64_t k0 = (int64_t)blockIdx.y * SPLIT_A_KT;
  uint32_t x[2][16];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t kb = k0 + 128 * h + 16 * w;
    const uint32_t* s = src + kb * ld + i;
    if (i < M && kb + 16 <= K) {
#pragma unroll
      for (int t = 0; t < 16; ++t) x[h][t] = s[(int64_t)t * ld];
    } else {
#pragma unroll
      for (int t = 0; t < 16; ++t) x[h][t] = (i < M && kb + t < K) ? s[(int64_t)t * ld] : 0u;
    }
    if (src2) {
      const uint32_t* s2 = src2 + kb * ld2 + i;
#pragma unroll
      for (int t = 0; t < 16; ++t)
        if (i < M && kb + t < K) x[h][t] += s2[(int64_t)t * ld2];
    }
  }
  const int64_t pstride = rowsP * Kp;
  const int r = threadIdx.x >> 3, j = threadIdx.x & 7;  // write-out role: 16-byte chunk j of tile row r
  const int st_idx = lane * 8 + (w ^ (lane & 7));
  const int ld_idx = r * 8 + (j ^ (r & 7));
  int buf = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (k0 + 128 * h >= Kp) break;  // block-uniform
    Encoder<MODE, 16> enc;
    enc.prepare(x[h], sp);
    uint8_t* out = planes + (int64_t)(i0 + r) * Kp + k0 + 128 * h + 16 * j;
#pragma unroll 1
    for (int pl = 0; pl < sp.nplanes; ++pl) {
      uint32_t wd[4];
      enc.plane(pl, sp, wd);
      tile[buf][st_idx] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
      __syncthreads();
      if (i0 + r < M) *reinterpret_cast<uint4*>(out + pl * pstride) = tile[buf][ld_idx];
      buf ^= 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// CRT epilogue kernel (RNS): planes e_t = (x * (M/m_t)^-1) mod m_t  ->  x mod P, fused C +/-= and the Karatsuba
// carry split (reference KaratsubaKernels.jl:141-158: C1 = cc mod N1, carry = cc div N1).
//   x = S - round(S/M) * M,  S = sum_t e_t * (M/m_t);  frac(S/M) = sum_t e_t / m_t tracked in 2^-32 fixed point
//   (the moduli are chosen with |x|/M < 1/2 - 2^-7, so 20 bits of fraction are ample).
// HBM-bound by design: each thread handles 4 consecutive rows (one 32-bit load per plane, all planes in flight, one
// 128-bit store); per (element, modulus) the work is one byte extract and two 32x32+64 multiply-adds.
// ---------------------------------------------------------------------------------------------------
struct CrtParams {
  int s;
  int mode;
  int balanced;
  ModP modP;
  uint64_t w[MAX_MODS];   // (M/m_t) mod P
  uint32_t f[MAX_MODS];   // floor(2^32 / m_t)  (m_t >= 2)
  uint64_t Wq[MAX_MODS + 2];  // (-q * M) mod P for q = 0..s
  uint64_t kara_N1;       // != 0: C <- (x mod P) mod N1, hi <- (x mod P) div N1
  // byte-sliced constants of crt_fast_kernel (P < 2^32, no carry split): group g = planes 4g..4g+3, one byte lane per plane
  int fast;               // 1: crt_fast_kernel applies; 2: and 2^16 < P < 2^30 (single-multiply reduction)
  uint32_t wb[4][4];      // byte j of w_t, packed over the 4 planes of group g: wb[g][j]
  uint32_t fb[4][2];      // byte j of floor(2^23 / m_t)
  uint32_t Wq32[MAX_MODS + 2];
  uint32_t mu48;          // floor(2^48 / P)
};

template <bool WIDE>  // WIDE: P >= 2^32 (Karatsuba P1), 64-bit weights
__global__ void __launch_bounds__(256)
crt_kernel(const uint8_t* __restrict__ E, int64_t lde, int64_t plane_stride, int m, int n, uint32_t* __restrict__ C,
           int64_t ldc, uint32_t* __restrict__ hi, int64_t ldhi, const __grid_constant__ CrtParams cp) {
  const int i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int j = blockIdx.y;
  if (i4 >= m || j >= n) return;
  const uint8_t* e = E + (int64_t)j * lde + i4;  // lde is a multiple of 128 and i4 of 4: aligned 32-bit loads
  uint32_t ew[MAX_MODS];
#pragma unroll
  for (int t = 0; t < MAX_MODS; ++t) ew[t] = (t < cp.s) ? *reinterpret_cast<const uint32_t*>(e + (int64_t)t * plane_stride) : 0u;
  uint64_t acc[4] = {0, 0, 0, 0}, F[4] = {0, 0, 0, 0};
#pragma unroll
  for (int t = 0; t < MAX_MODS; ++t) {
    if (t < cp.s) {
      const uint32_t ft = cp.f[t];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t et = __byte_perm(ew[t], 0, 0x4440 + q);
        if constexpr (WIDE) acc[q] += (uint64_t)et * cp.w[t];
        else acc[q] += (uint64_t)et * (uint32_t)cp.w[t];
        F[q] += (uint64_t)et * ft;
      }
    }
  }
  uint32_t r32[4];
  uint32_t* dst = C + (int64_t)j * ldc + i4;
  const int nv = min(4, m - i4);
  const uint64_t rnd = cp.balanced ? ((1ull << 31) + (1ull << 12)) : (1ull << 12);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t qq = (uint32_t)((F[q] + rnd) >> 32);
    uint64_t r;
    if constexpr (WIDE) {
      const uint64_t t1 = mod_u64(acc[q], cp.modP);
      const uint64_t t2 = cp.Wq[qq];
      r = t1 + t2;
      if (r >= cp.modP.P) r -= cp.modP.P;
    } else {
      r = mod_u64(acc[q] + cp.Wq[qq], cp.modP);
    }
    if (cp.kara_N1) {
      if (q < nv) hi[(int64_t)j * ldhi + i4 + q] = (uint32_t)(r / cp.kara_N1);
      r32[q] = (uint32_t)(r % cp.kara_N1);
    } else {
      r32[q] = (uint32_t)r;
    }
  }
  const bool vec = nv == 4 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  if (!cp.kara_N1 && cp.mode != GFFM_GEMM_STORE) {
    uint32_t old[4];
    if (vec) {
      const uint4 o = *reinterpret_cast<const uint4*>(dst);
      old[0] = o.x; old[1] = o.y; old[2] = o.z; old[3] = o.w;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) old[q] = q < nv ? dst[q] : 0u;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      r32[q] = cp.mode == GFFM_GEMM_ADD ? addmod_u32(old[q], r32[q], (uint32_t)cp.modP.P) : submod_u32(old[q], r32[q], (uint32_t)cp.modP.P);
  }
  if (vec) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(r32[0], r32[1], r32[2], r32[3]);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q < nv) dst[q] = r32[q];
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// planes: [nplanes][rowsP][Kp] bytes; box = {128 bytes of K, box_rows, 1 plane}, 128-byte swizzle
int32_t make_plane_tmap(CUtensorMap* tm, void* planes, int64_t Kp, int64_t rowsP, int64_t nplanes, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) GFFM_FAIL(GFFM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rowsP, (cuuint64_t)nplanes};
  cuuint64_t strides[2] = {(cuuint64_t)Kp, (cuuint64_t)(Kp * rowsP)};
  cuuint32_t box[3] = {(cuuint32_t)BK_BYTES, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, planes, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) GFFM_FAIL(GFFM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return GFFM_OK;
}

int32_t make_plane_tmap(CUtensorMap* tm, void* planes, int64_t Kp, int64_t rowsP, int64_t nplanes, int box_rows);

// GFFM_MCAST: 0 = single-CTA kernel, 1 = cluster-of-2 kernel with TMA-multicast B halves, 2 = CTA-pair kernel (cta_group::2)
inline int mcast_mode() {
  static const int m = getenv("GFFM_MCAST") ? atoi(getenv("GFFM_MCAST")) : 0;
  return m;
}

template <class S>
int32_t launch_gemm(gffm_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t st = nullptr,
                    const CUtensorMap* tmB_half = nullptr) {
  using C = Cfg<S>;
  if (tmB_half && mcast_mode() == 2 && p.num_m_blk >= 2) {
    using C2 = Cfg2<S>;
    static PerDeviceOnce attr_2;
    attr_2.run(ctx->device, [] { cudaFuncSetAttribute(gemm_tc_kernel_2cta<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2::SMEM_BYTES); });
    const int num_mp = (p.num_m_blk + 1) / 2;
    const int total_c = p.batches * num_mp * p.num_n_blk;
    int clusters = ctx->num_sms / 2;
    if (total_c < clusters) clusters = total_c;
    GemmParams q = p;
    q.l2_hints = 0;
    q.fake_loads = 0;
    gemm_tc_kernel_2cta<S><<<clusters * 2, 192, C2::SMEM_BYTES, st ? st : ctx->stream>>>(tmA, *tmB_half, q);
    GFFM_LAUNCH_CHECK(ctx);
    return GFFM_OK;
  }
  if (tmB_half && mcast_mode() == 1 && p.num_m_blk >= 2) {
    static PerDeviceOnce attr_mc;
    attr_mc.run(ctx->device, [] { cudaFuncSetAttribute(gemm_tc_kernel_mc<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES); });
    const int num_mp = (p.num_m_blk + 1) / 2;
    const int total_c = p.batches * num_mp * p.num_n_blk;
    int clusters = ctx->num_sms / 2;
    if (total_c < clusters) clusters = total_c;
    static const int hints2 = getenv("GFFM_L2_HINTS") ? atoi(getenv("GFFM_L2_HINTS")) : 0;
    GemmParams q = p;
    q.l2_hints = hints2;
    q.fake_loads = 0;
    gemm_tc_kernel_mc<S><<<clusters * 2, 192, C::SMEM_BYTES, st ? st : ctx->stream>>>(tmA, *tmB_half, q);
    GFFM_LAUNCH_CHECK(ctx);
    return GFFM_OK;
  }
  static PerDeviceOnce attr_set;  // a failed attribute call surfaces as a launch error below (GFFM_LAUNCH_CHECK)
  attr_set.run(ctx->device, [] { cudaFuncSetAttribute(gemm_tc_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES); });
  const int total = p.batches * p.num_m_blk * p.num_n_blk;
  static const int env_ctas = getenv("GFFM_GEMM_CTAS") ? atoi(getenv("GFFM_GEMM_CTAS")) : 0;
  const int max_ctas = ctx->gemm_ctas > 0 ? ctx->gemm_ctas : env_ctas;  // leave SMs to concurrent kernels (NCCL in the multi-GPU layer)
  const int sms = (max_ctas > 0 && max_ctas < ctx->num_sms) ? max_ctas : ctx->num_sms;
  const int grid = total < sms ? total : sms;
  static const int hints = getenv("GFFM_L2_HINTS") ? atoi(getenv("GFFM_L2_HINTS")) : 0;
  static const int fake = getenv("GFFM_FAKE_LOADS") ? atoi(getenv("GFFM_FAKE_LOADS")) : 0;
  GemmParams q = p;
  q.l2_hints = hints;
  q.fake_loads = fake;
  gemm_tc_kernel<S><<<grid, 192, C::SMEM_BYTES, st ? st : ctx->stream>>>(tmA, tmB, q);
  GFFM_LAUNCH_CHECK(ctx);
  return GFFM_OK;
}

int32_t run_split(gffm_ctx* ctx, bool is_a, const MatView& X, const MatView* X2, int64_t k_off, int64_t kc, uint8_t* planes,
                  int64_t Kp, int64_t rowsP, const SplitParams& sp, cudaStream_t st = nullptr, const PlaneDests* dests = nullptr) {
  if (!st) st = ctx->stream;
  if (dests && !is_a) {  // fused split + push: `planes` is ignored, every destination gets the rows the caller offset into dests
    const int64_t items = (Kp / 16) * X.cols;
    const int push_sms = dests->sms > 0 ? dests->sms : 20;
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(ceil_div(items, 1024), push_sms));
    constexpr int kPushSmem = 100 * 1024;  // + the GEMM CTA's ~197 KiB > 227 KiB: never co-resident with a GEMM CTA
    static PerDeviceOnce attr_push;
    attr_push.run(ctx->device, [] {
      cudaFuncSetAttribute(split_b_push_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPushSmem);
      cudaFuncSetAttribute(split_b_push_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPushSmem);
      cudaFuncSetAttribute(split_b_push_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPushSmem);
    });
    const uint32_t* s = X.p + k_off;
    const uint32_t* s2 = X2 ? X2->p + k_off : nullptr;
#define GFFM_SPLIT_BP(MODE) split_b_push_kernel<MODE><<<(unsigned)grid, 1024, kPushSmem, st>>>(s, s2, X.ld, X2 ? X2->ld : 0, (int)kc, (int)X.cols, Kp, rowsP, sp, *dests)
    if (sp.mode == 0) GFFM_SPLIT_BP(0);
    else if (sp.mode == 1) GFFM_SPLIT_BP(1);
    else GFFM_SPLIT_BP(2);
#undef GFFM_SPLIT_BP
    GFFM_LAUNCH_CHECK(ctx);
    return GFFM_OK;
  }
  if (is_a) {
    // A is m x K: rows = i, K along columns of the view
    const uint32_t* s = X.p + k_off * X.ld;
    const uint32_t* s2 = X2 ? X2->p + k_off * X2->ld : nullptr;
    dim3 grid((unsigned)ceil_div(X.rows, 32), (unsigned)ceil_div(Kp, SPLIT_A_KT));
#define GFFM_SPLIT_A(MODE) split_a_kernel<MODE><<<grid, 256, 0, st>>>(s, s2, X.ld, X2 ? X2->ld : 0, (int)X.rows, (int)kc, planes, Kp, rowsP, sp)
    if (sp.mode == 0) GFFM_SPLIT_A(0);
    else if (sp.mode == 1) GFFM_SPLIT_A(1);
    else GFFM_SPLIT_A(2);
#undef GFFM_SPLIT_A
  } else {
    // B is K x n: K along rows of the view
    // gridDim.y carries the column: slabs of at most 65535 columns per launch
    for (int64_t c0 = 0; c0 < X.cols; c0 += 65535) {
      const int64_t nc = std::min<int64_t>(65535, X.cols - c0);
      const uint32_t* s = X.p + k_off + c0 * X.ld;
      const uint32_t* s2 = X2 ? X2->p + k_off + c0 * X2->ld : nullptr;
      uint8_t* pl = planes + c0 * Kp;
      dim3 grid((unsigned)ceil_div(Kp / 16, 128), (unsigned)nc);
#define GFFM_SPLIT_B(MODE) split_b_kernel<MODE><<<grid, 128, 0, st>>>(s, s2, X.ld, X2 ? X2->ld : 0, (int)kc, (int)nc, pl, Kp, rowsP, sp)
      if (sp.mode == 0) GFFM_SPLIT_B(0);
      else if (sp.mode == 1) GFFM_SPLIT_B(1);
      else GFFM_SPLIT_B(2);
#undef GFFM_SPLIT_B
      if (c0 + nc < X.cols) GFFM_LAUNCH_CHECK(ctx);
    }
  }
  GFFM_LAUNCH_CHECK(ctx);
  return GFFM_OK;
}

// Returns the plane buffer for operand X (role 0 = A, 1 = B) and runs the split unless a valid cached copy exists.
int32_t acquire_planes(gffm_ctx* ctx, int role, const MatView& X, const MatView* X2, int64_t k0, int64_t kc, int64_t Kp, int64_t rowsP,
                       const SplitParams& sp, int balanced, uint64_t R, gffm_workspace* ws, uint8_t** out) {
  const size_t bytes = (size_t)sp.nplanes * rowsP * Kp;
  gffm_mat* own = X2 ? nullptr : X.owner;
  if (own) {
    gffm_plane_cache& c = own->cache[role];
    const bool hit = c.valid && c.version == own->version && c.mode == sp.mode && c.nplanes == sp.nplanes && c.balanced == balanced &&
                     c.R == R && c.r0 == X.r0 && c.c0 == X.c0 && c.rows == X.rows && c.cols == X.cols && c.k0 == k0 && c.kc == kc &&
                     c.Kp == Kp && c.rowsP == rowsP;
    if (hit) {
      *out = (uint8_t*)c.ptr;
      return GFFM_OK;
    }
    if (c.bytes < bytes) {
      if (c.ptr) {
        gffm_dev_free(ctx, c.ptr);
        c.ptr = nullptr;
        c.bytes = 0;
      }
      if (gffm_dev_alloc(ctx, &c.ptr, bytes) != cudaSuccess) {  // no memory for a cache: fall back to the shared workspace
        c.ptr = nullptr;
        own = nullptr;
      } else {
        c.bytes = bytes;
      }
    }
    if (own) {
      c.valid = false;
      GFFM_TRY(run_split(ctx, role == 0, X, X2, k0, kc, (uint8_t*)c.ptr, Kp, rowsP, sp));
      c.valid = true;
      c.version = own->version;
      c.mode = sp.mode; c.nplanes = sp.nplanes; c.balanced = balanced; c.R = R;
      c.r0 = X.r0; c.c0 = X.c0; c.rows = X.rows; c.cols = X.cols; c.k0 = k0; c.kc = kc; c.Kp = Kp; c.rowsP = rowsP;
      *out = (uint8_t*)c.ptr;
      return GFFM_OK;
    }
  }
  GFFM_TRY(gffm_ws_reserve(ctx, ws, bytes));
  GFFM_TRY(run_split(ctx, role == 0, X, X2, k0, kc, (uint8_t*)ws->ptr, Kp, rowsP, sp));
  *out = (uint8_t*)ws->ptr;
  return GFFM_OK;
}

inline void prof_mark(gffm_ctx* ctx, int idx) {
  if (ctx->profile && idx < 8) {
    cudaEventRecord(ctx->ev[idx], ctx->stream);
    ctx->n_ev = idx + 1;
  }
}

// CRT for P < 2^32 on the dot-product unit.  A thread owns 4 consecutive rows; the 4 x 4 byte block (4 planes x 4 rows)
// of each plane group is transposed in registers (8 PRMT), then per row and group 4 dp4a accumulate the byte slices of
// S = sum_t e_t * w_t and 2 dp4a the slices of the quotient estimate F = sum_t e_t * floor(2^23 / m_t)  (error < s * 2^-15,
// the plan keeps |x| / M at least 2^-7 away from the rounding boundary).  ~40 instructions per element instead of ~170.
template <bool FASTRED>
__global__ void __launch_bounds__(256)
crt_fast_kernel(const uint8_t* __restrict__ E, int64_t lde, int64_t plane_stride, int m, int n, uint32_t* __restrict__ C, int64_t ldc,
                const __grid_constant__ CrtParams cp) {
  const int i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int j = blockIdx.y;
  if (i4 >= m || j >= n) return;
  const uint8_t* e = E + (int64_t)j * lde + i4;  // lde is a multiple of 128 and i4 of 4: aligned 32-bit loads
  const int ngroups = (cp.s + 3) >> 2;
  // planes past s re-read plane s-1 (their weight bytes are zero): no per-plane predicates or address multiplies
  uint32_t ew[16];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (g < ngroups) {
#pragma unroll
      for (int t = 4 * g; t < 4 * g + 4; ++t) {
        ew[t] = *reinterpret_cast<const uint32_t*>(e);
        if (t + 1 < cp.s) e += plane_stride;
      }
    } else {
      ew[4 * g] = ew[4 * g + 1] = ew[4 * g + 2] = ew[4 * g + 3] = 0u;
    }
  }
  uint32_t a[4][4], F[4][2];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    a[q][0] = a[q][1] = a[q][2] = a[q][3] = 0;
    F[q][0] = F[q][1] = 0;
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (g < ngroups) {
      const uint32_t t0 = __byte_perm(ew[4 * g], ew[4 * g + 1], 0x5140), t1 = __byte_perm(ew[4 * g], ew[4 * g + 1], 0x7362);
      const uint32_t u0 = __byte_perm(ew[4 * g + 2], ew[4 * g + 3], 0x5140), u1 = __byte_perm(ew[4 * g + 2], ew[4 * g + 3], 0x7362);
      uint32_t v[4];
      v[0] = __byte_perm(t0, u0, 0x5410);
      v[1] = __byte_perm(t0, u0, 0x7632);
      v[2] = __byte_perm(t1, u1, 0x5410);
      v[3] = __byte_perm(t1, u1, 0x7632);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int b = 0; b < 4; ++b) a[q][b] = __dp4a(v[q], cp.wb[g][b], a[q][b]);
        F[q][0] = __dp4a(v[q], cp.fb[g][0], F[q][0]);
        F[q][1] = __dp4a(v[q], cp.fb[g][1], F[q][1]);
      }
    }
  }
  uint32_t r32[4];
  const uint32_t rnd = cp.balanced ? ((1u << 22) + (1u << 11)) : (1u << 12);
  const uint32_t P = (uint32_t)cp.modP.P;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t qq = (F[q][0] + (F[q][1] << 8) + rnd) >> 23;
    const uint32_t lo = a[q][0] + (a[q][1] << 8);   // < 2^29
    const uint32_t hi = a[q][2] + (a[q][3] << 8);   // < 2^29
    const uint64_t x = (uint64_t)hi * 65536u + lo + cp.Wq32[qq];  // < 2^46
    if constexpr (FASTRED) {
      const uint32_t xh = (uint32_t)(x >> 16);
      uint32_t r = (uint32_t)x - __umulhi(xh, cp.mu48) * P;  // in [0, 3P), 3P < 2^32
      if (r >= P) r -= P;
      if (r >= P) r -= P;
      r32[q] = r;
    } else {
      r32[q] = (uint32_t)mod_u64(x, cp.modP);
    }
  }
  uint32_t* dst = C + (int64_t)j * ldc + i4;
  const int nv = min(4, m - i4);
  const bool vec = nv == 4 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  if (cp.mode != GFFM_GEMM_STORE) {
    uint32_t old[4];
    if (vec) {
      const uint4 o = *reinterpret_cast<const uint4*>(dst);
      old[0] = o.x; old[1] = o.y; old[2] = o.z; old[3] = o.w;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) old[q] = q < nv ? dst[q] : 0u;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) r32[q] = cp.mode == GFFM_GEMM_ADD ? addmod_u32(old[q], r32[q], P) : submod_u32(old[q], r32[q], P);
  }
  if (vec) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(r32[0], r32[1], r32[2], r32[3]);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q < nv) dst[q] = r32[q];
  }
}

const uint32_t kModuli[MAX_MODS] = {256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197};

// All per-modulus constants of one RNS product with inner dimension kc (<= 65536), inputs < R, result mod P.
struct RnsPlan {
  int s = 0;
  SplitParams sp;
  CrtParams cp;
  RnsModDev mods[MAX_MODS];
};

int32_t make_rns_plan(int64_t kc, uint64_t R, uint64_t P, bool balanced, int crt_mode, uint64_t kara_N1, RnsPlan* plan) {
  // number of moduli: M > 2*X (balanced, |x| <= X = kc*floor(R/2)^2) or M > X (X = kc*(R-1)^2), with margin
  unsigned __int128 X = balanced ? (unsigned __int128)kc * (R / 2) * (R / 2) * 2 : (unsigned __int128)kc * (R - 1) * (R - 1);
  X += (X >> 6) + 2;  // |x|/M < 1/2 - 2^-7: the CRT rounding decision has a wide margin
  unsigned __int128 Mprod = 1;
  int s = 0;
  while (Mprod <= X) {
    if (s >= MAX_MODS) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "dynamic range exceeds %d moduli", MAX_MODS);
    Mprod *= kModuli[s++];
  }
  plan->s = s;
  SplitParams& spa = plan->sp;
  memset(&spa, 0, sizeof(spa));
  spa.nplanes = s;
  spa.mode = 1;
  spa.R = balanced ? (uint32_t)R : 0;
  spa.half = (uint32_t)(R / 2);
  if (R <= (1ull << 27)) spa.mode = 2;  // fast residue path: |v| <= 2^27, y < 2^24
  CrtParams& cp = plan->cp;
  memset(&cp, 0, sizeof(cp));
  cp.s = s;
  cp.mode = crt_mode;
  cp.balanced = balanced ? 1 : 0;
  cp.modP = make_modp(P);
  cp.kara_N1 = kara_N1;
  {
    const uint64_t Wm = (uint64_t)(Mprod % P);
    for (int q = 0; q <= s + 1 && q < MAX_MODS + 2; ++q) cp.Wq[q] = (uint64_t)((P - (uint64_t)(((unsigned __int128)q * Wm) % P)) % P);
  }
  cp.fast = (P < (1ull << 32) && !kara_N1) ? ((P > (1ull << 16) && P < (1ull << 30)) ? 2 : 1) : 0;
  if (cp.fast) {
    for (int q = 0; q < MAX_MODS + 2; ++q) cp.Wq32[q] = (uint32_t)cp.Wq[q];
    cp.mu48 = cp.fast == 2 ? (uint32_t)((1ull << 48) / P) : 0u;
  }
  memset(plan->mods, 0, sizeof(plan->mods));
  for (int t = 0; t < s; ++t) {
    const uint32_t mt = kModuli[t];
    uint64_t others_mod_mt = 1;
    unsigned __int128 others_mod_P = 1;
    for (int u = 0; u < s; ++u) {
      if (u == t) continue;
      others_mod_mt = (others_mod_mt * (kModuli[u] % mt)) % mt;
      others_mod_P = (others_mod_P * kModuli[u]) % P;
    }
    const uint64_t ut = modinv_u64(others_mod_mt, mt);
    spa.m[t] = mt;
    spa.mu[t] = (uint32_t)((1ull << 32) / mt);
    spa.cneg[t] = (uint32_t)((mt - (R % mt)) % mt);
    spa.c14[t] = (float)(16384u % mt);
    spa.inv[t] = (float)(1.0 / (double)mt);
    spa.mneg[t] = -(float)mt;
    cp.w[t] = (uint64_t)others_mod_P;
    cp.f[t] = (uint32_t)((1ull << 32) / mt);
    if (cp.fast) {
      const uint32_t f23 = (1u << 23) / mt;  // < 2^16 because every modulus exceeds 128
      for (int b = 0; b < 4; ++b) cp.wb[t >> 2][b] |= (uint32_t)(((uint64_t)others_mod_P >> (8 * b)) & 255u) << (8 * (t & 3));
      for (int b = 0; b < 2; ++b) cp.fb[t >> 2][b] |= ((f23 >> (8 * b)) & 255u) << (8 * (t & 3));
    }
    plan->mods[t].m = mt;
    plan->mods[t].mu = (uint32_t)((1ull << 32) / mt);
    plan->mods[t].off = (uint32_t)(((1ull << 31) + mt - 1) / mt * mt);
    plan->mods[t].u = (uint32_t)ut;
  }
  return GFFM_OK;
}

int32_t launch_crt(gffm_ctx* ctx, cudaStream_t st, const CrtParams& cp, const uint8_t* E, int64_t lde, int64_t e_plane, int64_t m, int64_t n,
                   uint32_t* C, int64_t ldc, uint32_t* hi, int64_t ldhi) {
  // gridDim.y carries the column: slabs of at most 65535 columns per launch
  for (int64_t c0 = 0; c0 < n; c0 += 65535) {
    const int nc = (int)std::min<int64_t>(65535, n - c0);
    dim3 grid((unsigned)ceil_div(m, 1024), (unsigned)nc);
    const uint8_t* Es = E + c0 * lde;
    uint32_t* Cs = C + c0 * ldc;
    uint32_t* his = hi ? hi + c0 * ldhi : nullptr;
    if (cp.fast == 2) crt_fast_kernel<true><<<grid, 256, 0, st>>>(Es, lde, e_plane, (int)m, nc, Cs, ldc, cp);
    else if (cp.fast == 1) crt_fast_kernel<false><<<grid, 256, 0, st>>>(Es, lde, e_plane, (int)m, nc, Cs, ldc, cp);
    else if (cp.modP.P >= (1ull << 32)) crt_kernel<true><<<grid, 256, 0, st>>>(Es, lde, e_plane, (int)m, nc, Cs, ldc, his, ldhi, cp);
    else crt_kernel<false><<<grid, 256, 0, st>>>(Es, lde, e_plane, (int)m, nc, Cs, ldc, his, ldhi, cp);
    GFFM_LAUNCH_CHECK(ctx);
  }
  return GFFM_OK;
}

// largest K of one limb launch such that no int32 accumulator can exceed 2^31 (exactness budget)
int64_t limb_kmax(uint64_t R) {
  const int L = R <= 256 ? 1 : 2;
  const uint64_t top = (R - 1) >> (8 * (L - 1));
  const uint64_t lo = L == 1 ? 0 : 255;
  uint64_t worst = L == 1 ? top * top : (2 * lo * top > lo * lo ? 2 * lo * top : lo * lo);
  if (L == 2 && top * top > worst) worst = top * top;
  if (worst == 0) worst = 1;
  int64_t kmax = (int64_t)(((1ull << 31) - 1) / worst);
  kmax = kmax / 128 * 128;
  if (kmax > (1 << 20)) kmax = 1 << 20;
  return kmax;
}

}  // namespace

bool gffm_tc_available(gffm_ctx* ctx) {
  (void)ctx;
  return get_encode_fn() != nullptr;
}

// Optional second addends (A2,B2) implement the Karatsuba prologue fusion: the planes are those of A+A2 / B+B2.
int32_t gffm_gemm_tc_limb_ex(gffm_ctx* ctx, MatView Cv, MatView A, const MatView* A2, MatView B, const MatView* B2,
                             uint64_t R, uint64_t P, int mode) {
  const int64_t m = A.rows, K = A.cols, n = B.cols;
  if (m == 0 || n == 0) return GFFM_OK;
  if (R > 65536 || P >= (1ull << 32) || P == 0) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "limb GEMM needs R <= 2^16 and P < 2^32");
  const int L = R <= 256 ? 1 : 2;
  const int BN = L == 1 ? SchemeL1::BN : SchemeL2::BN;
  const int64_t kmax = limb_kmax(R);
  if (kmax < 128) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "accumulator budget");
  const int64_t rowsPA = round_up(m, BM), rowsPB = round_up(n, BN);
  const int64_t kchunk_max = K < kmax ? round_up(K > 0 ? K : 1, 128) : kmax;
  (void)kchunk_max;
  SplitParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.nplanes = L;
  sp.mode = 0;
  if (K == 0) {
    if (mode == GFFM_GEMM_STORE) return gffm_fill_view(ctx, Cv, 0);
    return GFFM_OK;
  }
  for (int64_t k0 = 0; k0 < K; k0 += kmax) {
    const int64_t kc = (K - k0) < kmax ? (K - k0) : kmax;
    const int64_t Kp = round_up(kc, 128);
    uint8_t *pa = nullptr, *pb = nullptr;
    prof_mark(ctx, 0);
    GFFM_TRY(acquire_planes(ctx, 0, A, A2, k0, kc, Kp, rowsPA, sp, 0, R, &ctx->ws_planes_a, &pa));
    GFFM_TRY(acquire_planes(ctx, 1, B, B2, k0, kc, Kp, rowsPB, sp, 0, R, &ctx->ws_planes_b, &pb));
    prof_mark(ctx, 1);
    CUtensorMap tmA, tmB;
    CUtensorMap tmBh;
    GFFM_TRY(make_plane_tmap(&tmA, pa, Kp, rowsPA, L, BM));
    GFFM_TRY(make_plane_tmap(&tmB, pb, Kp, rowsPB, L, BN));
    GFFM_TRY(make_plane_tmap(&tmBh, pb, Kp, rowsPB, L, BN / 2));
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.m = (int)m;
    p.n = (int)n;
    p.num_kb = (int)(Kp / 128);
    p.num_m_blk = (int)(rowsPA / BM);
    p.num_n_blk = (int)(rowsPB / BN);
    p.batches = 1;
    p.C = Cv.p;
    p.ldc = Cv.ld;
    p.mode = (k0 == 0) ? mode : (mode == GFFM_GEMM_SUB ? GFFM_GEMM_SUB : GFFM_GEMM_ADD);
    p.modP = make_modp(P);
    if (L == 1) GFFM_TRY(launch_gemm<SchemeL1>(ctx, tmA, tmB, p, nullptr, &tmBh));
    else GFFM_TRY(launch_gemm<SchemeL2>(ctx, tmA, tmB, p, nullptr, &tmBh));
    prof_mark(ctx, 2);
  }
  return GFFM_OK;
}

int32_t gffm_gemm_tc_limb(gffm_ctx* ctx, MatView C, MatView A, MatView B, uint64_t R, uint64_t P, int mode) {
  return gffm_gemm_tc_limb_ex(ctx, C, A, nullptr, B, nullptr, R, P, mode);
}

// RNS path.  balanced: inputs are residues mod R and the result is only needed mod P with P | R (or P == R), so
// inputs may be shifted into (-R/2, R/2] -- halves the dynamic range twice.  kara_hi != nullptr selects the
// Karatsuba carry split with N1 = kara_N1 (P = N1*N2 <= 2^52).
int32_t gffm_gemm_tc_rns_ex(gffm_ctx* ctx, MatView Cv, MatView A, const MatView* A2, MatView B, const MatView* B2, uint64_t R,
                            uint64_t P, int mode, bool balanced, uint32_t* kara_hi, int64_t ldhi, uint64_t kara_N1) {
  const int64_t m = A.rows, K = A.cols, n = B.cols;
  if (m == 0 || n == 0) return GFFM_OK;
  if (R == 0 || R >= (1ull << 32) || P == 0 || P > (1ull << 52)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "RNS GEMM needs R < 2^32, P <= 2^52");
  if (!kara_hi && P >= (1ull << 32)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "uint32 output needs P < 2^32");
  if (K == 0) {
    if (mode == GFFM_GEMM_STORE && !kara_hi) return gffm_fill_view(ctx, Cv, 0);
    return GFFM_OK;
  }
  const int64_t kmax = 65536;  // |acc| <= K * 128 * 128 < 2^31
  const int BN = SchemeRNS::BN;
  const int64_t rowsPA = round_up(m, BM), rowsPB = round_up(n, BN);
  for (int64_t k0 = 0; k0 < K; k0 += kmax) {
    const int64_t kc = (K - k0) < kmax ? (K - k0) : kmax;
    const int64_t Kp = round_up(kc, 128);
    RnsPlan plan;
    GFFM_TRY(make_rns_plan(kc, R, P, balanced, (k0 == 0) ? mode : (mode == GFFM_GEMM_SUB ? GFFM_GEMM_SUB : GFFM_GEMM_ADD), kara_hi ? kara_N1 : 0,
                           &plan));
    const int s = plan.s;
    const SplitParams& spa = plan.sp;
    const CrtParams& cp = plan.cp;
    const int64_t lde = round_up(m, 128);
    const int64_t e_plane = lde * n;
    GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_eplanes, (size_t)s * e_plane));
    GemmParams p;
    memset(&p, 0, sizeof(p));
    memcpy(p.mods, plan.mods, sizeof(p.mods));
    const SplitParams& spb = spa;
    uint8_t *pa = nullptr, *pb = nullptr;
    prof_mark(ctx, 0);
    GFFM_TRY(acquire_planes(ctx, 0, A, A2, k0, kc, Kp, rowsPA, spa, balanced ? 1 : 0, R, &ctx->ws_planes_a, &pa));
    GFFM_TRY(acquire_planes(ctx, 1, B, B2, k0, kc, Kp, rowsPB, spb, balanced ? 1 : 0, R, &ctx->ws_planes_b, &pb));
    prof_mark(ctx, 1);
    CUtensorMap tmA, tmB, tmBh;
    GFFM_TRY(make_plane_tmap(&tmA, pa, Kp, rowsPA, s, BM));
    GFFM_TRY(make_plane_tmap(&tmB, pb, Kp, rowsPB, s, BN));
    GFFM_TRY(make_plane_tmap(&tmBh, pb, Kp, rowsPB, s, BN / 2));
    p.m = (int)m;
    p.n = (int)n;
    p.num_kb = (int)(Kp / 128);
    p.num_m_blk = (int)(rowsPA / BM);
    p.num_n_blk = (int)(rowsPB / BN);
    p.batches = s;
    p.E = (uint8_t*)ctx->ws_eplanes.ptr;
    p.lde = lde;
    p.e_plane_stride = e_plane;
    GFFM_TRY(launch_gemm<SchemeRNS>(ctx, tmA, tmB, p, nullptr, &tmBh));
    prof_mark(ctx, 2);
    GFFM_TRY(launch_crt(ctx, ctx->stream, cp, (const uint8_t*)ctx->ws_eplanes.ptr, lde, e_plane, m, n, Cv.p, Cv.ld, kara_hi, ldhi));
    prof_mark(ctx, 3);
    if (kara_hi && K > kmax) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "Karatsuba carry split with K > 65536");
  }
  return GFFM_OK;
}

int32_t gffm_gemm_tc_rns(gffm_ctx* ctx, MatView C, MatView A, MatView B, uint64_t R, uint64_t P, int mode, bool balanced,
                         uint32_t* kara_hi, uint64_t kara_N1) {
  return gffm_gemm_tc_rns_ex(ctx, C, A, nullptr, B, nullptr, R, P, mode, balanced, kara_hi, C.ld, kara_N1);
}

// ---------------------------------------------------------------------------------------------------
// Tiled, multi-stream modular GEMM.  A is cut into row blocks and B into column panels; the 8-bit plane split of block
// t+1 and the CRT of the tiles of step t run on an auxiliary stream WHILE the tensor-core GEMM of step t occupies the
// compute stream (the GEMM is tensor-bound and leaves the CUDA cores almost idle, the helpers are issue-bound), so
// only the first split and the last CRT are exposed.  Two front ends:
//   device mode : operands are resident views (gffm_gemm on large shapes)
//   host mode   : gffm_gemm_host -- blocks arrive by H2D copies on a third stream and C tiles leave by D2H copies on
//                 a fourth, replacing the reference sequence CuModMatrix(A); CuModMatrix(B); mul!(C,A,B); Array(C)
//                 (reference CuModMatrix.jl:53-99, :767-787, :256-261) whose transfers and product are serialised.
// Tile (i,j) is issued in the step t = max(i,j) in which its last operand block becomes available.
// ---------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
mod_inplace_kernel(uint32_t* __restrict__ X, int64_t ld, int64_t rows, int64_t cols, const __grid_constant__ ModP mp) {
  const int64_t total = rows * cols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx / rows, i = idx - j * rows;
    uint32_t v = X[j * ld + i];
    if (v >= mp.P) X[j * ld + i] = (uint32_t)mod_u64(v, mp);
  }
}

struct HostIO {  // host mode only
  const uint32_t* A = nullptr;
  int64_t lda = 0;
  const uint32_t* B = nullptr;
  int64_t ldb = 0;
  uint32_t* C = nullptr;
  int64_t ldc = 0;
};

// device mode, B arriving in column panels (multi-GPU layer: NCCL broadcast of B on another stream): panel p =
// columns [off[p], off[p+1]); ready[p] is recorded by the caller once panel p is in place, consumed[p] is recorded
// here once panel p has been turned into operand planes (its residues may be overwritten from then on)
struct PanelFeed {
  int npanels = 0;
  const int64_t* off = nullptr;
  cudaEvent_t const* ready = nullptr;
  cudaEvent_t const* consumed = nullptr;
};

// returns true when the planes of X are already cached (no split needed); otherwise *out points to the buffer to fill
// (the owner's cache buffer, (re)allocated here, or the shared workspace) and *fill_cache says whether to validate it
int32_t plane_buffer(gffm_ctx* ctx, int role, const MatView& X, int64_t k, int64_t Kp, int64_t rowsP, const SplitParams& sp,
                     int balanced, uint64_t R, gffm_workspace* ws, uint8_t** out, bool* hit, gffm_plane_cache** fill_cache) {
  const size_t bytes = (size_t)sp.nplanes * rowsP * Kp;
  *hit = false;
  *fill_cache = nullptr;
  gffm_mat* own = X.owner;
  if (own) {
    gffm_plane_cache& c = own->cache[role];
    if (c.valid && c.version == own->version && c.mode == sp.mode && c.nplanes == sp.nplanes && c.balanced == balanced && c.R == R &&
        c.r0 == X.r0 && c.c0 == X.c0 && c.rows == X.rows && c.cols == X.cols && c.k0 == 0 && c.kc == k && c.Kp == Kp && c.rowsP == rowsP) {
      *out = (uint8_t*)c.ptr;
      *hit = true;
      return GFFM_OK;
    }
    if (c.bytes < bytes) {
      if (c.ptr) {
        gffm_dev_free(ctx, c.ptr);
        c.ptr = nullptr;
        c.bytes = 0;
      }
      if (gffm_dev_alloc(ctx, &c.ptr, bytes) == cudaSuccess) c.bytes = bytes;
      else c.ptr = nullptr;
    }
    if (c.ptr) {
      c.valid = false;
      c.mode = sp.mode; c.nplanes = sp.nplanes; c.balanced = balanced; c.R = R;
      c.r0 = X.r0; c.c0 = X.c0; c.rows = X.rows; c.cols = X.cols; c.k0 = 0; c.kc = k; c.Kp = Kp; c.rowsP = rowsP;
      *out = (uint8_t*)c.ptr;
      *fill_cache = &c;
      return GFFM_OK;
    }
  }
  GFFM_TRY(gffm_ws_reserve(ctx, ws, bytes));
  *out = (uint8_t*)ws->ptr;
  return GFFM_OK;
}

int32_t tiled_gemm(gffm_ctx* ctx, MatView Cv, MatView A, MatView B, const HostIO* io, int64_t m, int64_t n, int64_t k, uint64_t R, uint64_t P,
                   int mode, bool balanced, const PanelFeed* feed = nullptr) {
  const bool host = io != nullptr;
  const bool rns = R > 65536;
  const int L = R <= 256 ? 1 : 2;
  const int BN = rns ? SchemeRNS::BN : (L == 1 ? SchemeL1::BN : SchemeL2::BN);
  RnsPlan plan;
  SplitParams sp;
  int nplanes;
  if (rns) {
    GFFM_TRY(make_rns_plan(k, R, P, balanced, mode, 0, &plan));
    sp = plan.sp;
    nplanes = plan.s;
  } else {
    memset(&sp, 0, sizeof(sp));
    sp.nplanes = nplanes = L;
    sp.mode = 0;
  }
  // block boundaries.  Host mode: up to 8 blocks per operand with DECREASING sizes (weights 6,6,5,5,4,3,2,1): the tiles that
  // can only start after the last upload are the last block row/column, so small final blocks shorten the exposed tail.
  // Device mode: up to 4 equal blocks (fewer launches).
  auto boundaries = [&](int64_t extent, int64_t align) {
    std::vector<int64_t> off{0};
    const int64_t min_block = host ? 1024 : 2048;
    if (extent < 2 * min_block) {
      off.push_back(extent);
      return off;
    }
    if (!host) {
      const int64_t nb = std::min<int64_t>(4, extent / min_block);
      const int64_t b = round_up(ceil_div(extent, nb), align);
      for (int64_t o = b; o < extent; o += b) off.push_back(o);
      off.push_back(extent);
      return off;
    }
    static const int wts[8] = {6, 6, 5, 5, 4, 3, 2, 1};
    const int64_t nb = std::min<int64_t>(8, extent / min_block);
    int64_t wsum = 0;
    for (int i = 0; i < nb; ++i) wsum += wts[8 - nb + i];
    int64_t acc = 0;
    for (int i = 0; i < nb - 1; ++i) {
      acc += wts[8 - nb + i];
      int64_t o = round_up(extent * acc / wsum, align);
      if (o > off.back() && o < extent) off.push_back(o);
    }
    off.push_back(extent);
    return off;
  };
  std::vector<int64_t> offA = boundaries(m, BM), offB = boundaries(n, BN);
  // GFFM_HOST_SCHED=1d: upload B first (one block, copied and split in 8 column sub-panels), then 16 equal row blocks of A, each
  // multiplied with all of B as soon as it has arrived -- the GPU idles while B is in flight but then runs flat out, and every
  // finished row block of C leaves immediately (see profiles/r01_notes.md for the comparison with the 2-D interleaving)
  static const bool sched_1d = getenv("GFFM_HOST_SCHED") && !strcmp(getenv("GFFM_HOST_SCHED"), "1d");
  if (host && sched_1d && m >= 16 * 256) {
    offB = {0, n};
    offA.clear();
    const int64_t b = round_up(ceil_div(m, 16), BM);
    for (int64_t o = 0; o < m; o += b) offA.push_back(o);
    offA.push_back(m);
  }
  if (feed) {  // the caller's panels of B; A (this rank's row block) is one block, split once while panel 0 is in flight
    offA = {0, m};
    offB.assign(feed->off, feed->off + feed->npanels + 1);
  }
  const int nbA = (int)offA.size() - 1, nbB = (int)offB.size() - 1;
  const int64_t Kp = round_up(k, 128), rowsPA = round_up(m, BM), rowsPB = round_up(n, BN);
  const int64_t lde = round_up(m, 128), e_plane = lde * n;
  // buffers
  uint32_t *dA = A.p, *dB = B.p, *dC = Cv.p;
  int64_t ldA = A.ld, ldB = B.ld, ldC = Cv.ld;
  uint8_t *pA = nullptr, *pB = nullptr, *E = nullptr;
  bool hitA = false, hitB = false;
  gffm_plane_cache *fillA = nullptr, *fillB = nullptr;
  if (host) {
    ldA = round_up(m, 32); ldB = round_up(k, 32); ldC = round_up(m, 32);
    size_t off = 0;
    auto carve = [&](size_t bytes) {
      size_t o = off;
      off = (off + bytes + 255) & ~(size_t)255;
      return o;
    };
    const size_t oA = carve((size_t)ldA * k * 4), oB = carve((size_t)ldB * n * 4), oC = carve((size_t)ldC * n * 4);
    const size_t oPA = carve((size_t)nplanes * rowsPA * Kp), oPB = carve((size_t)nplanes * rowsPB * Kp);
    const size_t oE = rns ? carve((size_t)nplanes * e_plane) : 0;
    GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_host, off));
    char* base = (char*)ctx->ws_host.ptr;
    dA = (uint32_t*)(base + oA); dB = (uint32_t*)(base + oB); dC = (uint32_t*)(base + oC);
    pA = (uint8_t*)(base + oPA); pB = (uint8_t*)(base + oPB); E = (uint8_t*)(base + oE);
  } else {
    GFFM_TRY(plane_buffer(ctx, 0, A, k, Kp, rowsPA, sp, balanced ? 1 : 0, R, &ctx->ws_planes_a, &pA, &hitA, &fillA));
    GFFM_TRY(plane_buffer(ctx, 1, B, k, Kp, rowsPB, sp, balanced ? 1 : 0, R, &ctx->ws_planes_b, &pB, &hitB, &fillB));
    if (rns) {
      GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_eplanes, (size_t)nplanes * e_plane));
      E = (uint8_t*)ctx->ws_eplanes.ptr;
    }
  }
  if (!ctx->s_aux) GFFM_CUDA(cudaStreamCreateWithFlags(&ctx->s_aux, cudaStreamNonBlocking));
  if (host && !ctx->s_h2d) {
    GFFM_CUDA(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    GFFM_CUDA(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
  }
  cudaStream_t sc = ctx->stream, sx = ctx->s_aux, sh = ctx->s_h2d, sd = ctx->s_d2h;
  const int steps = std::max(nbA, nbB);
  // events: pool owned by the context (grow-only)
  const size_t need_ev = 2 + 3 * (size_t)steps + 2 * (size_t)nbA * nbB + 16;
  while (ctx->ev_pool.size() < need_ev) {
    cudaEvent_t e;
    GFFM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->ev_pool.push_back(e);
  }
  size_t evi = 0;
  auto next_ev = [&]() { return ctx->ev_pool[evi++]; };
  cudaEvent_t ev0 = next_ev(), ev_end = next_ev();
  GFFM_CUDA(cudaEventRecord(ev0, sc));  // inputs / workspaces may still be in use by earlier work on the compute stream
  GFFM_CUDA(cudaStreamWaitEvent(sx, ev0, 0));
  if (host) {
    GFFM_CUDA(cudaStreamWaitEvent(sh, ev0, 0));
    GFFM_CUDA(cudaStreamWaitEvent(sd, ev0, 0));
  }
  // GFFM_TILED_TRACE=1: timeline of every kernel of the pipeline (timing events around each launch, printed relative to the
  // start of the call; synchronises at the end of the call -- diagnostics only)
  static const bool trace_on = getenv("GFFM_TILED_TRACE") != nullptr;
  struct TraceRec {
    const char* what;
    int idx;
    cudaEvent_t a, b;
  };
  std::vector<TraceRec> trace;
  cudaEvent_t trace0 = nullptr;
  auto trace_begin = [&](cudaStream_t st) -> cudaEvent_t {
    if (!trace_on) return nullptr;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    return e;
  };
  auto trace_end = [&](const char* what, int idx, cudaEvent_t a, cudaStream_t st) {
    if (!trace_on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    trace.push_back(TraceRec{what, idx, a, e});
  };
  if (trace_on) trace0 = trace_begin(sc);
  CUtensorMap tmA, tmB, tmBh;
  GFFM_TRY(make_plane_tmap(&tmA, pA, Kp, rowsPA, nplanes, BM));
  GFFM_TRY(make_plane_tmap(&tmB, pB, Kp, rowsPB, nplanes, BN));
  GFFM_TRY(make_plane_tmap(&tmBh, pB, Kp, rowsPB, nplanes, BN / 2));
  const ModP mpR = make_modp(R), mpP = make_modp(P);
  const int nblk_mod = ctx->num_sms * 8;
  std::vector<cudaEvent_t> ev_split(steps);

  // step t: make block t of A and of B available as planes (aux stream), one step ahead of the GEMM
  auto prepare = [&](int t) -> int32_t {
    if (t < nbA && !hitA) {
      const int64_t i0 = offA[t], mi = offA[t + 1] - i0;
      if (host) {
        GFFM_CUDA(cudaMemcpy2DAsync(dA + i0, (size_t)ldA * 4, io->A + i0, (size_t)io->lda * 4, (size_t)mi * 4, (size_t)k, cudaMemcpyHostToDevice, sh));
        cudaEvent_t e = next_ev();
        GFFM_CUDA(cudaEventRecord(e, sh));
        GFFM_CUDA(cudaStreamWaitEvent(sx, e, 0));
        mod_inplace_kernel<<<nblk_mod, 256, 0, sx>>>(dA + i0, ldA, mi, k, mpR);
        GFFM_LAUNCH_CHECK(ctx);
      }
      MatView v{dA + i0, ldA, mi, k};
      cudaEvent_t ta = trace_begin(sx);
      GFFM_TRY(run_split(ctx, true, v, nullptr, 0, k, pA + i0 * Kp, Kp, rowsPA, sp, sx));
      trace_end("splitA", t, ta, sx);
    }
    if (t < nbB && feed && feed->ready && feed->ready[t]) GFFM_CUDA(cudaStreamWaitEvent(sx, feed->ready[t], 0));
    if (t < nbB && !hitB) {
      const int64_t jb = offB[t], njb = offB[t + 1] - jb;
      // a single B block (1-D schedule) is copied and split in 8 sub-panels so that the split overlaps the upload
      const int nsub = (host && nbB == 1 && njb >= 8 * BN) ? 8 : 1;
      const int64_t sub = round_up(ceil_div(njb, nsub), BN);
      for (int64_t j0 = jb; j0 < jb + njb; j0 += sub) {
        const int64_t nj = std::min(sub, jb + njb - j0);
        if (host) {
          GFFM_CUDA(cudaMemcpy2DAsync(dB + j0 * ldB, (size_t)ldB * 4, io->B + j0 * io->ldb, (size_t)io->ldb * 4, (size_t)k * 4, (size_t)nj,
                                      cudaMemcpyHostToDevice, sh));
          cudaEvent_t e = next_ev();
          GFFM_CUDA(cudaEventRecord(e, sh));
          GFFM_CUDA(cudaStreamWaitEvent(sx, e, 0));
          mod_inplace_kernel<<<nblk_mod, 256, 0, sx>>>(dB + j0 * ldB, ldB, k, nj, mpR);
          GFFM_LAUNCH_CHECK(ctx);
        }
        MatView v{dB + j0 * ldB, ldB, k, nj};
        cudaEvent_t tb = trace_begin(sx);
        GFFM_TRY(run_split(ctx, false, v, nullptr, 0, k, pB + j0 * Kp, Kp, rowsPB, sp, sx));
        trace_end("splitB", t, tb, sx);
      }
    }
    ev_split[t] = next_ev();
    GFFM_CUDA(cudaEventRecord(ev_split[t], sx));
    if (t < nbB && feed && feed->consumed && feed->consumed[t]) GFFM_CUDA(cudaEventRecord(feed->consumed[t], sx));
    return GFFM_OK;
  };
  struct Pending {
    int i, j;
    cudaEvent_t done;
  };
  std::vector<Pending> pending;
  auto gemm_tile = [&](int i, int j) -> int32_t {
    const int64_t i0 = offA[i], mi = offA[i + 1] - i0, j0 = offB[j], nj = offB[j + 1] - j0;
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.m = (int)mi;
    p.n = (int)nj;
    p.num_kb = (int)(Kp / 128);
    p.num_m_blk = (int)ceil_div(mi, BM);
    p.num_n_blk = (int)ceil_div(nj, BN);
    p.mb0 = (int)(i0 / BM);
    p.nb0 = (int)(j0 / BN);
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (ctx->profile) {  // kernel-only duration of every GEMM launch (the compute stream carries nothing else)
      GFFM_CUDA(cudaEventCreate(&t0));
      GFFM_CUDA(cudaEventCreate(&t1));
      GFFM_CUDA(cudaEventRecord(t0, sc));
    }
    cudaEvent_t tg = trace_begin(sc);
    if (rns) {
      memcpy(p.mods, plan.mods, sizeof(p.mods));
      p.batches = plan.s;
      p.E = E + j0 * lde + i0;
      p.lde = lde;
      p.e_plane_stride = e_plane;
      GFFM_TRY(launch_gemm<SchemeRNS>(ctx, tmA, tmB, p, sc, &tmBh));
    } else {
      p.batches = 1;
      p.C = dC + j0 * ldC + i0;
      p.ldc = ldC;
      p.mode = mode;
      p.modP = mpP;
      if (L == 1) GFFM_TRY(launch_gemm<SchemeL1>(ctx, tmA, tmB, p, sc, &tmBh));
      else GFFM_TRY(launch_gemm<SchemeL2>(ctx, tmA, tmB, p, sc, &tmBh));
    }
    trace_end("gemm", i * 100 + j, tg, sc);
    if (ctx->profile) {
      GFFM_CUDA(cudaEventRecord(t1, sc));
      ctx->tile_events.push_back(t0);
      ctx->tile_events.push_back(t1);
    }
    cudaEvent_t e = next_ev();
    GFFM_CUDA(cudaEventRecord(e, sc));
    pending.push_back(Pending{i, j, e});
    return GFFM_OK;
  };
  // after the GEMM of a tile: CRT (aux stream) and, in host mode, the D2H copy of the finished C tile
  auto finish_tiles = [&]() -> int32_t {
    for (const Pending& t : pending) {
      const int64_t i0 = offA[t.i], mi = offA[t.i + 1] - i0, j0 = offB[t.j], nj = offB[t.j + 1] - j0;
      cudaEvent_t ready = t.done;
      if (rns) {
        GFFM_CUDA(cudaStreamWaitEvent(sx, t.done, 0));
        cudaEvent_t tc = trace_begin(sx);
        GFFM_TRY(launch_crt(ctx, sx, plan.cp, E + j0 * lde + i0, lde, e_plane, mi, nj, dC + j0 * ldC + i0, ldC, nullptr, 0));
        trace_end("crt", t.i * 100 + t.j, tc, sx);
        ready = next_ev();
        GFFM_CUDA(cudaEventRecord(ready, sx));
      }
      if (host) {
        GFFM_CUDA(cudaStreamWaitEvent(sd, ready, 0));
        GFFM_CUDA(cudaMemcpy2DAsync(io->C + j0 * io->ldc + i0, (size_t)io->ldc * 4, dC + j0 * ldC + i0, (size_t)ldC * 4, (size_t)mi * 4, (size_t)nj,
                                    cudaMemcpyDeviceToHost, sd));
      }
    }
    pending.clear();
    return GFFM_OK;
  };

  if (ctx->profile) {
    for (auto e : ctx->tile_events) cudaEventDestroy(e);
    ctx->tile_events.clear();
    ctx->n_ev = 0;
  }
  const bool hprof = host && getenv("GFFM_HOST_PROF") != nullptr;
  cudaEvent_t hp[4] = {nullptr, nullptr, nullptr, nullptr};
  if (hprof) {
    for (auto& e : hp) cudaEventCreate(&e);
    cudaEventRecord(hp[0], sh);
  }
  int32_t st = prepare(0);
  for (int t = 0; t < steps && st == GFFM_OK; ++t) {
    if (t + 1 < steps) st = prepare(t + 1);  // issued first so that it overlaps the GEMMs of step t
    if (st != GFFM_OK) break;
    GFFM_CUDA(cudaStreamWaitEvent(sc, ev_split[t], 0));
    for (int i = 0; i < std::min(t, nbA) && st == GFFM_OK; ++i)
      if (t < nbB) st = gemm_tile(i, t);
    for (int j = 0; j < std::min(t, nbB) && st == GFFM_OK; ++j)
      if (t < nbA) st = gemm_tile(t, j);
    if (t < nbA && t < nbB && st == GFFM_OK) st = gemm_tile(t, t);
    if (st == GFFM_OK) st = finish_tiles();
  }
  if (hprof) {
    cudaEventRecord(hp[1], sh);
    cudaEventRecord(hp[2], sc);
    cudaEventRecord(hp[3], sd);
  }
  // the compute stream continues only after the helpers are done (C complete, planes reusable)
  cudaEventRecord(ev_end, sx);
  cudaStreamWaitEvent(sc, ev_end, 0);
  if (trace_on) {
    cudaEvent_t tend = trace_begin(sc);
    cudaStreamSynchronize(sc);
    float total = 0.f;
    cudaEventElapsedTime(&total, trace0, tend);
    fprintf(stderr, "[tiled trace] call: %.3f ms (blocks %dx%d)\n", total, nbA, nbB);
    for (const TraceRec& r : trace) {
      float a = 0.f, b = 0.f;
      cudaEventElapsedTime(&a, trace0, r.a);
      cudaEventElapsedTime(&b, trace0, r.b);
      fprintf(stderr, "[tiled trace]   %-7s %4d  %8.3f -> %8.3f  (%.3f ms)\n", r.what, r.idx, a, b, b - a);
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    cudaEventDestroy(trace0);
    cudaEventDestroy(tend);
  }
  if (host) {
    cudaStreamSynchronize(sd);
    cudaStreamSynchronize(sh);
    cudaStreamSynchronize(sc);
    if (hprof) {
      float a = 0, b2 = 0, c2 = 0;
      cudaEventElapsedTime(&a, hp[0], hp[1]);
      cudaEventElapsedTime(&b2, hp[0], hp[2]);
      cudaEventElapsedTime(&c2, hp[0], hp[3]);
      fprintf(stderr, "[gemm_host] blocks %dx%d: H2D done %.2f ms, last GEMM done %.2f ms, last D2H done %.2f ms\n", nbA, nbB, a, b2, c2);
      for (auto& e : hp) cudaEventDestroy(e);
    }
  }
  if (st == GFFM_OK) {
    if (fillA) {
      fillA->valid = true;
      fillA->version = A.owner->version;
    }
    if (fillB) {
      fillB->valid = true;
      fillB->version = B.owner->version;
    }
    if (host) GFFM_CUDA(cudaGetLastError());
  }
  return st;
}
}  // namespace

// largest inner dimension one launch may accumulate exactly (int32 accumulators) for inputs < R
int64_t gffm_gemm_kchunk(uint64_t R, bool rns) { return rns ? 65536 : limb_kmax(R); }

// device-resident front end: used by the dispatcher for large single-chunk products
int32_t gffm_gemm_tiled(gffm_ctx* ctx, MatView C, MatView A, MatView B, uint64_t R, uint64_t P, int mode, bool balanced) {
  return tiled_gemm(ctx, C, A, B, nullptr, A.rows, B.cols, A.cols, R, P, mode, balanced);
}

extern "C" int32_t gffm_gemm_host(gffm_ctx* ctx, void* C_host, int64_t ldc, const void* A_host, int64_t lda, const void* B_host,
                                  int64_t ldb, int64_t m, int64_t n, int64_t k, int32_t dtype, uint64_t N) {
  GFFM_ENTER_CTX(ctx);
  if (!ctx || !C_host || !A_host || !B_host) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (m <= 0 || n <= 0 || k <= 0) GFFM_FAIL(GFFM_ERR_INVALID, "empty product");
  if (lda < m || ldb < k || ldc < m) GFFM_FAIL(GFFM_ERR_INVALID, "leading dimension too small");
  if (dtype != GFFM_U32) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "gffm_gemm_host takes uint32 residues (use upload/gemm/download for other host types)");
  if (N < 2 || N >= (1ull << 32)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "2 <= N < 2^32 required");
  if (!gffm_tc_available(ctx)) GFFM_FAIL(GFFM_ERR_CUDA, "tensor-map encoder unavailable");
  cudaSetDevice(ctx->device);
  const int64_t kmax = N > 65536 ? 65536 : limb_kmax(N);
  if (k > kmax)
    GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "inner dimension %lld exceeds one accumulation chunk (%lld) of the pipelined host GEMM", (long long)k, (long long)kmax);
  HostIO io;
  io.A = (const uint32_t*)A_host; io.lda = lda;
  io.B = (const uint32_t*)B_host; io.ldb = ldb;
  io.C = (uint32_t*)C_host; io.ldc = ldc;
  MatView none{nullptr, 0, 0, 0};
  return tiled_gemm(ctx, none, none, none, &io, m, n, k, N, N, GFFM_GEMM_STORE, /*balanced=*/true);
}

// Sharded product step of the multi-GPU layer (new; the reference is single-GPU).  This rank's row block A is multiplied with
// a B that ARRIVES in column panels (NCCL broadcast issued by the caller on its own stream).  The 8-bit plane split of panel
// p+1 and the CRT of panel p run on the auxiliary stream while the tensor-core GEMM of panel p occupies the compute stream;
// consumed[p] lets the caller start the NEXT step's broadcast of panel p while this step is still multiplying.
extern "C" int32_t gffm_gemm_panels(gffm_mat* C, gffm_mat* A, gffm_mat* B, int32_t npanels, const int64_t* col_off, void* const* ready,
                                    void* const* consumed, uint64_t R, uint64_t P) {
  GFFM_ENTER_MAT(C);
  if (!C || !A || !B || !col_off) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  GFFM_NARROW_ONLY(C, A, B);
  if (C == A || C == B) GFFM_FAIL(GFFM_ERR_INVALID, "gemm_panels: C must not alias an operand");
  if (!P && (A->N != B->N || A->N != C->N)) GFFM_FAIL(GFFM_ERR_MODULUS_MISMATCH, "gemm operands have different moduli");
  if (A->cols != B->rows || C->rows != A->rows || C->cols != B->cols)
    GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "gemm_panels: C %lldx%lld = A %lldx%lld * B %lldx%lld", (long long)C->rows, (long long)C->cols,
              (long long)A->rows, (long long)A->cols, (long long)B->rows, (long long)B->cols);
  if (npanels < 1 || col_off[0] != 0 || col_off[npanels] != B->cols) GFFM_FAIL(GFFM_ERR_INVALID, "gemm_panels: panels must cover [0, cols(B))");
  for (int p = 0; p < npanels; ++p)
    if (col_off[p + 1] < col_off[p]) GFFM_FAIL(GFFM_ERR_INVALID, "gemm_panels: panel offsets must be non-decreasing");
  if (!P) P = C->N;
  if (!R) R = A->N > B->N ? A->N : B->N;
  if (P >= (1ull << 32)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "gemm needs 0 < P < 2^32");
  gffm_ctx* ctx = C->ctx;
  cudaSetDevice(ctx->device);
  gffm_touch(C);
  const int64_t m = A->rows, k = A->cols, n = B->cols;
  bool aligned = true;
  for (int p = 1; p < npanels; ++p) aligned = aligned && (col_off[p] % 256 == 0);  // tile offsets into the operand planes
  for (int p = 0; p < npanels; ++p) aligned = aligned && (col_off[p + 1] > col_off[p]);
  const bool rns = R > 65536;
  const bool pipelined = aligned && R < (1ull << 32) && gffm_tc_available(ctx) && k >= 128 && k <= gffm_gemm_kchunk(R, rns) && m >= 128 && n >= 256;
  if (!pipelined) {  // small or unaligned shapes: panel by panel on the compute stream through the ordinary dispatcher
    for (int p = 0; p < npanels; ++p) {
      if (ready && ready[p]) GFFM_CUDA(cudaStreamWaitEvent(ctx->stream, (cudaEvent_t)ready[p], 0));
      const int64_t c0 = col_off[p], nc = col_off[p + 1] - c0;
      if (nc > 0 && m > 0)
        GFFM_TRY(gffm_gemm_views(ctx, sub_view(view_of(C), 0, c0, m, nc), cached_view_of(A), sub_view(view_of(B), 0, c0, k, nc), R, P, GFFM_GEMM_STORE,
                                 GFFM_ALGO_AUTO));
      if (consumed && consumed[p]) GFFM_CUDA(cudaEventRecord((cudaEvent_t)consumed[p], ctx->stream));
    }
    return GFFM_OK;
  }
  PanelFeed feed;
  feed.npanels = npanels;
  feed.off = col_off;
  feed.ready = (cudaEvent_t const*)ready;
  feed.consumed = (cudaEvent_t const*)consumed;
  // B is never taken from / put into the plane cache here: its residues change behind the library's back every step
  return tiled_gemm(ctx, view_of(C), cached_view_of(A), view_of(B), nullptr, m, n, k, R, P, GFFM_GEMM_STORE, rns && (R % P) == 0, &feed);
}

// ---------------------------------------------------------------------------------------------------
// B planes produced elsewhere (multi-GPU layer, mg.cu): plan / split / multiply as three separate steps (gemm_internal.cuh)
// ---------------------------------------------------------------------------------------------------
#include "gemm_internal.cuh"

struct GemmBPlan {
  bool rns = false;
  int L = 1;
  int64_t kc = 0, Kp = 0;
  uint64_t R = 0, P = 0, kara_N1 = 0;
  bool balanced = false;
  int mode = GFFM_GEMM_STORE;
  RnsPlan plan;
  SplitParams sp;
  BPlaneSpec spec;
};

int32_t gffm_bplan_create(int64_t kc, uint64_t R, uint64_t P, bool balanced, int crt_mode, uint64_t kara_N1, GemmBPlan** out) {
  if (!out) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (kc < 1) GFFM_FAIL(GFFM_ERR_INVALID, "empty inner dimension");
  if (R == 0 || R >= (1ull << 32) || P == 0 || P > (1ull << 52)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "product plan needs R < 2^32, P <= 2^52");
  GemmBPlan* g = new GemmBPlan();
  g->kc = kc;
  g->Kp = round_up(kc, 128);
  g->R = R;
  g->P = P;
  g->kara_N1 = kara_N1;
  g->mode = crt_mode;
  g->rns = kara_N1 != 0 || R > 65536 || P >= (1ull << 32);
  if (!kara_N1 && P >= (1ull << 32)) {
    delete g;
    GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "uint32 output needs P < 2^32");
  }
  if (g->rns) {
    if (kc > 65536) {
      delete g;
      GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "one RNS chunk covers at most 65536 inner columns");
    }
    g->balanced = balanced;
    const int32_t st = make_rns_plan(kc, R, P, balanced, crt_mode, kara_N1, &g->plan);
    if (st != GFFM_OK) {
      delete g;
      return st;
    }
    g->sp = g->plan.sp;
    g->spec = BPlaneSpec{g->plan.s, SchemeRNS::BN, g->Kp, 1};
  } else {
    g->L = R <= 256 ? 1 : 2;
    if (kc > limb_kmax(R)) {
      delete g;
      GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "inner dimension %lld exceeds one limb accumulation chunk (%lld)", (long long)kc, (long long)limb_kmax(R));
    }
    memset(&g->sp, 0, sizeof(g->sp));
    g->sp.nplanes = g->L;
    g->sp.mode = 0;
    g->spec = BPlaneSpec{g->L, g->L == 1 ? SchemeL1::BN : SchemeL2::BN, g->Kp, 0};
  }
  *out = g;
  return GFFM_OK;
}

void gffm_bplan_destroy(GemmBPlan* plan) { delete plan; }
const BPlaneSpec* gffm_bplan_spec(const GemmBPlan* plan) { return &plan->spec; }

int32_t gffm_bplan_split(gffm_ctx* ctx, const GemmBPlan* g, MatView B, const MatView* B2, uint8_t* planes, int64_t rowsPB, int64_t row0,
                         cudaStream_t st) {
  if (B.rows != g->kc) GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "plane split: %lld rows, plan has %lld", (long long)B.rows, (long long)g->kc);
  if (B.cols == 0) return GFFM_OK;
  if (row0 < 0 || row0 + B.cols > rowsPB) GFFM_FAIL(GFFM_ERR_INVALID, "plane split: rows [%lld, %lld) outside the buffer", (long long)row0, (long long)(row0 + B.cols));
  return run_split(ctx, false, B, B2, 0, g->kc, planes + row0 * g->Kp, g->Kp, rowsPB, g->sp, st);
}

int32_t gffm_bplan_split_push(gffm_ctx* ctx, const GemmBPlan* g, MatView B, const MatView* B2, uint8_t* const* plane_bufs, int nbufs, int64_t rowsPB,
                              int64_t row0, cudaStream_t st, int push_sms) {
  if (B.rows != g->kc) GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "plane split: %lld rows, plan has %lld", (long long)B.rows, (long long)g->kc);
  if (B.cols == 0 || nbufs <= 0) return GFFM_OK;
  if (nbufs > 32) GFFM_FAIL(GFFM_ERR_INVALID, "at most 32 destinations");
  if (row0 < 0 || row0 + B.cols > rowsPB) GFFM_FAIL(GFFM_ERR_INVALID, "plane split: rows outside the buffer");
  PlaneDests d;
  d.n = nbufs;
  d.sms = push_sms;
  for (int i = 0; i < nbufs; ++i) d.p[i] = plane_bufs[i] + row0 * g->Kp;
  return run_split(ctx, false, B, B2, 0, g->kc, nullptr, g->Kp, rowsPB, g->sp, st, &d);
}

int32_t gffm_bplan_gemm(gffm_ctx* ctx, const GemmBPlan* g, MatView Cv, MatView A, const MatView* A2, int64_t n, const ExtBPlanes& ext,
                        uint32_t* kara_hi, int64_t ldhi) {
  const int64_t m = A.rows;
  if (A.cols != g->kc || Cv.rows != m || Cv.cols != n) GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "external-plane GEMM: inconsistent shapes");
  if (m == 0 || n == 0) return GFFM_OK;
  if (!ext.planes || !ext.off || ext.npanels < 1 || ext.off[ext.npanels] != n) GFFM_FAIL(GFFM_ERR_INVALID, "external-plane GEMM: bad panel list");
  if ((g->kara_N1 != 0) != (kara_hi != nullptr)) GFFM_FAIL(GFFM_ERR_INVALID, "external-plane GEMM: carry output does not match the plan");
  const int BN = g->spec.BN, nplanes = g->spec.nplanes;
  const int64_t Kp = g->Kp, rowsPA = round_up(m, BM);
  for (int p = 0; p < ext.npanels; ++p)
    if ((ext.off[p + 1] > ext.off[p] && ext.off[p] % BN != 0) || ext.off[p + 1] < ext.off[p]) GFFM_FAIL(GFFM_ERR_INVALID, "external-plane GEMM: panel offsets must be multiples of %d", BN);
  if (ext.rowsPB < round_up(n, BN)) GFFM_FAIL(GFFM_ERR_INVALID, "external-plane GEMM: plane buffer has too few rows");
  uint8_t* pa = nullptr;
  {
    cudaEvent_t ta = gffm_trace_begin(ctx, ctx->stream);
    GFFM_TRY(acquire_planes(ctx, 0, A, A2, 0, g->kc, Kp, rowsPA, g->sp, g->balanced ? 1 : 0, g->R, &ctx->ws_planes_a, &pa));
    gffm_trace_end(ctx, "splitA", 0, 0, ta, ctx->stream);
  }
  const int64_t lde = round_up(m, 128), e_plane = lde * n;
  uint8_t* E = nullptr;
  if (g->rns) {
    GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_eplanes, (size_t)nplanes * e_plane));
    E = (uint8_t*)ctx->ws_eplanes.ptr;
  }
  if (!ctx->s_aux) GFFM_CUDA(cudaStreamCreateWithFlags(&ctx->s_aux, cudaStreamNonBlocking));
  cudaStream_t sc = ctx->stream, sx = ctx->s_aux;
  const size_t need_ev = 2 + (size_t)ext.npanels + 4;
  while (ctx->ev_pool.size() < need_ev) {
    cudaEvent_t e;
    GFFM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->ev_pool.push_back(e);
  }
  size_t evi = 0;
  cudaEvent_t ev0 = ctx->ev_pool[evi++], ev_end = ctx->ev_pool[evi++];
  GFFM_CUDA(cudaEventRecord(ev0, sc));
  GFFM_CUDA(cudaStreamWaitEvent(sx, ev0, 0));
  CUtensorMap tmA, tmB, tmBh;
  GFFM_TRY(make_plane_tmap(&tmA, pa, Kp, rowsPA, nplanes, BM));
  GFFM_TRY(make_plane_tmap(&tmB, (void*)ext.planes, Kp, ext.rowsPB, nplanes, BN));
  GFFM_TRY(make_plane_tmap(&tmBh, (void*)ext.planes, Kp, ext.rowsPB, nplanes, BN / 2));
  const ModP mpP = make_modp(g->P);
  for (int idx = 0; idx < ext.npanels; ++idx) {
    const int p = ext.order ? ext.order[idx] : idx;
    const int64_t j0 = ext.off[p], nj = ext.off[p + 1] - j0;
    if (nj <= 0) continue;
    cudaEvent_t tw = gffm_trace_begin(ctx, sc);
    if (ext.ready && ext.ready[p]) GFFM_CUDA(cudaStreamWaitEvent(sc, ext.ready[p], 0));
    gffm_trace_end(ctx, "waitB", p, 0, tw, sc);
    cudaEvent_t tg = gffm_trace_begin(ctx, sc);
    GemmParams q;
    memset(&q, 0, sizeof(q));
    q.m = (int)m;
    q.n = (int)nj;
    q.num_kb = (int)(Kp / 128);
    q.num_m_blk = (int)(rowsPA / BM);
    q.num_n_blk = (int)ceil_div(nj, BN);
    q.nb0 = (int)(j0 / BN);
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (ctx->profile) {
      GFFM_CUDA(cudaEventCreate(&t0));
      GFFM_CUDA(cudaEventCreate(&t1));
      GFFM_CUDA(cudaEventRecord(t0, sc));
    }
    if (g->rns) {
      memcpy(q.mods, g->plan.mods, sizeof(q.mods));
      q.batches = g->plan.s;
      q.E = E + j0 * lde;
      q.lde = lde;
      q.e_plane_stride = e_plane;
      GFFM_TRY(launch_gemm<SchemeRNS>(ctx, tmA, tmB, q, sc, &tmBh));
    } else {
      q.batches = 1;
      q.C = Cv.p + j0 * Cv.ld;
      q.ldc = Cv.ld;
      q.mode = g->mode;
      q.modP = mpP;
      if (g->L == 1) GFFM_TRY(launch_gemm<SchemeL1>(ctx, tmA, tmB, q, sc, &tmBh));
      else GFFM_TRY(launch_gemm<SchemeL2>(ctx, tmA, tmB, q, sc, &tmBh));
    }
    if (ctx->profile) {
      GFFM_CUDA(cudaEventRecord(t1, sc));
      ctx->tile_events.push_back(t0);
      ctx->tile_events.push_back(t1);
    }
    gffm_trace_end(ctx, "gemm", p, 0, tg, sc);
    if (g->rns) {
      cudaEvent_t e = ctx->ev_pool[evi++];
      GFFM_CUDA(cudaEventRecord(e, sc));
      GFFM_CUDA(cudaStreamWaitEvent(sx, e, 0));
      cudaEvent_t tc = gffm_trace_begin(ctx, sx);
      GFFM_TRY(launch_crt(ctx, sx, g->plan.cp, E + j0 * lde, lde, e_plane, m, nj, Cv.p + j0 * Cv.ld, Cv.ld, kara_hi ? kara_hi + j0 * ldhi : nullptr, ldhi));
      gffm_trace_end(ctx, "crt", p, 1, tc, sx);
    }
  }
  GFFM_CUDA(cudaEventRecord(ev_end, sx));
  GFFM_CUDA(cudaStreamWaitEvent(sc, ev_end, 0));
  return GFFM_OK;
}

// profiling hook of the multi-GPU layer: forget the GEMM launch events of earlier products
void gffm_profile_reset_tiles(gffm_ctx* ctx) {
  for (auto e : ctx->tile_events) cudaEventDestroy(e);
  ctx->tile_events.clear();
  ctx->n_ev = 0;
}
