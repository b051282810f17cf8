// pluq.cu -- blocked right-looking elimination over Z/N (N prime, N < 2^32): PLUQ / LU / RREF / rank / inverse
// and the triangular inverse.  Replaces the reference's unblocked host-driven loop (one pivot per iteration,
// >= 6 launches + a device->host findmax + CUDA.synchronize per pivot:
// reference src/CuModMatrix/rref_lu_pluq/pluq_kernels.jl:46-157) by
//
//   outer blocks of NB columns   trailing update A22 -= L21*U12 and U12 = L11^-1*A12 on the tensor-core GEMM
//   inner panels of w columns    ONE thread-block-cluster kernel per panel: the panel lives in the distributed
//                                shared memory of 8 CTAs, pivot search = warp-shuffle argmax + DSMEM exchange,
//                                two cluster barriers per pivot, no host round trip, batched row swaps
//
// Conventions kept from the reference (parity): pivot = MAXIMUM residue at/below the current row, first index on
// ties (findmax, pluq_kernels.jl:189); the pivot row is scaled to 1 (:314) so U has unit pivots and L carries the
// pivot value on its diagonal with the un-normalised sub-column below it (:343,:389); row permutation returned as
// an ordered list of transpositions (:254-260).  A column without pivot is skipped (row echelon form); pluq's
// column permutation moves the pivot columns to the front (GFFM_PIVOT_CORRECT, see DESIGN.md for why the
// reference's swap-with-fixed-last-column quirk is not reproduced by the blocked path).
#include <cooperative_groups.h>
#include <algorithm>
#include <chrono>
#include <stdlib.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int PANEL_THREADS_MAX = 512;
constexpr int PANEL_CLUSTER = 8;
constexpr int PANEL_W_MAX = 32;
constexpr int PANEL_SMEM_BUDGET = 200 * 1024;

struct PluqState {
  int r;         // pivots found so far (== next pivot row)
  int rb;        // value of r when the last panel started
  int n_gather;  // entries of the composed row-gather list of the last panel
  int pad;
};

struct PluqBufs {
  PluqState* st;
  int* pivcol;       // [t] column of pivot t
  uint32_t* pinv;    // [t] inverse of pivot t
  int* swp;          // [t] row swapped with row t when pivot t was chosen
  int* g_dst;        // composed gather list of the last panel: W[g_dst[e]] <- old W[g_src[e]]
  int* g_src;
  const uint32_t* inv_table;  // [x] = x^-1 mod N for N <= 2^26 (batched once per modulus), else nullptr
  long long* prof;            // optional per-phase cycle counters of the panel kernel (GFFM_PANEL_PROF=1), else nullptr
};

__device__ __forceinline__ bool better(uint32_t v, int i, uint32_t bv, int bi) { return v > bv || (v == bv && i < bi); }

// ---------------------------------------------------------------------------------------------------
// panel kernel: columns [j0, j0+w) of W, rows [r, m).  One cluster; CTA c owns rows rb + c*rows_c + [0,rows_c).
// The panel stays in (distributed) shared memory for all w pivots; the L multipliers of a pivot are kept IN PLACE in
// its own panel column (below the pivot) so that row swaps move them for free -- no global traffic inside the pivot
// loop.  Per pivot: local argmax (warp shuffles) + the candidate's modular inverse (table lookup / 32-bit Euclid,
// computed by every CTA for its own candidate BEFORE the barrier), cluster barrier, DSMEM read of the 8/16
// candidates, pivot row / current row published in shared memory, cluster barrier, DSMEM read, rank-1 update.
// ---------------------------------------------------------------------------------------------------
template <bool SMALL>
__device__ __forceinline__ uint32_t panel_mulmod(uint32_t a, uint32_t b, const ModP& mp, uint32_t mu32) {
  if constexpr (SMALL) {  // N <= 2^16: 32-bit product, 32-bit Barrett
    const uint32_t x = a * b;
    uint32_t r = x - __umulhi(x, mu32) * (uint32_t)mp.P;
    if (r >= (uint32_t)mp.P) r -= (uint32_t)mp.P;
    return r;
  } else {
    return mulmod_u32(a, b, mp);
  }
}

// (a + nl * u) mod P with a, nl, u < P: one multiply-add and one Barrett step per eliminated element
template <bool SMALL>
__device__ __forceinline__ uint32_t panel_fmamod(uint32_t a, uint32_t nl, uint32_t u, const ModP& mp, uint32_t mu32) {
  if constexpr (SMALL) {  // P <= 2^16: nl*u + a < 2^32
    const uint32_t x = nl * u + a;
    const uint32_t r = x - __umulhi(x, mu32) * (uint32_t)mp.P;  // in [0, 2P)
    return min(r, r - (uint32_t)mp.P);                          // unsigned: picks r - P when r >= P
  } else {
    return (uint32_t)mod_u64((uint64_t)nl * u + a, mp);
  }
}

// compose the transpositions rb..r-1 of one panel into a single gather list (executed by one warp of CTA 0)
__device__ __forceinline__ void compose_gather_list(const PluqBufs& b, int rb, int r, int m, int lane) {

    const int k = r - rb;  // <= 32
    int near = rb + lane;  // content (original row) now sitting at position rb+lane
    int far_pos = -1, far_src = -1, nf = 0;
    for (int s = 0; s < k; ++s) {
      const int p = b.swp[rb + s];  // written by tid 0 of this CTA above
      if (p == rb + s) continue;
      const int mine = __shfl_sync(0xffffffffu, near, s);
      if (p < rb + 32) {
        const int other = __shfl_sync(0xffffffffu, near, p - rb);
        if (lane == s) near = other;
        if (lane == p - rb) near = mine;
      } else {
        unsigned hit = __ballot_sync(0xffffffffu, lane < nf && far_pos == p);
        int e;
        if (hit) e = __ffs(hit) - 1;
        else {
          e = nf++;
          if (lane == e) {
            far_pos = p;
            far_src = p;
          }
        }
        const int other = __shfl_sync(0xffffffffu, far_src, e);
        if (lane == s) near = other;
        if (lane == e) far_src = mine;
      }
    }
    int cnt = 0;
    // near entries that moved
    const bool nm = lane < 32 && near != rb + lane && (rb + lane) < m;
    unsigned nmask = __ballot_sync(0xffffffffu, nm);
    if (nm) {
      const int e = __popc(nmask & ((1u << lane) - 1));
      b.g_dst[e] = rb + lane;
      b.g_src[e] = near;
    }
    cnt = __popc(nmask);
    const bool fm = lane < nf && far_src != far_pos;
    unsigned fmask = __ballot_sync(0xffffffffu, fm);
    if (fm) {
      const int e = cnt + __popc(fmask & ((1u << lane) - 1));
      b.g_dst[e] = far_pos;
      b.g_src[e] = far_src;
    }
    cnt += __popc(fmask);
    if (lane == 0) {
      b.st->rb = rb;
      b.st->n_gather = cnt;
    }
  }

template <bool SMALL, int PANEL_THREADS>
__global__ void __launch_bounds__(PANEL_THREADS, 1)
pluq_panel_kernel(uint32_t* __restrict__ W, int64_t ldw, int m, int j0, int w, uint32_t* __restrict__ Lm, int64_t ldl,
                  PluqBufs b, const __grid_constant__ ModP mp, uint32_t mu32) {
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int NWARPS = PANEL_THREADS / 32;
  constexpr int MAXC = 16;  // largest cluster
  extern __shared__ uint32_t panel[];  // [w][rows_c]
  // published per pivot (double-buffered by parity): candidate (value, row, inverse), its row, and -- on its owner -- row r
  __shared__ uint32_t cand_val[2], cand_inv[2];
  __shared__ int cand_idx[2];
  __shared__ uint32_t rowbuf_p[2][PANEL_W_MAX], rowbuf_r[2][PANEL_W_MAX];
  // block-local scratch
  __shared__ uint32_t red_val[NWARPS], red_inv[NWARPS];
  __shared__ int red_idx[NWARPS];
  __shared__ uint32_t g_val[MAXC], g_inv[MAXC];  // gathered candidates of all CTAs
  __shared__ int g_idx[MAXC];
  __shared__ uint32_t allrows[MAXC][PANEL_W_MAX];
  __shared__ uint32_t u_s[PANEL_W_MAX], oldr_s[PANEL_W_MAX];
  __shared__ int colpiv[PANEL_W_MAX];       // panel column -> pivot ordinal within this panel, or -1
  __shared__ uint32_t pivval[PANEL_W_MAX];  // pivot value of ordinal s

  const int rb = b.st->r;
  const int rows_total = m - rb;
  const int rows_c = (rows_total + cs - 1) / cs;
  const int my_lo = rb + rank * rows_c;
  int my_n = rows_total - rank * rows_c;
  my_n = my_n < 0 ? 0 : (my_n > rows_c ? rows_c : my_n);
  const uint32_t P = (uint32_t)mp.P;
  const bool do_prof = b.prof != nullptr && rank == 0 && tid == 0;
  long long tprev = do_prof ? clock64() : 0;
#define PANEL_TICK(slot)                         \
  if (do_prof) {                                 \
    const long long tn = clock64();              \
    b.prof[slot] += tn - tprev;                  \
    tprev = tn;                                  \
  }

  for (int c = 0; c < w; ++c)
    for (int q = tid; q < my_n; q += PANEL_THREADS) panel[c * rows_c + q] = W[(int64_t)(j0 + c) * ldw + my_lo + q];
  if (tid < PANEL_W_MAX) colpiv[tid] = -1;
  cluster.sync();  // everybody has read st->r before anyone can finish and overwrite it
  PANEL_TICK(0)

  int r = rb;
  for (int jj = 0; jj < w && r < m; ++jj) {
    const int par = jj & 1;
    // ---- phase A: local argmax (max residue, first index) over rows >= r of column jj
    uint32_t bv = 0;
    int bi = 0x7fffffff;
    for (int q = tid; q < my_n; q += PANEL_THREADS) {
      const int i = my_lo + q;
      if (i >= r) {
        const uint32_t v = panel[jj * rows_c + q];
        if (v > bv) {
          bv = v;
          bi = i;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint32_t ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      // inverse of this WARP's candidate: every warp does it at once, so the table-load latency (or the Euclid chain)
      // is paid once per pivot and overlaps the block barrier below
      red_val[warp] = bv;
      red_idx[warp] = bi;
      red_inv[warp] = bv ? (b.inv_table ? b.inv_table[bv] : modinv_u32(bv, P)) : 0u;
    }
    __syncthreads();  // S1: also orders the previous pivot's phase C before the row copies below
    if (warp == 0) {
      uint32_t biv = 0;
      bv = lane < NWARPS ? red_val[lane] : 0u;
      bi = lane < NWARPS ? red_idx[lane] : 0x7fffffff;
      biv = lane < NWARPS ? red_inv[lane] : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const uint32_t ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const uint32_t oiv = __shfl_xor_sync(0xffffffffu, biv, o);
        if (better(ov, oi, bv, bi)) {
          bv = ov;
          bi = oi;
          biv = oiv;
        }
      }
      if (lane == 0) {
        cand_val[par] = bv;
        cand_idx[par] = bi;
        cand_inv[par] = biv;
      }
      // publish this CTA's candidate row (all w columns): ONE cluster barrier per pivot suffices
      if (bv != 0 && lane < w) rowbuf_p[par][lane] = panel[lane * rows_c + (bi - my_lo)];
    } else if (warp == 1) {
      if (r >= my_lo && r < my_lo + my_n && lane < w) rowbuf_r[par][lane] = panel[lane * rows_c + (r - my_lo)];
    }
    PANEL_TICK(1)
    cluster.sync();
    PANEL_TICK(2)
    // ---- phase B: ONE round of DSMEM reads, all in flight together: warp c fetches candidate + row of CTA c, one more
    // warp fetches row r from its owner
    {
      const int owner_r = (r - rb) / rows_c;
      for (int c = warp; c <= cs; c += NWARPS) {
        if (c < cs) {
          if (lane == 0) {
            g_val[c] = *cluster.map_shared_rank(&cand_val[par], c);
            g_idx[c] = *cluster.map_shared_rank(&cand_idx[par], c);
            g_inv[c] = *cluster.map_shared_rank(&cand_inv[par], c);
          }
          if (lane < w) allrows[c][lane] = *cluster.map_shared_rank(&rowbuf_p[par][lane], c);
        } else if (lane < w) {
          oldr_s[lane] = *cluster.map_shared_rank(&rowbuf_r[par][lane], owner_r);
        }
      }
    }
    __syncthreads();  // S2
    // every warp selects the winner redundantly from the gathered candidates (cheap, avoids a broadcast)
    uint32_t pv = lane < cs ? g_val[lane] : 0u, pinv = lane < cs ? g_inv[lane] : 0u;
    int p = lane < cs ? g_idx[lane] : 0x7fffffff, src = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint32_t ov = __shfl_xor_sync(0xffffffffu, pv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, p, o);
      const uint32_t oiv = __shfl_xor_sync(0xffffffffu, pinv, o);
      const int os = __shfl_xor_sync(0xffffffffu, src, o);
      if (better(ov, oi, pv, p)) {
        pv = ov;
        p = oi;
        pinv = oiv;
        src = os;
      }
    }
    PANEL_TICK(3)
    if (pv == 0) continue;  // no pivot in this column (uniform over the cluster): skip it, row stays
    if (warp == 0) {
      if (lane < w) {
        const uint32_t xp = allrows[src][lane];
        // new row r = old row p: multipliers of earlier pivots (columns < jj) move with the row, pivot -> 1, rest scaled
        u_s[lane] = lane > jj ? panel_mulmod<SMALL>(xp, pinv, mp, mu32) : (lane == jj ? 1u % P : xp);
      }
      if (lane == 0) {
        const int s_ord = r - rb;
        colpiv[jj] = s_ord;
        pivval[s_ord] = pv;
        if (rank == 0) {
          b.pivcol[r] = j0 + jj;
          b.pinv[r] = pinv;
          b.swp[r] = p;
        }
      }
    }
    __syncthreads();  // S3
    PANEL_TICK(4)
    // ---- phase C: swap rows r <-> p, scale, eliminate.  Every thread owns fixed rows; columns go in chunks of 4 whose
    // loads are issued before their stores (the compiler cannot reorder shared loads across stores of the same array)
    for (int q = tid; q < my_n; q += PANEL_THREADS) {
      const int i = my_lo + q;
      if (i < r) continue;
      uint32_t* prow = panel + q;
      if (i == r) {
        for (int c = 0; c < w; ++c) prow[c * rows_c] = u_s[c];
        continue;
      }
      if (i == p)  // position p receives the old row r (its multipliers of earlier pivots included), then is eliminated like any row
        for (int c = 0; c < w; ++c) prow[c * rows_c] = oldr_s[c];
      const uint32_t l = prow[jj * rows_c];  // stays in place: becomes L[i][t] at store-back
      if (l != 0) {
        const uint32_t nl = P - l;
        for (int c0 = jj + 1; c0 < w; c0 += 4) {
          uint32_t* pc = prow + c0 * rows_c;
          uint32_t a[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c0 + k < w) a[k] = pc[k * rows_c];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c0 + k < w) pc[k * rows_c] = panel_fmamod<SMALL>(a[k], nl, u_s[c0 + k], mp, mu32);
        }
      }
    }
    ++r;
    PANEL_TICK(5)
  }
  __syncthreads();
  // ---- store-back: pivot columns split into U (rows <= pivot row) and L (rows below), everything else is W
  for (int c = 0; c < w; ++c) {
    const int s_ord = colpiv[c];
    const int t = rb + s_ord;
    for (int q = tid; q < my_n; q += PANEL_THREADS) {
      const int i = my_lo + q;
      const uint32_t v = panel[c * rows_c + q];
      uint32_t* wdst = W + (int64_t)(j0 + c) * ldw + i;
      if (s_ord < 0 || i < t) {
        *wdst = v;
      } else if (i == t) {
        *wdst = v;  // == 1
        Lm[(int64_t)t * ldl + i] = pivval[s_ord];
      } else {
        *wdst = 0;
        Lm[(int64_t)t * ldl + i] = v;
      }
    }
  }

  __syncthreads();
  PANEL_TICK(6)
  // ---- compose this panel's transpositions (rb..r-1) into one gather list (warp 0 of CTA 0)
  if (rank == 0 && warp == 0) compose_gather_list(b, rb, r, m, lane);
  cluster.sync();  // all CTAs read st->r long ago; order the final write after every CTA's last use
  if (rank == 0 && tid == 0) b.st->r = r;
}

// ---------------------------------------------------------------------------------------------------
// register-resident panel kernel (default when a CTA's row slice fits 2 rows per thread).  Same cluster layout and
// same results as pluq_panel_kernel, different data flow:
//   * every thread keeps its (up to) two panel rows in registers -- the pivot loop is fully unrolled so all column
//     indices are static; the rank-1 update is pure register arithmetic (u is read from shared memory as a broadcast);
//   * the argmax needs no scan: the thread-local candidates of column jj+1 are produced while column jj is being
//     eliminated (column jj+1 is updated first), and the candidate's inverse is a table load issued at that point and
//     consumed only after the warp reduction of the next pivot step;
//   * candidates are PUSHED: every CTA scales its own candidate row by its own inverse and stores row + (value, index,
//     inverse) into the shared memory of all CTAs before the cluster barrier, so after the barrier the winner, its
//     normalised row u and the displaced row r are local -- one cluster barrier and two block barriers per pivot.
// ---------------------------------------------------------------------------------------------------
// ---- cluster push primitives: asynchronous remote shared-memory stores that credit their bytes to an mbarrier of the
// DESTINATION CTA (st.async ... mbarrier::complete_tx::bytes).  The consumer waits on its own mbarrier for the expected
// byte count: no fence, no cluster-wide barrier, and a CTA proceeds as soon as ITS data has landed.
__device__ __forceinline__ uint32_t psm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t psm_remote(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(psm_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint4 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void pmbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(psm_u32(bar)), "r"(count));
}
__device__ __forceinline__ void pmbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(psm_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pmbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(psm_u32(bar)), "r"(parity)
        : "memory");
  }
}

struct PanelAr {
  uint32_t P, mu, sh;  // ARITH 0: mu = floor(2^32 / P); ARITH 1: mu = floor(2^(32+sh) / P), sh = bits(P) - 1
  uint32_t mu2, sh2;   // left-looking kernel, any P < 2^29: sh2 = bits(P) - 1, mu2 = min(floor(2^(32+sh2) / P), 2^32 - 1)
  ModP mp;
};

// x < 32 * P^2 -> x mod P.  For P < 2^26 the top 32 bits of x give a quotient estimate that is at most 4 short, so the
// remainder fits 32 bits and three conditional subtractions finish the job; larger P take the generic 64-bit Barrett.
__device__ __forceinline__ uint32_t ll_reduce(uint64_t x, const PanelAr& ar) {
  if (ar.sh2 <= 25) {
    const uint32_t xh = (uint32_t)(x >> ar.sh2);
    uint32_t r = (uint32_t)x - __umulhi(xh, ar.mu2) * ar.P;
    r = min(r, r - ar.P);
    r = min(r, r - ar.P);
    return min(r, r - ar.P);
  }
  return (uint32_t)mod_u64(x, ar.mp);
}

// ARITH 0: P <= 2^16 (32-bit Barrett);  1: 2^16 < P < 2^30 (quotient estimate from the top 32 bits);  2: generic 64-bit
// Barrett;  3: P < 2^16 LAZY -- eliminated values stay in [0, 2P) (x = nl*u + a <= (P-1)^2 + 2P - 1 < 2^32), only the next
// pivot column, published rows and the final store are brought back to [0, P);  4: 2^16 < P < 2^30 LAZY -- values stay in
// [0, 2.6 P): with x < P^2 + 2.6 P the quotient estimate is still at most 2.6 short.
template <int ARITH>
__device__ __forceinline__ uint32_t pa_reduce(uint64_t x, const PanelAr& ar) {  // x < P^2 (ARITH 3: <= P^2), result < P
  if constexpr (ARITH == 0 || ARITH == 3) {
    const uint32_t x32 = (uint32_t)x;
    const uint32_t r = x32 - __umulhi(x32, ar.mu) * ar.P;  // in [0, 2P)
    return min(r, r - ar.P);
  } else if constexpr (ARITH == 1 || ARITH == 4) {
    const uint32_t xh = (uint32_t)(x >> ar.sh);
    uint32_t r = (uint32_t)x - __umulhi(xh, ar.mu) * ar.P;  // in [0, 2.5 P)
    r = min(r, r - ar.P);
    return min(r, r - ar.P);
  } else {
    return (uint32_t)mod_u64(x, ar.mp);
  }
}
template <int ARITH>
__device__ __forceinline__ uint32_t pa_canon(uint32_t x, const PanelAr& ar) {
  if constexpr (ARITH == 3) return min(x, x - ar.P);
  if constexpr (ARITH == 4) {
    x = min(x, x - ar.P);
    return min(x, x - ar.P);
  }
  return x;
}
// (a + nl * u) mod P, nl, u < P; ARITH 3: a < 2P and the result is only < 2P
template <int ARITH>
__device__ __forceinline__ uint32_t pa_fmamod(uint32_t a, uint32_t nl, uint32_t u, const PanelAr& ar) {
  if constexpr (ARITH == 3) {
    const uint32_t x32 = nl * u + a;
    return x32 - __umulhi(x32, ar.mu) * ar.P;
  } else if constexpr (ARITH == 4) {
    const uint64_t x = (uint64_t)nl * u + a;
    return (uint32_t)x - __umulhi((uint32_t)(x >> ar.sh), ar.mu) * ar.P;
  } else if constexpr (ARITH == 0) {
    return pa_reduce<0>((uint64_t)(nl * u + a), ar);
  } else {
    return pa_reduce<ARITH>((uint64_t)nl * u + a, ar);
  }
}
template <int ARITH>
__device__ __forceinline__ uint32_t pa_mulmod(uint32_t a, uint32_t c, const PanelAr& ar) {  // a, c < P -> result < P
  if constexpr (ARITH == 0 || ARITH == 3) return pa_reduce<ARITH>((uint64_t)(a * c), ar);
  return pa_reduce<ARITH>((uint64_t)a * c, ar);
}

constexpr int PREG_ROWS = 1024;  // rows per CTA: PREG_RPT rows per thread, PREG_ROWS / PREG_RPT threads
constexpr int PREG_MAXC = 16;    // largest cluster

template <int ARITH, int PW, int PREG_RPT>
__global__ void __launch_bounds__(PREG_ROWS / PREG_RPT, 1)
pluq_panel_reg_kernel(uint32_t* __restrict__ Wm, int64_t ldw, int m, int j0, int w, uint32_t* __restrict__ Lm, int64_t ldl, PluqBufs b,
                      const __grid_constant__ PanelAr ar) {
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int PREG_T = PREG_ROWS / PREG_RPT;
  constexpr int NW = PREG_T / 32;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ uint32_t red_val[NW], red_inv[NW];
  __shared__ int red_idx[NW];
  __shared__ __align__(16) uint32_t rowbuf_w[NW][PW], rowbuf_r[PW];  // every warp publishes its own candidate row: no second block barrier
  // filled by the asynchronous remote stores of every CTA (double-buffered by pivot parity); xbar[par] counts their bytes
  __shared__ __align__(16) uint4 g_cand[2][PREG_MAXC];  // (value, row, inverse, -)
  __shared__ __align__(16) uint32_t allrows[2][PREG_MAXC][PW];
  __shared__ __align__(16) uint32_t oldr[2][PW];
  __shared__ __align__(8) uint64_t xbar[2];
  __shared__ int colpiv[PW];       // panel column -> pivot ordinal within this panel, or -1
  __shared__ uint32_t pivval[PW];  // pivot value of ordinal s
  __shared__ uint32_t pinv_s[PW];  // per ordinal: inverse, swap partner, column -- written to global memory once, at the end
  __shared__ int swp_s[PW], pcol_s[PW];

  const int rb = b.st->r;
  const int rows_total = m - rb;
  const int rows_c = (rows_total + cs - 1) / cs;
  const int my_lo = rb + rank * rows_c;
  int my_n = rows_total - rank * rows_c;
  my_n = my_n < 0 ? 0 : (my_n > rows_c ? rows_c : my_n);
  const uint32_t P = ar.P;
  // optional cycle counters (GFFM_PANEL_PROF=1): accumulated in shared memory by thread 0 of CTA 0, flushed once at the end
  __shared__ long long prof_s[8];
  const bool do_prof = b.prof != nullptr && rank == 0 && tid == 0;
  if (do_prof)
    for (int i = 0; i < 8; ++i) prof_s[i] = 0;
  long long tprev = do_prof ? clock64() : 0;
#define PREG_TICK(slot)                          \
  if (do_prof) {                                 \
    const long long tn = clock64();              \
    prof_s[slot] += tn - tprev;                  \
    tprev = clock64();                           \
  }

  uint32_t a[PREG_RPT][PW];
  int gi[PREG_RPT];  // global row of slot k, or -1
#pragma unroll
  for (int k = 0; k < PREG_RPT; ++k) {
    const int q = tid + PREG_T * k;
    gi[k] = q < my_n ? my_lo + q : -1;
    const uint32_t* src = Wm + (int64_t)j0 * ldw + (gi[k] >= 0 ? gi[k] : 0);
#pragma unroll
    for (int c = 0; c < PW; ++c) {
      a[k][c] = (gi[k] >= 0 && c < w) ? *src : 0u;
      src += ldw;
    }
  }
  if (tid < PW) colpiv[tid] = -1;
  if (tid == 0) {
    pmbar_init(&xbar[0], 1);
    pmbar_init(&xbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t xbytes = (uint32_t)cs * (PW * 4 + 16) + PW * 4;  // per pivot: every CTA's row + candidate, and row r once
  cluster.sync();  // everybody has read st->r before anyone can finish and overwrite it; shared arrays + mbarriers exist cluster-wide
  PREG_TICK(0)

  int r = rb;
  // The panel is kept ROTATED: at pivot step jj register position pos holds logical column (jj + pos) mod PW, so the
  // pivot column is always position 0 and every register index below is static although jj is a run-time loop counter.
  // The rotation by one position is folded into the elimination (results are written one position to the left).
  // thread-local candidate of the current column: (value, row) and -- table path -- its inverse (load in flight)
  uint32_t lv = 0, linv = 0;
  int li = 0x7fffffff;
  auto local_cand = [&](int rmin) {
    lv = 0;
    li = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < PREG_RPT; ++k) {
      if (gi[k] >= rmin && better(a[k][0], gi[k], lv, li)) {
        lv = a[k][0];
        li = gi[k];
      }
    }
    linv = (lv && b.inv_table) ? b.inv_table[lv] : 0u;
  };
  // copy the row held in slot `hi` to shared memory: one branch per slot keeps the register indices static; 128-bit stores
  auto store_row = [&](uint32_t* dst, bool hi) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    if (PREG_RPT == 2 && hi) {
#pragma unroll
      for (int c = 0; c < PW; c += 4) d4[c >> 2] = make_uint4(a[PREG_RPT - 1][c], a[PREG_RPT - 1][c + 1], a[PREG_RPT - 1][c + 2], a[PREG_RPT - 1][c + 3]);
    } else {
#pragma unroll
      for (int c = 0; c < PW; c += 4) d4[c >> 2] = make_uint4(a[0][c], a[0][c + 1], a[0][c + 2], a[0][c + 3]);
    }
  };
  local_cand(r);

  int jj = 0;
  for (; jj < w && r < m; ++jj) {
    const int par = jj & 1;
    if (tid == 0) pmbar_expect_tx(&xbar[par], xbytes);
    // the owner of row r publishes it (raw) before anything modifies it
    const int qr = r - my_lo;
    const bool owns_r = qr >= 0 && qr < my_n;
    if (owns_r && ((qr & (PREG_T - 1)) >> 5) == warp) {  // warp-uniform: 15 of 16 warps skip the copy code entirely
      if ((qr & 31) == lane) store_row(rowbuf_r, qr >= PREG_T);
    }
    // ---- warp argmax (max residue, first index) on the REDUX unit; the winning lane publishes value, row index, inverse and
    // its whole row, so the block-level decision below needs no further barrier
    {
      const uint32_t wv = __reduce_max_sync(FULL, lv);
      const int wi = __reduce_min_sync(FULL, lv == wv ? li : 0x7fffffff);
      if (wv == 0) {
        if (lane == 0) red_val[warp] = 0u;
      } else if (lv == wv && li == wi) {
        red_val[warp] = wv;
        red_idx[warp] = wi;
        red_inv[warp] = b.inv_table ? linv : modinv_u32(wv, P);
        store_row(rowbuf_w[warp], wi == gi[PREG_RPT - 1]);
      }
    }
    __syncthreads();  // S1
    PREG_TICK(1)
    // ---- block argmax, redundantly in every warp
    uint32_t bv, biv = 0;
    int bi, ww;
    {
      const uint32_t v = lane < NW ? red_val[lane] : 0u;
      const int i = (lane < NW && v) ? red_idx[lane] : 0x7fffffff;
      bv = __reduce_max_sync(FULL, v);
      bi = __reduce_min_sync(FULL, v == bv ? i : 0x7fffffff);
      const unsigned hit = __ballot_sync(FULL, v == bv && i == bi);
      ww = hit ? __ffs(hit) - 1 : 0;
      if (bv) biv = red_inv[ww];
    }
    PREG_TICK(2)
    // ---- push: warp d sends this CTA's candidate (normalised row, value, index, inverse) and, if owned, row r to CTA d.
    // Lanes 0..PW/4-1 carry four columns each (one 16-byte st.async), the next lane the candidate record.
    if (warp < cs) {
      const uint32_t rbar = psm_remote(&xbar[par], warp);
      if (lane < PW / 4) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (bv != 0) {
          const uint4 x4 = *reinterpret_cast<const uint4*>(&rowbuf_w[ww][4 * lane]);
          const uint32_t x[4] = {x4.x, x4.y, x4.z, x4.w};
          uint32_t y[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int pos = 4 * lane + e;
            const uint32_t xc = pa_canon<ARITH>(x[e], ar);
            // pivot -> 1, columns right of it (and the zero padding) scaled by the inverse; the wrapped-around positions
            // >= PW - jj hold the multipliers of earlier pivots, which move with the row unchanged
            y[e] = pos == 0 ? 1u % P : (pos < PW - jj ? pa_mulmod<ARITH>(xc, biv, ar) : xc);
          }
          v = make_uint4(y[0], y[1], y[2], y[3]);
        }
        st_async_v4(psm_remote(&allrows[par][rank][4 * lane], warp), v, rbar);
        if (owns_r) {
          const uint4 o4 = *reinterpret_cast<const uint4*>(&rowbuf_r[4 * lane]);
          st_async_v4(psm_remote(&oldr[par][4 * lane], warp),
                      make_uint4(pa_canon<ARITH>(o4.x, ar), pa_canon<ARITH>(o4.y, ar), pa_canon<ARITH>(o4.z, ar), pa_canon<ARITH>(o4.w, ar)), rbar);
        }
      } else if (lane == PW / 4) {
        st_async_v4(psm_remote(&g_cand[par][rank], warp), make_uint4(bv, (uint32_t)bi, biv, 0u), rbar);
      }
    }
    PREG_TICK(3)
    pmbar_wait_cluster(&xbar[par], (uint32_t)(jj >> 1) & 1u);
    PREG_TICK(4)
    // ---- global winner, selected redundantly by every warp from local shared memory
    uint32_t pv;
    int p, src;
    {
      const uint32_t v = lane < cs ? g_cand[par][lane].x : 0u;
      const int i = (lane < cs && v) ? (int)g_cand[par][lane].y : 0x7fffffff;
      pv = __reduce_max_sync(FULL, v);
      p = __reduce_min_sync(FULL, v == pv ? i : 0x7fffffff);
      const unsigned hit = __ballot_sync(FULL, v == pv && i == p);
      src = hit ? __ffs(hit) - 1 : 0;
    }
    const bool have = pv != 0;  // false: no pivot in this column (uniform over the cluster): skip it, row r stays
    if (have && tid == 0) {
      const int s_ord = r - rb;
      colpiv[jj] = s_ord;
      pivval[s_ord] = pv;
      pinv_s[s_ord] = g_cand[par][src].z;
      swp_s[s_ord] = p;
      pcol_s[s_ord] = j0 + jj;
    }
    PREG_TICK(5)
    // ---- swap rows r <-> p, eliminate and rotate (registers only)
    const uint32_t* u = allrows[par][src];
    uint32_t nl[PREG_RPT], keep[PREG_RPT];
#pragma unroll
    for (int k = 0; k < PREG_RPT; ++k) {
      // rows r and p are replaced; the test is made warp-uniform first so that the 2 x 32 copies are real branches that
      // (almost) every warp skips instead of predicated code executed by all
      if (have && __any_sync(FULL, gi[k] == r || gi[k] == p)) {
        if (gi[k] == r) {
#pragma unroll
          for (int c = 0; c < PW; ++c) a[k][c] = u[c];
        } else if (gi[k] == p) {  // position p receives the old row r (its multipliers of earlier pivots included)
#pragma unroll
          for (int c = 0; c < PW; ++c) a[k][c] = oldr[par][c];
        }
      }
      const uint32_t l = a[k][0];  // below the pivot: becomes L[i][t] at store-back
      nl[k] = (have && gi[k] > r && l) ? P - l : 0u;
      keep[k] = l;
    }
    const int nact = have ? PW - 1 - jj : 0;  // positions 1..nact are the columns right of the pivot
    const int rnext = have ? r + 1 : r;
    // position 1 first: it is the next pivot column, its candidates (and the inverse-table load) start right away
    {
      const uint32_t u1 = u[1];
#pragma unroll
      for (int k = 0; k < PREG_RPT; ++k) a[k][0] = pa_canon<ARITH>(nact >= 1 ? pa_fmamod<ARITH>(a[k][1], nl[k], u1, ar) : a[k][1], ar);
      local_cand(rnext);
    }
#pragma unroll
    for (int g = 2; g < PW; g += 4) {
      if (g <= nact) {  // uniform: groups of four positions that (mostly) need the update
#pragma unroll
        for (int pos = g; pos < g + 4 && pos < PW; ++pos) {
          const uint32_t uc = u[pos];
#pragma unroll
          for (int k = 0; k < PREG_RPT; ++k) a[k][pos - 1] = pos <= nact ? pa_fmamod<ARITH>(a[k][pos], nl[k], uc, ar) : a[k][pos];
        }
      } else {
#pragma unroll
        for (int pos = g; pos < g + 4 && pos < PW; ++pos)
#pragma unroll
          for (int k = 0; k < PREG_RPT; ++k) a[k][pos - 1] = a[k][pos];
      }
    }
#pragma unroll
    for (int k = 0; k < PREG_RPT; ++k) a[k][PW - 1] = keep[k];
    r = rnext;
    PREG_TICK(6)
  }
  // undo the remaining rotation so that position == logical column again
  for (; jj < PW; ++jj) {
#pragma unroll
    for (int k = 0; k < PREG_RPT; ++k) {
      const uint32_t t0 = a[k][0];
#pragma unroll
      for (int pos = 1; pos < PW; ++pos) a[k][pos - 1] = a[k][pos];
      a[k][PW - 1] = t0;
    }
  }
  __syncthreads();
  // ---- store-back: pivot columns split into U (rows <= pivot row) and L (rows below), everything else is W
  {
    uint32_t* wcol = Wm + (int64_t)j0 * ldw;
#pragma unroll
    for (int c = 0; c < PW; ++c) {
      if (c < w) {
        const int s_ord = colpiv[c];  // uniform
        if (s_ord < 0) {
#pragma unroll
          for (int k = 0; k < PREG_RPT; ++k)
            if (gi[k] >= 0) wcol[gi[k]] = pa_canon<ARITH>(a[k][c], ar);
        } else {
          const int t = rb + s_ord;
          uint32_t* lcol = Lm + (int64_t)t * ldl;
          const uint32_t pvv = pivval[s_ord];
#pragma unroll
          for (int k = 0; k < PREG_RPT; ++k) {
            const int i = gi[k];
            if (i >= 0) {
              const uint32_t v = pa_canon<ARITH>(a[k][c], ar);
              wcol[i] = i <= t ? v : 0u;  // the pivot itself is stored as 1
              if (i >= t) lcol[i] = i == t ? pvv : v;
            }
          }
        }
      }
      wcol += ldw;
    }
  }
  if (rank == 0 && tid < r - rb) {
    b.pivcol[rb + tid] = pcol_s[tid];
    b.pinv[rb + tid] = pinv_s[tid];
    b.swp[rb + tid] = swp_s[tid];
  }
  __syncthreads();
  PREG_TICK(7)
  if (do_prof)
    for (int i = 0; i < 8; ++i) b.prof[i] += prof_s[i];
  if (rank == 0 && warp == 0) compose_gather_list(b, rb, r, m, lane);
  cluster.sync();  // all CTAs read st->r long ago; order the final write after every CTA's last use
  if (rank == 0 && tid == 0) b.st->r = r;
}

// ---------------------------------------------------------------------------------------------------
// LEFT-LOOKING register panel kernel (default for P < 2^29).  Same layout and exchange protocol as
// pluq_panel_reg_kernel, but a pivot no longer updates all remaining panel columns (3 multiplies per element and
// pivot on the half-rate integer pipe).  Instead a column is brought up to date only when it becomes the pivot column:
//   a[i][jj] -= sum_{t < jj} l[i][t] * U[t][jj]      one 32x32+64 multiply-add per term, ONE reduction per row and pivot
// with the multipliers l[i][t] read from the row's own registers and U (the normalised pivot rows of this panel) from a
// PW x PW table in shared memory.  The candidate row of every CTA gets the same treatment for ALL its remaining
// columns by the pushing warps (lane = column), so the pushed row is the final normalised pivot row.
// Registers are rotated by four positions after every fourth pivot; inside a group of four the pivot position is static.
// ---------------------------------------------------------------------------------------------------
template <int PW, bool PROF>
__global__ void __launch_bounds__(512, 1)
pluq_panel_ll_kernel(uint32_t* __restrict__ Wm, int64_t ldw, int m, int j0, int w, uint32_t* __restrict__ Lm, int64_t ldl, PluqBufs b,
                     const __grid_constant__ PanelAr ar) {
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int T = 512, NW = T / 32, RPT = 2;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ uint32_t red_val[NW], red_inv[NW];
  __shared__ int red_idx[NW];
  __shared__ __align__(16) uint32_t rowbuf_w[NW][PW], rowbuf_r[PW];
  __shared__ __align__(16) uint4 g_cand[2][PREG_MAXC];  // (value, row, inverse, -)
  __shared__ __align__(16) uint32_t allrows[2][PREG_MAXC][PW];
  __shared__ __align__(16) uint32_t oldr[2][PW];
  __shared__ __align__(8) uint64_t xbar[2];
  __shared__ uint32_t Utab[PW][PW];  // [pivot column t][logical column c]: normalised pivot rows (zero row for a skipped column)
  __shared__ int colpiv[PW];
  __shared__ uint32_t pivval[PW], pinv_s[PW];
  __shared__ int swp_s[PW], pcol_s[PW];

  const int rb = b.st->r;
  const int rows_total = m - rb;
  const int rows_c = (rows_total + cs - 1) / cs;
  const int my_lo = rb + rank * rows_c;
  int my_n = rows_total - rank * rows_c;
  my_n = my_n < 0 ? 0 : (my_n > rows_c ? rows_c : my_n);
  const uint32_t P = ar.P;
  __shared__ long long prof_s[8];
  const bool do_prof = PROF && b.prof != nullptr && rank == 0 && tid == 0;
  if (do_prof)
    for (int i = 0; i < 8; ++i) prof_s[i] = 0;
  long long tprev = do_prof ? clock64() : 0;
#define PLL_TICK(slot)                           \
  if (PROF && do_prof) {                         \
    const long long tn = clock64();              \
    prof_s[slot] += tn - tprev;                  \
    tprev = clock64();                           \
  }

  uint32_t a[RPT][PW];
  int gi[RPT];
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int q = tid + T * k;
    gi[k] = q < my_n ? my_lo + q : -1;
    const uint32_t* src = Wm + (int64_t)j0 * ldw + (gi[k] >= 0 ? gi[k] : 0);
#pragma unroll
    for (int c = 0; c < PW; ++c) {
      a[k][c] = (gi[k] >= 0 && c < w) ? *src : 0u;
      src += ldw;
    }
  }
  if (tid < PW) colpiv[tid] = -1;
  if (tid == 0) {
    pmbar_init(&xbar[0], 1);
    pmbar_init(&xbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t xbytes = (uint32_t)cs * (PW * 4 + 16) + PW * 4;
  cluster.sync();
  PLL_TICK(0)

  int r = rb;
  uint32_t lv = 0, linv = 0;
  int li = 0x7fffffff;
  auto store_row = [&](uint32_t* dst, bool hi) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    if (hi) {
#pragma unroll
      for (int c = 0; c < PW; c += 4) d4[c >> 2] = make_uint4(a[1][c], a[1][c + 1], a[1][c + 2], a[1][c + 3]);
    } else {
#pragma unroll
      for (int c = 0; c < PW; c += 4) d4[c >> 2] = make_uint4(a[0][c], a[0][c + 1], a[0][c + 2], a[0][c + 3]);
    }
  };
  // candidates of the column at register position 0..3 (static S)
#define PLL_LOCAL_CAND(S, RMIN)                                             \
  {                                                                         \
    lv = 0;                                                                 \
    li = 0x7fffffff;                                                        \
    if (gi[0] >= (RMIN) && better(a[0][S], gi[0], lv, li)) {                \
      lv = a[0][S];                                                         \
      li = gi[0];                                                           \
    }                                                                       \
    if (gi[1] >= (RMIN) && better(a[1][S], gi[1], lv, li)) {                \
      lv = a[1][S];                                                         \
      li = gi[1];                                                           \
    }                                                                       \
    linv = (lv && b.inv_table) ? b.inv_table[lv] : 0u;                      \
  }
  PLL_LOCAL_CAND(0, r)

  int J4 = 0;  // 4 * (group index): register position pos holds logical column (J4 + pos) mod PW
  bool done = false;
  for (; J4 < PW && !done; J4 += 4) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int jj = J4 + s;
      if (jj >= w || r >= m) {
        done = true;
        break;
      }
      const int par = jj & 1;
      if (tid == 0) pmbar_expect_tx(&xbar[par], xbytes);
      const int qr = r - my_lo;
      const bool owns_r = qr >= 0 && qr < my_n;
      if (owns_r && ((qr & (T - 1)) >> 5) == warp) {
        if ((qr & 31) == lane) store_row(rowbuf_r, qr >= T);
      }
      // ---- warp argmax; the winning lane publishes value, row index, inverse and its whole row
      {
        const uint32_t wv = __reduce_max_sync(FULL, lv);
        const int wi = __reduce_min_sync(FULL, lv == wv ? li : 0x7fffffff);
        if (wv == 0) {
          if (lane == 0) red_val[warp] = 0u;
        } else if (lv == wv && li == wi) {
          red_val[warp] = wv;
          red_idx[warp] = wi;
          red_inv[warp] = b.inv_table ? linv : modinv_u32(wv, P);
          store_row(rowbuf_w[warp], wi == gi[1]);
        }
      }
      __syncthreads();  // S1
      PLL_TICK(1)
      uint32_t bv, biv = 0;
      int bi, ww;
      {
        const uint32_t v = lane < NW ? red_val[lane] : 0u;
        const int i = (lane < NW && v) ? red_idx[lane] : 0x7fffffff;
        bv = __reduce_max_sync(FULL, v);
        bi = __reduce_min_sync(FULL, v == bv ? i : 0x7fffffff);
        const unsigned hit = __ballot_sync(FULL, v == bv && i == bi);
        ww = hit ? __ffs(hit) - 1 : 0;
        if (bv) biv = red_inv[ww];
      }
      PLL_TICK(2)
      // ---- push: lane = register position.  Columns right of the pivot are first brought up to date with the jj earlier
      // pivots of this panel (dot product of the row's multipliers with a column of Utab), then scaled by the inverse.
      if (warp < cs) {
        const uint32_t rbar = psm_remote(&xbar[par], warp);
        if (lane < PW) {
          uint32_t v = 0;
          if (bv != 0) {
            const uint32_t* rw = rowbuf_w[ww];
            const uint32_t x = rw[lane];
            if (lane == s) {
              v = 1u % P;
            } else if (lane > s && lane < PW - J4) {  // a column right of the pivot (or zero padding)
              const int col = J4 + lane;
              uint64_t acc0 = 0, acc1 = 0;  // two chains, loads of four terms in flight together
              for (int e = 0; e < J4; e += 4) {  // older groups (wrapped positions)
                const uint32_t l0 = rw[PW - 1 - e], l1 = rw[PW - 2 - e], l2 = rw[PW - 3 - e], l3 = rw[PW - 4 - e];
                const uint32_t u0 = Utab[J4 - 1 - e][col], u1 = Utab[J4 - 2 - e][col], u2 = Utab[J4 - 3 - e][col], u3 = Utab[J4 - 4 - e][col];
                acc0 += (uint64_t)l0 * u0;
                acc1 += (uint64_t)l1 * u1;
                acc0 += (uint64_t)l2 * u2;
                acc1 += (uint64_t)l3 * u3;
              }
#pragma unroll
              for (int sp = 0; sp < s; ++sp) acc0 += (uint64_t)rw[sp] * Utab[J4 + sp][col];  // this group
              const uint32_t am = ll_reduce(acc0 + acc1, ar);
              const uint32_t xu = x >= am ? x - am : x + P - am;
              v = ll_reduce((uint64_t)xu * biv, ar);
            } else {
              v = x;  // multipliers of earlier pivots move with the row unchanged
            }
          }
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(
                           psm_remote(&allrows[par][rank][lane], warp)),
                       "r"(v), "r"(rbar)
                       : "memory");
          if (owns_r)
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(
                             psm_remote(&oldr[par][lane], warp)),
                         "r"(rowbuf_r[lane]), "r"(rbar)
                         : "memory");
        }
        if (lane == 0) st_async_v4(psm_remote(&g_cand[par][rank], warp), make_uint4(bv, (uint32_t)bi, biv, 0u), rbar);
      }
      PLL_TICK(3)
      pmbar_wait_cluster(&xbar[par], (uint32_t)(jj >> 1) & 1u);
      PLL_TICK(4)
      uint32_t pv;
      int p, src;
      {
        const uint32_t v = lane < cs ? g_cand[par][lane].x : 0u;
        const int i = (lane < cs && v) ? (int)g_cand[par][lane].y : 0x7fffffff;
        pv = __reduce_max_sync(FULL, v);
        p = __reduce_min_sync(FULL, v == pv ? i : 0x7fffffff);
        const unsigned hit = __ballot_sync(FULL, v == pv && i == p);
        src = hit ? __ffs(hit) - 1 : 0;
      }
      const bool have = pv != 0;
      const uint32_t* u = allrows[par][src];
      // every warp records the new pivot row in the U table (identical values; the warp reads back only its own writes
      // before the next block barrier)
      if (lane < PW) Utab[jj][(J4 + lane) & (PW - 1)] = have ? u[lane] : 0u;
      __syncwarp();
      if (have && tid == 0) {
        const int s_ord = r - rb;
        colpiv[jj] = s_ord;
        pivval[s_ord] = pv;
        pinv_s[s_ord] = g_cand[par][src].z;
        swp_s[s_ord] = p;
        pcol_s[s_ord] = j0 + jj;
      }
      PLL_TICK(5)
      // ---- swap rows r <-> p (register copies, skipped by every warp that owns neither)
      if (have && __any_sync(FULL, gi[0] == r || gi[0] == p || gi[1] == r || gi[1] == p)) {
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          if (gi[k] == r) {
#pragma unroll
            for (int c = 0; c < PW; ++c) a[k][c] = u[c];
          } else if (gi[k] == p) {
#pragma unroll
            for (int c = 0; c < PW; ++c) a[k][c] = oldr[par][c];
          }
        }
      }
      const int rnext = have ? r + 1 : r;
      // ---- bring the NEXT pivot column (register position s + 1) up to date for the rows below the pivot, then its candidates
      if (jj + 1 < PW) {
        const int col = jj + 1;
        uint64_t acc[RPT][2] = {{0, 0}, {0, 0}};  // two chains per row
#pragma unroll
        for (int e0 = 0; e0 < PW - 4; e0 += 4) {
          if (e0 < J4) {  // uniform: one older group of four pivots
            uint32_t uv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) uv[e] = Utab[J4 - 1 - e0 - e][col];
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
              for (int k = 0; k < RPT; ++k) acc[k][e & 1] += (uint64_t)a[k][PW - 1 - e0 - e] * uv[e];
          }
        }
#pragma unroll
        for (int sp = 0; sp <= s; ++sp) {
          const uint32_t uv = Utab[J4 + sp][col];
#pragma unroll
          for (int k = 0; k < RPT; ++k) acc[k][sp & 1] += (uint64_t)a[k][sp] * uv;
        }
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          if (gi[k] >= rnext) {
            const uint32_t am = ll_reduce(acc[k][0] + acc[k][1], ar);
            const uint32_t x = a[k][s + 1];
            a[k][s + 1] = x >= am ? x - am : x + P - am;
          }
        }
        if (s < 3) {
          PLL_LOCAL_CAND(s + 1, rnext)
        }
      }
      r = rnext;
      PLL_TICK(6)
    }
    if (done) break;
    // rotate by four: position 4 (the next pivot column, already up to date) becomes position 0
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const uint32_t t0 = a[k][0], t1 = a[k][1], t2 = a[k][2], t3 = a[k][3];
#pragma unroll
      for (int pos = 4; pos < PW; ++pos) a[k][pos - 4] = a[k][pos];
      a[k][PW - 4] = t0;
      a[k][PW - 3] = t1;
      a[k][PW - 2] = t2;
      a[k][PW - 1] = t3;
    }
    PLL_LOCAL_CAND(0, r)
  }
#undef PLL_LOCAL_CAND
  // undo the remaining rotation so that position == logical column again
  for (int q4 = J4 & (PW - 1); q4 != 0 && q4 < PW; q4 += 4) {
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const uint32_t t0 = a[k][0], t1 = a[k][1], t2 = a[k][2], t3 = a[k][3];
#pragma unroll
      for (int pos = 4; pos < PW; ++pos) a[k][pos - 4] = a[k][pos];
      a[k][PW - 4] = t0;
      a[k][PW - 3] = t1;
      a[k][PW - 2] = t2;
      a[k][PW - 1] = t3;
    }
  }
  __syncthreads();
  {
    uint32_t* wcol = Wm + (int64_t)j0 * ldw;
#pragma unroll
    for (int c = 0; c < PW; ++c) {
      if (c < w) {
        const int s_ord = colpiv[c];  // uniform
        if (s_ord < 0) {
#pragma unroll
          for (int k = 0; k < RPT; ++k)
            if (gi[k] >= 0) wcol[gi[k]] = a[k][c];
        } else {
          const int t = rb + s_ord;
          uint32_t* lcol = Lm + (int64_t)t * ldl;
          const uint32_t pvv = pivval[s_ord];
#pragma unroll
          for (int k = 0; k < RPT; ++k) {
            const int i = gi[k];
            if (i >= 0) {
              const uint32_t v = a[k][c];
              wcol[i] = i <= t ? v : 0u;  // the pivot itself is stored as 1
              if (i >= t) lcol[i] = i == t ? pvv : v;
            }
          }
        }
      }
      wcol += ldw;
    }
  }
  if (rank == 0 && tid < r - rb) {
    b.pivcol[rb + tid] = pcol_s[tid];
    b.pinv[rb + tid] = pinv_s[tid];
    b.swp[rb + tid] = swp_s[tid];
  }
  __syncthreads();
  PLL_TICK(7)
  if (do_prof)
    for (int i = 0; i < 8; ++i) b.prof[i] += prof_s[i];
  if (rank == 0 && warp == 0) compose_gather_list(b, rb, r, m, lane);
  cluster.sync();
  if (rank == 0 && tid == 0) b.st->r = r;
#undef PLL_TICK
}

// ---------------------------------------------------------------------------------------------------
// apply the composed gather list of the last panel to columns [c_lo, c_hi) of a matrix (one warp per column)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_rows_kernel(uint32_t* __restrict__ X, int64_t ld, int c_lo, int c_hi_static, const int* __restrict__ c_hi_dyn, PluqBufs b) {
  const int n = b.st->n_gather;
  if (n == 0) return;
  const int c_hi = c_hi_dyn ? *c_hi_dyn : c_hi_static;  // dynamic upper bound = st->rb for the L matrix
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int c = c_lo + wg; c < c_hi; c += nw) {
    uint32_t* col = X + (int64_t)c * ld;
    uint32_t v0 = 0, v1 = 0;
    if (lane < n) v0 = col[b.g_src[lane]];
    if (lane + 32 < n) v1 = col[b.g_src[lane + 32]];
    __syncwarp();
    if (lane < n) col[b.g_dst[lane]] = v0;
    if (lane + 32 < n) col[b.g_dst[lane + 32]] = v1;
  }
}

// ---------------------------------------------------------------------------------------------------
// inner TRSM: for columns c in [c_lo, c_hi): rows rb..r-1 of W  <-  L11'^-1 * W   (one warp per column,
// forward substitution through shuffles; L11' = L[rb..r, rb..r], pivots on the diagonal)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
trsm_small_kernel(uint32_t* __restrict__ W, int64_t ldw, int c_lo, int c_hi, const uint32_t* __restrict__ Lm, int64_t ldl, PluqBufs b,
                  const __grid_constant__ ModP mp) {
  const int rb = b.st->rb, k = b.st->r - rb;
  if (k <= 0) return;
  __shared__ uint32_t L11[32][33];
  __shared__ uint32_t pinv[32];
  for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
    const int i = e & 31, t = e >> 5;
    L11[i][t] = (i < k && t < k && i > t) ? Lm[(int64_t)(rb + t) * ldl + rb + i] : 0u;
  }
  if (threadIdx.x < 32) pinv[threadIdx.x] = threadIdx.x < k ? b.pinv[rb + threadIdx.x] : 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const uint32_t P = (uint32_t)mp.P;
  for (int c = c_lo + wg; c < c_hi; c += nw) {
    uint32_t* col = W + (int64_t)c * ldw + rb;
    uint32_t x = lane < k ? col[lane] : 0u;
    for (int t = 0; t < k; ++t) {
      uint32_t ut = mulmod_u32(__shfl_sync(0xffffffffu, x, t), pinv[t], mp);
      if (lane == t) x = ut;
      else if (lane > t) x = submod_u32(x, mulmod_u32(L11[lane][t], ut, mp), P);
    }
    if (lane < k) col[lane] = x;
  }
}

// ---------------------------------------------------------------------------------------------------
// inner rank-k' update: W[r.., c] -= L[r.., rb..r) * W[rb..r, c] for c in [c_lo, c_hi)  (k' <= 32, SIMT; the
// big trailing update of an outer block goes through the tensor-core GEMM instead)
// tile: 128 rows x 32 columns per CTA of 256 threads
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
update_small_kernel(uint32_t* __restrict__ W, int64_t ldw, int m, int c_lo, int c_hi, const uint32_t* __restrict__ Lm, int64_t ldl,
                    PluqBufs b, int big_mod, const __grid_constant__ ModP mp) {
  const int rb = b.st->rb, r = b.st->r, k = r - rb;
  if (k <= 0) return;
  const int i0 = r + blockIdx.x * 128;
  if (i0 >= m) return;
  __shared__ uint32_t sL[32][128 + 1];  // [t][i]
  __shared__ uint32_t sU[32][32 + 1];   // [t][c]
  const int tid = threadIdx.x;
  for (int e = tid; e < 32 * 128; e += 256) {
    const int i = e & 127, t = e >> 7;
    sL[t][i] = (t < k && i0 + i < m) ? Lm[(int64_t)(rb + t) * ldl + i0 + i] : 0u;
  }
  const int ti = tid & 127, tc = tid >> 7;  // row within tile, column parity
  const uint32_t P = (uint32_t)mp.P;
  for (int c0 = c_lo + blockIdx.y * 32; c0 < c_hi; c0 += gridDim.y * 32) {
    __syncthreads();
    for (int e = tid; e < 32 * 32; e += 256) {
      const int t = e & 31, c = e >> 5;
      sU[t][c] = (t < k && c0 + c < c_hi) ? W[(int64_t)(c0 + c) * ldw + rb + t] : 0u;
    }
    __syncthreads();
    if (i0 + ti < m) {
      // register tiling: this thread owns row ti and the 16 columns c = tc, tc+2, ...; the L value of a pivot is loaded
      // once and reused for all 16 columns, old C values are requested up front
      uint64_t acc[16];
      uint32_t oldv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        acc[j] = 0;
        const int c = tc + 2 * j;
        oldv[j] = (c0 + c < c_hi) ? W[(int64_t)(c0 + c) * ldw + i0 + ti] : 0u;
      }
      if (big_mod) {
        for (int t = 0; t < k; ++t) {
          const uint64_t l = sL[t][ti];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = mod_u64(acc[j] + mod_u64(l * sU[t][tc + 2 * j], mp), mp);
        }
      } else {
        for (int t = 0; t < k; ++t) {
          const uint32_t l = sL[t][ti];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += (uint64_t)l * sU[t][tc + 2 * j];
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = tc + 2 * j;
        if (c0 + c < c_hi) W[(int64_t)(c0 + c) * ldw + i0 + ti] = submod_u32(oldv[j], (uint32_t)mod_u64(acc[j], mp), P);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// triangular inverse, base case: b x b diagonal blocks, one CTA each, matrix block in shared memory, one thread
// per column of the inverse (reference: 32-thread forward/backward_sub_kernel_32, substitution_inplace.jl:6-89,
// with a per-row device Euclid).  The b modular inverses of the diagonal are computed first, in parallel.
// ---------------------------------------------------------------------------------------------------
constexpr int TRI_B = 64;

__global__ void __launch_bounds__(TRI_B)
triinv_base_kernel(const uint32_t* __restrict__ T, int64_t ldt, uint32_t* __restrict__ X, int64_t ldx, int n, int upper, int unit_diag,
                   int* __restrict__ singular, const __grid_constant__ ModP mp) {
  __shared__ uint32_t sT[TRI_B][TRI_B + 1];
  __shared__ uint32_t dinv[TRI_B];
  const int b0 = blockIdx.x * TRI_B;
  const int nb = min(TRI_B, n - b0);
  const int j = threadIdx.x;
  const uint32_t P = (uint32_t)mp.P;
  // load block as LOWER triangular (transpose when upper): sT[i][k], i >= k
  for (int c = 0; c < nb; ++c) {
    if (j < nb) {
      const uint32_t v = T[(int64_t)(b0 + c) * ldt + b0 + j];  // element (j, c)
      if (upper) sT[c][j] = v; else sT[j][c] = v;
    }
  }
  __syncthreads();
  if (j < nb) {
    uint32_t d = unit_diag ? 1u : modinv_u32(sT[j][j], P);
    if (!unit_diag && d == 0 && P != 1) atomicExch(singular, 1);
    dinv[j] = d;
  }
  __syncthreads();
  if (j >= nb) return;
  // column j of the inverse of the lower-triangular sT by forward substitution; x kept in registers? (dynamic
  // index) -> keep in shared column of a second array would double smem; TRI_B=64 values in local memory is fine.
  uint32_t x[TRI_B];
#pragma unroll 1
  for (int i = 0; i < nb; ++i) {
    uint32_t v = 0;
    if (i >= j) {
      uint64_t acc = 0;
      for (int k = j; k < i; ++k) {
        acc += (uint64_t)sT[i][k] * x[k];
        if (mp.P > (1ull << 29) || ((k - j) & 31) == 31) acc = mod_u64(acc, mp);
      }
      const uint32_t s = (uint32_t)mod_u64(acc, mp);
      const uint32_t rhs = submod_u32(i == j ? 1u % P : 0u, s, P);
      v = mulmod_u32(rhs, dinv[i], mp);
    }
    x[i] = v;
  }
  // write back: lower -> X(i,j) = x[i]; upper -> inverse of transpose is transpose of inverse: X(j,i) = x[i]
  for (int i = 0; i < nb; ++i) {
    if (upper) X[(int64_t)(b0 + i) * ldx + b0 + j] = x[i];
    else X[(int64_t)(b0 + j) * ldx + b0 + i] = x[i];
  }
}

__global__ void gather_cols_kernel(uint32_t* __restrict__ dst, int64_t ldd, const uint32_t* __restrict__ src, int64_t lds, int rows,
                                   int ncols, const int* __restrict__ order) {
  const int j = blockIdx.y;
  if (j >= ncols) return;
  const int sj = order[j];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x)
    dst[(int64_t)j * ldd + i] = src[(int64_t)sj * lds + i];
}

__global__ void swap_pairs_kernel(uint32_t* __restrict__ X, int64_t ld, int rows, int cols, const long long* __restrict__ pairs,
                                  int n_pairs, int on_cols, int inverse) {
  // sequential replay of the transposition list, one thread per line (reference permutations.jl:49-62,112-125)
  const int line = blockIdx.x * blockDim.x + threadIdx.x;
  if (on_cols) {
    if (line >= rows) return;
    for (int e = 0; e < n_pairs; ++e) {
      const int q = inverse ? n_pairs - 1 - e : e;
      const long long a = pairs[2 * q] - 1, c = pairs[2 * q + 1] - 1;
      const uint32_t t = X[a * ld + line];
      X[a * ld + line] = X[c * ld + line];
      X[c * ld + line] = t;
    }
  } else {
    if (line >= cols) return;
    uint32_t* col = X + (int64_t)line * ld;
    for (int e = 0; e < n_pairs; ++e) {
      const int q = inverse ? n_pairs - 1 - e : e;
      const long long a = pairs[2 * q] - 1, c = pairs[2 * q + 1] - 1;
      const uint32_t t = col[a];
      col[a] = col[c];
      col[c] = t;
    }
  }
}

__global__ void inv_table_kernel(uint32_t* __restrict__ tab, uint32_t N) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x < N) tab[x] = x ? modinv_u32(x, N) : 0u;
}

__global__ void modinv_batch_kernel(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, long long n,
                                    unsigned long long N) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = modinv_u64(in[i] % N, N);
}

// ---------------------------------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------------------------------
struct Elim {
  gffm_mat* W = nullptr;   // echelon form (U)
  gffm_mat* L = nullptr;
  int rank = 0;
  std::vector<int> pivcol, swp;
};

int panel_threads() {
  static const int t = (getenv("GFFM_PANEL_THREADS") && atoi(getenv("GFFM_PANEL_THREADS")) == 256) ? 256 : 512;
  return t;
}

template <bool SMALL, int THREADS>
int32_t launch_panel_t(gffm_ctx* ctx, gffm_mat* W, gffm_mat* L, int j0, int w, int rows_c_max, int cluster, const PluqBufs& b,
                       const ModP& mp) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cluster);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = (size_t)w * rows_c_max * 4;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const uint32_t mu32 = (uint32_t)((1ull << 32) / mp.P);
  GFFM_CUDA(cudaLaunchKernelEx(&cfg, (pluq_panel_kernel<SMALL, THREADS>), W->data, W->ld, (int)W->rows, j0, w, L->data, L->ld, b, mp, mu32));
  ctx->launches++;
  return GFFM_OK;
}

// cluster size of the panel kernel: 16 CTAs (non-portable size, opt-in) when the device can co-schedule them, else 8
int panel_cluster_size(gffm_ctx* ctx) {
  static std::mutex mu;
  static int chosen_dev[64] = {0};  // per device ordinal: the attributes below are per-device state
  std::lock_guard<std::mutex> guard(mu);
  int& chosen = chosen_dev[ctx->device & 63];
  if (chosen) return chosen;
  chosen = PANEL_CLUSTER;
  cudaFuncSetAttribute(pluq_panel_kernel<true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM_BUDGET);
  cudaFuncSetAttribute(pluq_panel_kernel<false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM_BUDGET);
  cudaFuncSetAttribute(pluq_panel_kernel<true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM_BUDGET);
  cudaFuncSetAttribute(pluq_panel_kernel<false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM_BUDGET);
  const char* env = getenv("GFFM_PANEL_CLUSTER");
  const int want = env ? atoi(env) : 16;
  if (want == 16 && cudaFuncSetAttribute(pluq_panel_kernel<true, 512>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
      cudaFuncSetAttribute(pluq_panel_kernel<false, 512>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
      cudaFuncSetAttribute(pluq_panel_kernel<true, 256>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
      cudaFuncSetAttribute(pluq_panel_kernel<false, 256>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(16);
    cfg.blockDim = dim3(PANEL_THREADS_MAX);
    cfg.dynamicSmemBytes = PANEL_SMEM_BUDGET;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 16;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, pluq_panel_kernel<false, 512>, &cfg) == cudaSuccess && nclusters >= 1) chosen = 16;
  } else if (want >= 1 && want <= 8) {
    chosen = want;
  }
  cudaGetLastError();
  return chosen;
}

PanelAr make_panel_ar(const ModP& mp) {
  PanelAr ar;
  memset(&ar, 0, sizeof(ar));
  ar.P = (uint32_t)mp.P;
  ar.mp = mp;
  {
    int bits = 0;
    while ((mp.P >> bits) != 0) ++bits;
    ar.sh2 = (uint32_t)(bits - 1);
    const unsigned __int128 q = ((unsigned __int128)1 << (32 + ar.sh2)) / mp.P;
    ar.mu2 = q > 0xffffffffull ? 0xffffffffu : (uint32_t)q;
  }
  if (mp.P <= 65536) {
    ar.mu = (uint32_t)((1ull << 32) / mp.P);
  } else if (mp.P < (1ull << 30)) {
    int bits = 0;
    while ((mp.P >> bits) != 0) ++bits;
    ar.sh = (uint32_t)(bits - 1);
    const unsigned __int128 q = ((unsigned __int128)1 << (32 + ar.sh)) / mp.P;
    ar.mu = q > 0xffffffffull ? 0xffffffffu : (uint32_t)q;
  }
  return ar;
}
int panel_arith(const ModP& mp) { return mp.P < 65536 ? 3 : (mp.P == 65536 ? 0 : (mp.P < (1ull << 30) ? 4 : 2)); }

template <int ARITH, int PW, int RPT>
int32_t launch_panel_reg_t(gffm_ctx* ctx, gffm_mat* W, gffm_mat* L, int j0, int w, int cluster, const PluqBufs& b, const ModP& mp) {
  static PerDeviceOnce attr;
  attr.run(ctx->device, [] {
    cudaFuncSetAttribute(pluq_panel_reg_kernel<ARITH, PW, RPT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaGetLastError();
  });
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cluster);
  cfg.blockDim = dim3(PREG_ROWS / RPT);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const PanelAr ar = make_panel_ar(mp);
  GFFM_CUDA(cudaLaunchKernelEx(&cfg, (pluq_panel_reg_kernel<ARITH, PW, RPT>), W->data, W->ld, (int)W->rows, j0, w, L->data, L->ld, b, ar));
  ctx->launches++;
  return GFFM_OK;
}

template <int PW, bool PROF>
int32_t launch_panel_ll_t(gffm_ctx* ctx, gffm_mat* W, gffm_mat* L, int j0, int w, int cluster, const PluqBufs& b, const ModP& mp) {
  static PerDeviceOnce attr;
  attr.run(ctx->device, [] {
    cudaFuncSetAttribute(pluq_panel_ll_kernel<PW, PROF>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaGetLastError();
  });
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cluster);
  cfg.blockDim = dim3(512);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const PanelAr ar = make_panel_ar(mp);
  GFFM_CUDA(cudaLaunchKernelEx(&cfg, (pluq_panel_ll_kernel<PW, PROF>), W->data, W->ld, (int)W->rows, j0, w, L->data, L->ld, b, ar));
  ctx->launches++;
  return GFFM_OK;
}
// left-looking variant: 31 products of two residues must fit a 64-bit accumulator
bool panel_ll_enabled(const ModP& mp) {
  static const bool on = !(getenv("GFFM_PANEL_LL") && atoi(getenv("GFFM_PANEL_LL")) == 0);
  return on && mp.P < (1ull << 29);
}

// register-resident panel: needs rows_c <= PREG_ROWS and a panel of at most 32 columns
bool panel_reg_enabled() {
  static const bool on = !(getenv("GFFM_PANEL_REG") && atoi(getenv("GFFM_PANEL_REG")) == 0);
  return on;
}
int32_t launch_panel_reg(gffm_ctx* ctx, gffm_mat* W, gffm_mat* L, int j0, int w, int cluster, const PluqBufs& b, const ModP& mp) {
  if (panel_ll_enabled(mp)) {
    if (b.prof) return w <= 16 ? launch_panel_ll_t<16, true>(ctx, W, L, j0, w, cluster, b, mp) : launch_panel_ll_t<32, true>(ctx, W, L, j0, w, cluster, b, mp);
    return w <= 16 ? launch_panel_ll_t<16, false>(ctx, W, L, j0, w, cluster, b, mp) : launch_panel_ll_t<32, false>(ctx, W, L, j0, w, cluster, b, mp);
  }
  const int ar = panel_arith(mp);
  static const int rpt = (getenv("GFFM_PANEL_RPT") && atoi(getenv("GFFM_PANEL_RPT")) == 1) ? 1 : 2;
#define GFFM_PREG(A)                                                                                                   \
  {                                                                                                                    \
    if (rpt == 2)                                                                                                      \
      return w <= 16 ? launch_panel_reg_t<A, 16, 2>(ctx, W, L, j0, w, cluster, b, mp)                                  \
                     : launch_panel_reg_t<A, 32, 2>(ctx, W, L, j0, w, cluster, b, mp);                                 \
    return w <= 16 ? launch_panel_reg_t<A, 16, 1>(ctx, W, L, j0, w, cluster, b, mp)                                    \
                   : launch_panel_reg_t<A, 32, 1>(ctx, W, L, j0, w, cluster, b, mp);                                   \
  }
  if (ar == 3) GFFM_PREG(3)
  if (ar == 0) GFFM_PREG(0)
  if (ar == 4) GFFM_PREG(4)
  GFFM_PREG(2)
#undef GFFM_PREG
}

int32_t launch_panel(gffm_ctx* ctx, gffm_mat* W, gffm_mat* L, int j0, int w, int rows_c_max, int cluster, const PluqBufs& b,
                     const ModP& mp) {
  if (panel_reg_enabled() && rows_c_max <= PREG_ROWS && w <= 32) return launch_panel_reg(ctx, W, L, j0, w, cluster, b, mp);
  if (panel_threads() == 256) {
    if (mp.P <= 65536) return launch_panel_t<true, 256>(ctx, W, L, j0, w, rows_c_max, cluster, b, mp);
    return launch_panel_t<false, 256>(ctx, W, L, j0, w, rows_c_max, cluster, b, mp);
  }
  if (mp.P <= 65536) return launch_panel_t<true, 512>(ctx, W, L, j0, w, rows_c_max, cluster, b, mp);
  return launch_panel_t<false, 512>(ctx, W, L, j0, w, rows_c_max, cluster, b, mp);
}

int32_t triinv_views(gffm_ctx* ctx, MatView T, MatView X, bool upper, bool unit_diag, uint64_t P, int* singular_dev,
                     const MatView* scratch_opt = nullptr);

// ---- recursive blocked elimination -----------------------------------------------------------------------------
// elim_rec(c_lo, c_hi): columns narrower than NB0 are factorised by cluster panels + SIMT rank-w updates; wider ranges are
// split in two, and after the left half the right half receives U12 = L11^-1 * A12 and the Schur update
// A22 -= L21 * U12 through the tensor-core GEMM with K = number of pivots found in the left half (up to n/2), so the
// modular GEMM runs at large K where it is tensor-bound instead of epilogue/HBM-bound.
// width of a base block (columns factorised by cluster panels + in-block updates before the tensor-core Schur update takes over);
// GFFM_ELIM_NB0 overrides it for experiments (multiple of 64)
static const int NB0 = [] {
  const char* e = getenv("GFFM_ELIM_NB0");
  const int v = e ? atoi(e) : 512;  // measured at n = 16384 (profiles/r02_notes.md): 128 -> 127 ms, 256 -> 112 ms, 512 -> 109 ms (mod 65521)
  return (v >= 64 && v <= 1024 && v % 64 == 0) ? v : 512;
}();

struct ElimState {
  gffm_ctx* ctx;
  gffm_mat *W, *L;
  int m, n;
  uint64_t N;
  ModP mp;
  PluqBufs b;
  int big_mod;
  int r;  // pivots found so far (host copy, refreshed after every base block)
  double t_panel = 0, t_u12 = 0, t_trail = 0;
  // stack allocator over ctx->ws_scratch for the Schur-update temporaries
  char* sbase = nullptr;
  size_t scap = 0;
  // one entry per base block that produced pivots: pivot rows [r_start, r_start+K) and the cached inverse of its
  // K x K diagonal block of L (leading dimension NB0), reused by every triangular solve above it in the recursion
  struct Block {
    int r_start, K;
    uint32_t* linv;
  };
  std::vector<Block> blocks;
  uint32_t* linv_pool = nullptr;
};

MatView scratch_view(ElimState& E, size_t& off, int64_t rows, int64_t cols) {
  const int64_t ld = round_up(std::max<int64_t>(rows, 1), 32);
  MatView v{reinterpret_cast<uint32_t*>(E.sbase + off), ld, rows, cols};
  off += (size_t)ld * std::max<int64_t>(cols, 1) * 4;
  off = (off + 255) & ~(size_t)255;
  return v;
}

int32_t base_block(ElimState& E, int c_lo, int c_hi) {
  gffm_ctx* ctx = E.ctx;
  gffm_mat *W = E.W, *L = E.L;
  const int m = E.m, n = E.n, r0 = E.r;
  const bool prof = ctx->profile;
  if (prof) cudaEventRecord(ctx->ev[4], ctx->stream);
  const int cluster = panel_cluster_size(ctx);
  const int rows_c_max = (int)ceil_div(m - r0, cluster);
  int w = PANEL_SMEM_BUDGET / (4 * std::max(rows_c_max, 1));
  // measured on B200 (profiles/r01_notes.md): 32-column panels win for 16-bit moduli (cheap 32-bit arithmetic in the rank-1
  // update), 16 columns for larger ones
  static const int w_env = getenv("GFFM_PANEL_W") ? std::max(1, std::min(PANEL_W_MAX, atoi(getenv("GFFM_PANEL_W")))) : 0;
  const bool reg_path = panel_reg_enabled() && rows_c_max <= PREG_ROWS;
  const int w_cap = w_env ? w_env : ((E.N <= 65536 || reg_path) ? 32 : 16);
  w = std::min(w, w_cap);
  if (w < 1) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "matrix has too many rows (%d) for the panel kernel", m);
  for (int j0 = c_lo; j0 < c_hi; j0 += w) {
    const int wj = std::min(w, c_hi - j0);
    GFFM_TRY(launch_panel(ctx, W, L, j0, wj, rows_c_max, cluster, E.b, E.mp));
    // row swaps of this panel: W columns right of the panel, all earlier L columns [0, rb)
    if (j0 + wj < n) {
      const int ncol = n - (j0 + wj);
      gather_rows_kernel<<<(unsigned)std::min<int64_t>(ceil_div(ncol, 8), 4 * ctx->num_sms), 256, 0, ctx->stream>>>(
          W->data, W->ld, j0 + wj, n, nullptr, E.b);
      GFFM_LAUNCH_CHECK(ctx);
    }
    if (j0 > 0) {
      gather_rows_kernel<<<(unsigned)std::min<int64_t>(ceil_div(std::max(std::min(j0, m), 1), 8), 4 * ctx->num_sms), 256, 0, ctx->stream>>>(
          L->data, L->ld, 0, 0, &E.b.st->rb, E.b);
      GFFM_LAUNCH_CHECK(ctx);
    }
    // rest of the base block: U12' = L11'^-1 W12', W22' -= L21' U12'
    const int lo = j0 + wj;
    if (lo < c_hi) {
      trsm_small_kernel<<<(unsigned)ceil_div(c_hi - lo, 8), 256, 0, ctx->stream>>>(W->data, W->ld, lo, c_hi, L->data, L->ld, E.b, E.mp);
      GFFM_LAUNCH_CHECK(ctx);
      dim3 grid((unsigned)ceil_div(m - r0, 128), (unsigned)std::min<int64_t>(ceil_div(c_hi - lo, 32), 8));
      update_small_kernel<<<grid, 256, 0, ctx->stream>>>(W->data, W->ld, m, lo, c_hi, L->data, L->ld, E.b, E.big_mod, E.mp);
      GFFM_LAUNCH_CHECK(ctx);
    }
  }
  if (prof) cudaEventRecord(ctx->ev[5], ctx->stream);
  PluqState hs;
  GFFM_CUDA(cudaMemcpyAsync(&hs, E.b.st, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
  GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
  E.r = hs.r;
  if (prof) {
    float a = 0;
    cudaEventElapsedTime(&a, ctx->ev[4], ctx->ev[5]);
    E.t_panel += a;
  }
  const int K = E.r - r0;
  if (K > 0 && E.linv_pool) {
    ElimState::Block blk{r0, K, E.linv_pool + (size_t)E.blocks.size() * NB0 * NB0};
    MatView X{blk.linv, NB0, K, K};
    GFFM_TRY(gffm_fill_view(ctx, X, 0));
    size_t off = 0;
    MatView tsc = scratch_view(E, off, NB0 / 2, NB0 / 2);
    GFFM_TRY(triinv_views(ctx, sub_view(view_of(L), r0, r0, K, K), X, /*upper=*/false, false, E.N, nullptr, &tsc));
    E.blocks.push_back(blk);
  }
  return GFFM_OK;
}

// W[rows of blocks b_lo..b_hi, c1..c2) <- L11^-1 * (same rows), L11 = the unit-free lower-triangular block of L spanned by
// those blocks.  Recursive over the block list: diagonal blocks use their cached inverses, off-diagonal parts are GEMMs
// whose K is the pivot count of the left part.
int32_t trsm_blocks(ElimState& E, int b_lo, int b_hi, int c1, int c2) {
  gffm_ctx* ctx = E.ctx;
  MatView Wv = view_of(E.W), Lv = view_of(E.L);
  const int nc = c2 - c1;
  if (b_hi - b_lo == 1) {
    const ElimState::Block& blk = E.blocks[b_lo];
    size_t off = 0;
    MatView tmp = scratch_view(E, off, blk.K, nc);
    if (off > E.scap) GFFM_FAIL(GFFM_ERR_OOM, "elimination scratch too small");
    MatView X{blk.linv, NB0, blk.K, blk.K};
    MatView rows = sub_view(Wv, blk.r_start, c1, blk.K, nc);
    GFFM_TRY(gffm_gemm_views(ctx, tmp, X, rows, E.N, E.N, GFFM_GEMM_STORE, GFFM_ALGO_AUTO));
    return gffm_copy_views(ctx, rows, tmp);
  }
  const int mid = (b_lo + b_hi) / 2;
  GFFM_TRY(trsm_blocks(E, b_lo, mid, c1, c2));
  const int r_lo = E.blocks[b_lo].r_start, r_mid = E.blocks[mid].r_start;
  const int r_hi = E.blocks[b_hi - 1].r_start + E.blocks[b_hi - 1].K;
  GFFM_TRY(gffm_gemm_views(ctx, sub_view(Wv, r_mid, c1, r_hi - r_mid, nc), sub_view(Lv, r_mid, r_lo, r_hi - r_mid, r_mid - r_lo),
                           sub_view(Wv, r_lo, c1, r_mid - r_lo, nc), E.N, E.N, GFFM_GEMM_SUB, GFFM_ALGO_AUTO));
  return trsm_blocks(E, mid, b_hi, c1, c2);
}

// columns [c1, c2): rows r0..r0+K become U12 = L11^-1 * W12, rows below get W22 -= L21 * U12
int32_t schur_update(ElimState& E, int r0, int K, int c1, int c2) {
  gffm_ctx* ctx = E.ctx;
  const bool prof = ctx->profile;
  const int nc = c2 - c1;
  if (prof) cudaEventRecord(ctx->ev[5], ctx->stream);
  MatView Wv = view_of(E.W), Lv = view_of(E.L);
  // blocks covering pivot rows [r0, r0+K)
  int b_lo = -1, b_hi = -1;
  for (int i = 0; i < (int)E.blocks.size(); ++i) {
    if (E.blocks[i].r_start == r0) b_lo = i;
    if (E.blocks[i].r_start + E.blocks[i].K == r0 + K) b_hi = i + 1;
  }
  if (b_lo < 0 || b_hi <= b_lo) GFFM_FAIL(GFFM_ERR_INVALID, "internal: pivot blocks do not tile [%d,%d)", r0, r0 + K);
  GFFM_TRY(trsm_blocks(E, b_lo, b_hi, c1, c2));
  if (prof) cudaEventRecord(ctx->ev[6], ctx->stream);
  if (r0 + K < E.m)
    GFFM_TRY(gffm_gemm_views(ctx, sub_view(Wv, r0 + K, c1, E.m - r0 - K, nc), sub_view(Lv, r0 + K, r0, E.m - r0 - K, K),
                             sub_view(Wv, r0, c1, K, nc), E.N, E.N, GFFM_GEMM_SUB, GFFM_ALGO_AUTO));
  if (prof) {
    cudaEventRecord(ctx->ev[7], ctx->stream);
    cudaEventSynchronize(ctx->ev[7]);
    float a = 0, b2 = 0;
    cudaEventElapsedTime(&a, ctx->ev[5], ctx->ev[6]);
    cudaEventElapsedTime(&b2, ctx->ev[6], ctx->ev[7]);
    E.t_u12 += a;
    E.t_trail += b2;
  }
  return GFFM_OK;
}

int32_t elim_rec(ElimState& E, int c_lo, int c_hi) {
  if (E.r >= E.m || c_lo >= c_hi) return GFFM_OK;
  if (c_hi - c_lo <= NB0) return base_block(E, c_lo, c_hi);
  const int mid = c_lo + (int)round_up((c_hi - c_lo + 1) / 2, NB0);
  const int r_before = E.r;
  GFFM_TRY(elim_rec(E, c_lo, mid));
  const int K = E.r - r_before;
  if (K > 0 && mid < c_hi) GFFM_TRY(schur_update(E, r_before, K, mid, c_hi));
  return elim_rec(E, mid, c_hi);
}

// Row-echelon elimination of A (copy) with L; leaves everything on the device, returns rank/pivots on the host.
int32_t eliminate(gffm_mat* A, Elim* out) {
  gffm_ctx* ctx = A->ctx;
  const int m = (int)A->rows, n = (int)A->cols;
  const uint64_t N = A->N;
  GFFM_NARROW_ONLY(A);
  if (N >= (1ull << 32)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "elimination needs N < 2^32");
  gffm_mat *W = nullptr, *L = nullptr;
  static const bool host_prof = getenv("GFFM_HOST_PROF") != nullptr;
  auto now_ms = [&]() {
    cudaStreamSynchronize(ctx->stream);
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  };
  const double hp0 = host_prof ? now_ms() : 0;
  GFFM_TRY(gffm_mat_create(ctx, m, n, N, A->pad, &W));
  out->W = W;
  GFFM_TRY(gffm_mat_create(ctx, m, m, N, A->pad, &L));
  out->L = L;
  GFFM_TRY(gffm_copy_views(ctx, view_of(W), view_of(A)));
  const int maxr = std::min(m, n);
  out->rank = 0;
  if (maxr == 0) return GFFM_OK;
  ElimState E;
  E.ctx = ctx;
  E.W = W;
  E.L = L;
  E.m = m;
  E.n = n;
  E.N = N;
  E.mp = make_modp(N);
  E.big_mod = N > (1ull << 29) ? 1 : 0;
  E.r = 0;
  // device bookkeeping
  const size_t need = sizeof(PluqState) + (size_t)maxr * (sizeof(int) * 2 + sizeof(uint32_t)) + 2 * 64 * sizeof(int) + 256;
  GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_misc2, need));
  char* base = (char*)ctx->ws_misc2.ptr;
  PluqBufs& b = E.b;
  b.st = (PluqState*)base;
  b.pivcol = (int*)(base + 64);
  b.swp = b.pivcol + maxr;
  b.pinv = (uint32_t*)(b.swp + maxr);
  b.g_dst = (int*)(b.pinv + maxr);
  b.g_src = b.g_dst + 64;
  GFFM_CUDA(cudaMemsetAsync(base, 0, need, ctx->stream));
  b.prof = nullptr;
  if (getenv("GFFM_PANEL_PROF")) {
    GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_misc, 256));
    GFFM_CUDA(cudaMemsetAsync(ctx->ws_misc.ptr, 0, 128, ctx->stream));
    b.prof = (long long*)ctx->ws_misc.ptr;
  }
  b.inv_table = nullptr;
  if (N <= (1u << 26)) {  // batched modular inverses: one kernel computes every inverse mod N (<= 256 MiB), cached per context
    if (ctx->inv_table_N != N) {
      GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_invtab, (size_t)N * 4));
      inv_table_kernel<<<(unsigned)ceil_div((int64_t)N, 256), 256, 0, ctx->stream>>>((uint32_t*)ctx->ws_invtab.ptr, (uint32_t)N);
      GFFM_LAUNCH_CHECK(ctx);
      ctx->inv_table_N = N;
    }
    b.inv_table = (const uint32_t*)ctx->ws_invtab.ptr;
  }
  // scratch: one NB0 x n tile for the triangular solves (+ the 128 x 128 scratch of the block inversions) and the pool
  // of cached NB0 x NB0 diagonal-block inverses
  if (n > NB0) {
    const size_t nblk = (size_t)ceil_div(n, NB0) + 1;
    const size_t tmp_bytes = 4 * ((size_t)NB0 * n + (size_t)NB0 * NB0) + 8192;
    GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_scratch, tmp_bytes + nblk * NB0 * NB0 * 4));
    E.sbase = (char*)ctx->ws_scratch.ptr;
    E.scap = tmp_bytes;
    E.linv_pool = reinterpret_cast<uint32_t*>(E.sbase + tmp_bytes);
  }
  const double hp1 = host_prof ? now_ms() : 0;
  GFFM_TRY(elim_rec(E, 0, n));
  if (host_prof) {
    const double hp2 = now_ms();
    fprintf(stderr, "[elim host prof] setup (alloc W,L + copy + tables) %.2f ms, elimination %.2f ms\n", hp1 - hp0, hp2 - hp1);
  }
  const int r0 = E.r;
  if (b.prof) {
    long long h[16];
    cudaMemcpyAsync(h, b.prof, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    static const char* names_smem[8] = {"load+sync", "A scan+inverse+S1+publish", "cluster.sync", "B dsmem gather+S2+select", "u row + S3",
                                        "C update (warp 0)", "store-back", "-"};
    static const char* names_reg[8] = {"load+sync", "warp argmax + S1", "block argmax + S1b", "push (remote stores)", "cluster.sync",
                                       "select winner", "eliminate+rotate (thread 0)", "store-back"};
    const char** names = panel_reg_enabled() ? names_reg : names_smem;
    long long tot = 0;
    for (int i = 0; i < 8; ++i) tot += h[i];
    fprintf(stderr, "[panel prof] total %.3f Mcycles over %d pivots\n", tot / 1e6, r0);
    for (int i = 0; i < 8; ++i) fprintf(stderr, "  %-24s %10.3f Mcyc  %5.1f%%  %8.1f cyc/pivot\n", names[i], h[i] / 1e6, 100.0 * h[i] / tot, (double)h[i] / std::max(r0, 1));
  }
  if (ctx->profile) {
    ctx->n_ev = 0;
    ctx->elim_timings = {E.t_panel, E.t_u12, E.t_trail};
  }
  out->rank = r0;
  out->pivcol.resize(r0);
  out->swp.resize(r0);
  if (r0 > 0) {
    std::vector<uint32_t> pinv(r0);
    GFFM_CUDA(cudaMemcpyAsync(out->pivcol.data(), b.pivcol, sizeof(int) * r0, cudaMemcpyDeviceToHost, ctx->stream));
    GFFM_CUDA(cudaMemcpyAsync(out->swp.data(), b.swp, sizeof(int) * r0, cudaMemcpyDeviceToHost, ctx->stream));
    GFFM_CUDA(cudaMemcpyAsync(pinv.data(), b.pinv, sizeof(uint32_t) * r0, cudaMemcpyDeviceToHost, ctx->stream));
    GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
    // The pivot rule is "largest residue" (pluq_kernels.jl:189), which over a composite modulus can pick a zero divisor: its
    // inverse is 0 (mod_inv, pluq_kernels.jl:11-31 has no answer either) and everything scaled by it is garbage.  The reference
    // declares CuModMatrixModulusNotPrimeException (CuModMatrix.jl:23-25) for this situation; report it instead of a wrong result.
    for (int t = 0; t < r0; ++t)
      if (pinv[t] == 0 && N > 1)
        GFFM_FAIL(GFFM_ERR_MODULUS_NOT_PRIME, "elimination mod %llu: pivot %d is a zero divisor (the modulus is not prime)", (unsigned long long)N, t);
  }
  return GFFM_OK;
}

// triangular inverse on views: X = T^-1 (T n x n lower or upper triangular; strict other half ignored, X's other
// half must be zero on entry).  Recursive halving at multiples of TRI_B; combine with two GEMMs:
//   lower: X21 = -X22 * T21 * X11        upper: X12 = -X11 * T12 * X22
// (reference triangular_inverse_no_copy.jl:153-157,182-186,406-410,435-439 -- there with unguarded cuBLAS float GEMM)
int32_t triinv_rec(gffm_ctx* ctx, MatView T, MatView X, bool upper, uint64_t P, int lo, int hi, const MatView& scratch) {
  const int len = hi - lo;
  if (len <= TRI_B) return GFFM_OK;  // base blocks were inverted in one batched launch
  const int half = (int)round_up((len + 1) / 2, TRI_B);
  const int mid = lo + half;
  GFFM_TRY(triinv_rec(ctx, T, X, upper, P, lo, mid, scratch));
  GFFM_TRY(triinv_rec(ctx, T, X, upper, P, mid, hi, scratch));
  const int n1 = mid - lo, n2 = hi - mid;
  if (!upper) {
    MatView T21 = sub_view(T, mid, lo, n2, n1), X11 = sub_view(X, lo, lo, n1, n1), X22 = sub_view(X, mid, mid, n2, n2),
            X21 = sub_view(X, mid, lo, n2, n1);
    MatView tmp = sub_view(scratch, 0, 0, n2, n1);
    GFFM_TRY(gffm_gemm_views(ctx, tmp, T21, X11, P, P, GFFM_GEMM_STORE, GFFM_ALGO_AUTO));
    GFFM_TRY(gffm_fill_view(ctx, X21, 0));
    GFFM_TRY(gffm_gemm_views(ctx, X21, X22, tmp, P, P, GFFM_GEMM_SUB, GFFM_ALGO_AUTO));
  } else {
    MatView T12 = sub_view(T, lo, mid, n1, n2), X11 = sub_view(X, lo, lo, n1, n1), X22 = sub_view(X, mid, mid, n2, n2),
            X12 = sub_view(X, lo, mid, n1, n2);
    MatView tmp = sub_view(scratch, 0, 0, n1, n2);
    GFFM_TRY(gffm_gemm_views(ctx, tmp, T12, X22, P, P, GFFM_GEMM_STORE, GFFM_ALGO_AUTO));
    GFFM_TRY(gffm_fill_view(ctx, X12, 0));
    GFFM_TRY(gffm_gemm_views(ctx, X12, X11, tmp, P, P, GFFM_GEMM_SUB, GFFM_ALGO_AUTO));
  }
  return GFFM_OK;
}

int32_t triinv_views(gffm_ctx* ctx, MatView T, MatView X, bool upper, bool unit_diag, uint64_t P, int* singular_dev,
                     const MatView* scratch_opt) {
  const int n = (int)T.rows;
  if (T.cols != n || X.rows != n || X.cols != n) GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "triinv: square views expected");
  if (n == 0) return GFFM_OK;
  int* sing = singular_dev;
  if (!sing) {
    GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_misc, 64));
    sing = (int*)ctx->ws_misc.ptr;
    GFFM_CUDA(cudaMemsetAsync(sing, 0, sizeof(int), ctx->stream));
  }
  triinv_base_kernel<<<(unsigned)ceil_div(n, TRI_B), TRI_B, 0, ctx->stream>>>(T.p, T.ld, X.p, X.ld, n, upper ? 1 : 0, unit_diag ? 1 : 0,
                                                                               sing, make_modp(P));
  GFFM_LAUNCH_CHECK(ctx);
  if (n <= TRI_B) return GFFM_OK;
  const int half = (int)round_up((n + 1) / 2, TRI_B);
  if (scratch_opt) {
    if (scratch_opt->rows < half || scratch_opt->cols < half) GFFM_FAIL(GFFM_ERR_INVALID, "triinv scratch too small");
    return triinv_rec(ctx, T, X, upper, P, 0, n, *scratch_opt);
  }
  gffm_mat* scratch = nullptr;
  GFFM_TRY(gffm_mat_create(ctx, half, half, P, 0, &scratch));
  int32_t st = triinv_rec(ctx, T, X, upper, P, 0, n, view_of(scratch));
  gffm_mat_destroy(scratch);
  return st;
}

void fill_row_pairs(const Elim& e, int64_t* pairs, int64_t* n_pairs) {
  int64_t k = 0;
  for (int t = 0; t < e.rank; ++t) {
    if (e.swp[t] != t) {  // reference pushes only when row != prow (pluq_kernels.jl:258-260)
      if (pairs) {
        pairs[2 * k] = t + 1;
        pairs[2 * k + 1] = e.swp[t] + 1;
      }
      ++k;
    }
  }
  if (n_pairs) *n_pairs = k;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" int32_t gffm_lu(gffm_mat* A, gffm_mat** U, gffm_mat** L, int64_t* prow_pairs, int64_t* n_prow, int64_t* pivcols,
                           int64_t* rank) {
  GFFM_ENTER_MAT(A);
  if (!A || !U || !L) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  Elim e;
  int32_t st = eliminate(A, &e);
  if (st != GFFM_OK) {
    gffm_mat_destroy(e.W);
    gffm_mat_destroy(e.L);
    return st;
  }
  fill_row_pairs(e, prow_pairs, n_prow);
  if (pivcols)
    for (int t = 0; t < e.rank; ++t) pivcols[t] = e.pivcol[t];
  if (rank) *rank = e.rank;
  *U = e.W;
  *L = e.L;
  return GFFM_OK;
}

extern "C" int32_t gffm_rank(gffm_mat* A, int64_t* rank) {
  GFFM_ENTER_MAT(A);
  if (!A || !rank) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  Elim e;
  int32_t st = eliminate(A, &e);
  gffm_mat_destroy(e.W);
  gffm_mat_destroy(e.L);
  GFFM_TRY(st);
  *rank = e.rank;
  return GFFM_OK;
}

int32_t gffm_pluq_quirk(gffm_mat* A, gffm_mat** U, gffm_mat** L, int64_t* prow_pairs, int64_t* n_prow, int64_t* pcol_pairs,
                        int64_t* n_pcol, int64_t* rank);

extern "C" int32_t gffm_pluq(gffm_mat* A, gffm_mat** U, gffm_mat** L, int64_t* prow_pairs, int64_t* n_prow, int64_t* pcol_pairs,
                             int64_t* n_pcol, int64_t* rank, int32_t col_pivot_mode) {
  GFFM_ENTER_MAT(A);
  if (!A || !U || !L) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (col_pivot_mode == GFFM_PIVOT_REFERENCE_QUIRK) return gffm_pluq_quirk(A, U, L, prow_pairs, n_prow, pcol_pairs, n_pcol, rank);
  gffm_ctx* ctx = A->ctx;
  Elim e;
  int32_t st = eliminate(A, &e);
  if (st != GFFM_OK) {
    gffm_mat_destroy(e.W);
    gffm_mat_destroy(e.L);
    return st;
  }
  fill_row_pairs(e, prow_pairs, n_prow);
  if (rank) *rank = e.rank;
  const int n = (int)A->cols;
  // column order: pivot columns first (stable); transposition list equivalent to it
  std::vector<int> order;
  order.reserve(n);
  std::vector<char> isp(n, 0);
  for (int t = 0; t < e.rank; ++t) {
    order.push_back(e.pivcol[t]);
    isp[e.pivcol[t]] = 1;
  }
  bool identity = true;
  for (int t = 0; t < e.rank; ++t)
    if (e.pivcol[t] != t) identity = false;
  for (int c = 0; c < n; ++c)
    if (!isp[c]) order.push_back(c);
  int64_t npc = 0;
  if (!identity) {
    std::vector<int> cur(n), pos(n);
    for (int c = 0; c < n; ++c) cur[c] = pos[c] = c;
    for (int j = 0; j < n; ++j) {
      const int want = order[j], pj = pos[want];
      if (pj != j) {
        if (pcol_pairs) {
          pcol_pairs[2 * npc] = j + 1;
          pcol_pairs[2 * npc + 1] = pj + 1;
        }
        ++npc;
        const int cj = cur[j];
        cur[j] = want;
        cur[pj] = cj;
        pos[want] = j;
        pos[cj] = pj;
      }
    }
    gffm_mat* Up = nullptr;
    st = gffm_mat_create(ctx, A->rows, A->cols, A->N, A->pad, &Up);
    if (st == GFFM_OK) st = gffm_ws_reserve(ctx, &ctx->ws_misc, sizeof(int) * n);
    if (st == GFFM_OK) {
      cudaMemcpyAsync(ctx->ws_misc.ptr, order.data(), sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream);
      for (int c0 = 0; c0 < n; c0 += 65535) {
        const int nc = std::min(65535, n - c0);
        dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(A->rows, 256), 64)), (unsigned)nc);
        gather_cols_kernel<<<grid, 256, 0, ctx->stream>>>(Up->data + (int64_t)c0 * Up->ld, Up->ld, e.W->data, e.W->ld, (int)A->rows, nc,
                                                          (const int*)ctx->ws_misc.ptr + c0);
        ctx->launches++;
      }
      cudaStreamSynchronize(ctx->stream);  // order[] is a host vector
      gffm_mat_destroy(e.W);
      e.W = Up;
    } else {
      gffm_mat_destroy(e.W);
      gffm_mat_destroy(e.L);
      return st;
    }
  }
  if (n_pcol) *n_pcol = npc;
  *U = e.W;
  *L = e.L;
  return GFFM_OK;
}

extern "C" int32_t gffm_rref(gffm_mat* A, gffm_mat** R, int64_t* pivcols, int64_t* rank) {
  GFFM_ENTER_MAT(A);
  if (!A || !R) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  gffm_ctx* ctx = A->ctx;
  Elim e;
  int32_t st = eliminate(A, &e);
  gffm_mat_destroy(e.L);
  if (st != GFFM_OK) {
    gffm_mat_destroy(e.W);
    return st;
  }
  if (pivcols)
    for (int t = 0; t < e.rank; ++t) pivcols[t] = e.pivcol[t];
  if (rank) *rank = e.rank;
  const int r = e.rank, n = (int)A->cols;
  if (r > 1) {
    // R[0:r,:] = T^-1 * E[0:r,:], T = E[0:r, pivcols] (unit upper triangular)
    gffm_mat *T = nullptr, *Ti = nullptr, *Out = nullptr;
    st = gffm_mat_create(ctx, r, r, A->N, 0, &T);
    if (st == GFFM_OK) st = gffm_mat_create(ctx, r, r, A->N, 0, &Ti);
    if (st == GFFM_OK) st = gffm_mat_create(ctx, A->rows, A->cols, A->N, A->pad, &Out);
    if (st == GFFM_OK) st = gffm_ws_reserve(ctx, &ctx->ws_misc2, sizeof(int) * r + 64);
    if (st == GFFM_OK) {
      int* dord = (int*)((char*)ctx->ws_misc2.ptr + 64);
      cudaMemcpyAsync(dord, e.pivcol.data(), sizeof(int) * r, cudaMemcpyHostToDevice, ctx->stream);
      for (int c0 = 0; c0 < r; c0 += 65535) {
        const int nc = std::min(65535, r - c0);
        dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(r, 256), 64)), (unsigned)nc);
        gather_cols_kernel<<<grid, 256, 0, ctx->stream>>>(T->data + (int64_t)c0 * T->ld, T->ld, e.W->data, e.W->ld, r, nc, dord + c0);
        ctx->launches++;
      }
      st = triinv_views(ctx, view_of(T), view_of(Ti), /*upper=*/true, /*unit=*/true, A->N, nullptr);
    }
    if (st == GFFM_OK)
      st = gffm_gemm_views(ctx, sub_view(view_of(Out), 0, 0, r, n), view_of(Ti), sub_view(view_of(e.W), 0, 0, r, n), A->N, A->N,
                           GFFM_GEMM_STORE, GFFM_ALGO_AUTO);
    gffm_mat_destroy(T);
    gffm_mat_destroy(Ti);
    if (st != GFFM_OK) {
      gffm_mat_destroy(Out);
      gffm_mat_destroy(e.W);
      return st;
    }
    gffm_mat_destroy(e.W);
    e.W = Out;
  }
  *R = e.W;
  return GFFM_OK;
}

extern "C" int32_t gffm_triinv(gffm_mat* A, int32_t upper, gffm_mat** out) {
  GFFM_ENTER_MAT(A);
  if (!A || !out) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  GFFM_NARROW_ONLY(A);
  gffm_ctx* ctx = A->ctx;
  const int64_t rows = A->rows, cols = A->cols;
  if (!upper && rows > cols) GFFM_FAIL(GFFM_ERR_INVERSE_NOT_DEFINED, "lower triangular inverse of a tall matrix is not defined");
  if (upper && rows > cols) GFFM_FAIL(GFFM_ERR_INVERSE_NOT_DEFINED, "upper triangular inverse of a tall matrix is not defined");
  // wide input: invert the leading rows x rows block, result is cols x rows = [T^-1; 0]
  // (reference triangular_inverse_no_copy.jl:197-228)
  gffm_mat* X = nullptr;
  GFFM_TRY(gffm_mat_create(ctx, cols, rows, A->N, A->pad, &X));
  GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_misc, 64));
  int* sing = (int*)ctx->ws_misc.ptr;
  cudaMemsetAsync(sing, 0, sizeof(int), ctx->stream);
  int32_t st = triinv_views(ctx, sub_view(view_of(A), 0, 0, rows, rows), sub_view(view_of(X), 0, 0, rows, rows), upper != 0, false, A->N, sing);
  int hs = 0;
  if (st == GFFM_OK) {
    cudaMemcpyAsync(&hs, sing, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    if (hs) {
      st = GFFM_ERR_NOT_INVERTIBLE;
      gffm_set_error("triangular matrix has a non-invertible diagonal entry");
    }
  }
  if (st != GFFM_OK) {
    gffm_mat_destroy(X);
    return st;
  }
  *out = X;
  return GFFM_OK;
}

extern "C" int32_t gffm_apply_perm(gffm_mat* A, const int64_t* pairs, int64_t n_pairs, int32_t on_cols, int32_t inverse) {
  GFFM_ENTER_MAT(A);
  if (!A || (n_pairs > 0 && !pairs)) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  GFFM_NARROW_ONLY(A);
  if (n_pairs <= 0) return GFFM_OK;
  gffm_touch(A);
  gffm_ctx* ctx = A->ctx;
  const int64_t lim = on_cols ? A->cols : A->rows;
  for (int64_t e = 0; e < 2 * n_pairs; ++e)
    if (pairs[e] < 1 || pairs[e] > lim) GFFM_FAIL(GFFM_ERR_INVALID, "BoundsError: permutation index %lld out of 1..%lld", (long long)pairs[e], (long long)lim);
  GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_misc, sizeof(long long) * 2 * n_pairs));
  GFFM_CUDA(cudaMemcpyAsync(ctx->ws_misc.ptr, pairs, sizeof(long long) * 2 * n_pairs, cudaMemcpyHostToDevice, ctx->stream));
  const int64_t lines = on_cols ? A->rows : A->cols;
  if (lines > 0) {
    swap_pairs_kernel<<<(unsigned)ceil_div(lines, 128), 128, 0, ctx->stream>>>(A->data, A->ld, (int)A->rows, (int)A->cols,
                                                                               (const long long*)ctx->ws_misc.ptr, (int)n_pairs, on_cols, inverse);
    GFFM_LAUNCH_CHECK(ctx);
  }
  GFFM_CUDA(cudaStreamSynchronize(ctx->stream));  // pairs is a caller buffer
  return GFFM_OK;
}

extern "C" int32_t gffm_modinv_batch(gffm_ctx* ctx, const uint64_t* in_host, uint64_t* out_host, int64_t n, uint64_t N) {
  GFFM_ENTER_CTX(ctx);
  if (!ctx || (n > 0 && (!in_host || !out_host))) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (N == 0 || N >= (1ull << 62)) GFFM_FAIL(GFFM_ERR_INVALID, "modulus out of range");
  if (n <= 0) return GFFM_OK;
  GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_misc, 16 * (size_t)n));
  unsigned long long* din = (unsigned long long*)ctx->ws_misc.ptr;
  unsigned long long* dout = din + n;
  GFFM_CUDA(cudaMemcpyAsync(din, in_host, 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  modinv_batch_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(din, dout, n, N);
  GFFM_LAUNCH_CHECK(ctx);
  GFFM_CUDA(cudaMemcpyAsync(out_host, dout, 8 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return GFFM_OK;
}

// inverse(A) = U^-1 * L^-1 * P  (reference CuModMatrix.jl:480-502: apply_col_inv_perm!(P, L_inv); U_inv * L_inv)
extern "C" int32_t gffm_inverse(gffm_mat* A, gffm_mat** Ainv, int32_t* invertible) {
  GFFM_ENTER_MAT(A);
  if (!A || !Ainv || !invertible) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  *Ainv = nullptr;
  *invertible = 0;
  if (A->rows != A->cols) GFFM_FAIL(GFFM_ERR_NOT_SQUARE, "inverse of a %lldx%lld matrix", (long long)A->rows, (long long)A->cols);
  gffm_ctx* ctx = A->ctx;
  const int n = (int)A->rows;
  Elim e;
  int32_t st = eliminate(A, &e);
  if (st != GFFM_OK || e.rank < n) {  // invertibility = pivot count (replaces the reference's CPU SVD rank, :340-347)
    gffm_mat_destroy(e.W);
    gffm_mat_destroy(e.L);
    return st;
  }
  gffm_mat *Ui = nullptr, *Li = nullptr, *Out = nullptr;
  st = gffm_mat_create(ctx, n, n, A->N, A->pad, &Ui);
  if (st == GFFM_OK) st = gffm_mat_create(ctx, n, n, A->N, A->pad, &Li);
  if (st == GFFM_OK) st = gffm_mat_create(ctx, n, n, A->N, A->pad, &Out);
  if (st == GFFM_OK) st = triinv_views(ctx, view_of(e.W), view_of(Ui), true, true, A->N, nullptr);
  if (st == GFFM_OK) st = triinv_views(ctx, view_of(e.L), view_of(Li), false, false, A->N, nullptr);
  if (st == GFFM_OK) {
    std::vector<int64_t> pairs(2 * (size_t)std::max(n, 1));
    int64_t np = 0;
    fill_row_pairs(e, pairs.data(), &np);
    st = gffm_apply_perm(Li, pairs.data(), np, /*on_cols=*/1, /*inverse=*/1);
  }
  if (st == GFFM_OK) st = gffm_gemm_views(ctx, view_of(Out), view_of(Ui), view_of(Li), A->N, A->N, GFFM_GEMM_STORE, GFFM_ALGO_AUTO);
  gffm_mat_destroy(e.W);
  gffm_mat_destroy(e.L);
  gffm_mat_destroy(Ui);
  gffm_mat_destroy(Li);
  if (st != GFFM_OK) {
    gffm_mat_destroy(Out);
    return st;
  }
  *Ainv = Out;
  *invertible = 1;
  return GFFM_OK;
}
