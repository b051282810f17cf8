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
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int PANEL_THREADS = 1024;
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
};

__device__ __forceinline__ bool better(uint32_t v, int i, uint32_t bv, int bi) { return v > bv || (v == bv && i < bi); }

// ---------------------------------------------------------------------------------------------------
// panel kernel: columns [j0, j0+w) of W, rows [r, m).  One cluster; CTA c owns rows rb + c*rows_c + [0,rows_c).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PANEL_THREADS, 1)
pluq_panel_kernel(uint32_t* __restrict__ W, int64_t ldw, int m, int j0, int w, uint32_t* __restrict__ Lm, int64_t ldl,
                  PluqBufs b, const __grid_constant__ ModP mp) {
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  extern __shared__ uint32_t panel[];  // [w][rows_c]
  __shared__ uint32_t cand_val[2];
  __shared__ int cand_idx[2];
  __shared__ uint32_t red_val[32];
  __shared__ int red_idx[32];
  __shared__ uint32_t sel_val;
  __shared__ int sel_idx;
  __shared__ uint32_t rowbuf_p[PANEL_W_MAX], rowbuf_r[PANEL_W_MAX], u_s[PANEL_W_MAX], oldr_s[PANEL_W_MAX];
  __shared__ uint32_t pinv_s;

  const int rb = b.st->r;
  const int rows_total = m - rb;
  const int rows_c = (rows_total + cs - 1) / cs;
  const int my_lo = rb + rank * rows_c;
  int my_n = rows_total - rank * rows_c;
  my_n = my_n < 0 ? 0 : (my_n > rows_c ? rows_c : my_n);
  const uint32_t P = (uint32_t)mp.P;

  for (int c = 0; c < w; ++c)
    for (int q = tid; q < my_n; q += PANEL_THREADS) panel[c * rows_c + q] = W[(int64_t)(j0 + c) * ldw + my_lo + q];
  cluster.sync();  // everybody has read st->r before anyone can finish and overwrite it

  int r = rb;
  for (int jj = 0; jj < w && r < m; ++jj) {
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
      red_val[warp] = bv;
      red_idx[warp] = bi;
    }
    __syncthreads();
    if (warp == 0) {
      bv = red_val[lane];
      bi = red_idx[lane];
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
        cand_val[jj & 1] = bv;
        cand_idx[jj & 1] = bi;
      }
    }
    cluster.sync();
    // ---- phase B: cluster-wide winner through distributed shared memory
    if (warp == 0) {
      uint32_t v = 0;
      int i = 0x7fffffff;
      if (lane < cs) {
        v = *cluster.map_shared_rank(&cand_val[jj & 1], lane);
        i = *cluster.map_shared_rank(&cand_idx[jj & 1], lane);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const uint32_t ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (better(ov, oi, v, i)) {
          v = ov;
          i = oi;
        }
      }
      if (lane == 0) {
        sel_val = v;
        sel_idx = i;
      }
    }
    __syncthreads();
    const uint32_t pv = sel_val;
    const int p = sel_idx;
    if (pv == 0) continue;  // no pivot in this column (uniform over the cluster): skip it, row stays
    const int owner_p = (p - rb) / rows_c, owner_r = (r - rb) / rows_c;
    if (rank == owner_p && tid < w) rowbuf_p[tid] = panel[tid * rows_c + (p - my_lo)];
    if (rank == owner_r && tid >= 32 && tid < 32 + w) rowbuf_r[tid - 32] = panel[(tid - 32) * rows_c + (r - my_lo)];
    cluster.sync();
    if (tid < w) {
      const uint32_t xp = *cluster.map_shared_rank(&rowbuf_p[tid], owner_p);
      const uint32_t xr = *cluster.map_shared_rank(&rowbuf_r[tid], owner_r);
      const uint32_t pinv = (uint32_t)modinv_u64(pv, P);  // batched: one inverse per pivot, on the device
      u_s[tid] = tid >= jj ? mulmod_u32(xp, pinv, mp) : 0u;
      oldr_s[tid] = xr;
      if (tid == 0) pinv_s = pinv;
    }
    __syncthreads();
    // ---- phase C: swap rows r <-> p, scale, eliminate (every thread owns fixed rows of the panel)
    const int t = r;  // pivot number == row of U
    for (int q = tid; q < my_n; q += PANEL_THREADS) {
      const int i = my_lo + q;
      if (i < r) continue;
      if (i == r) {
        for (int c = jj; c < w; ++c) panel[c * rows_c + q] = u_s[c];
        Lm[(int64_t)t * ldl + i] = pv;
      } else {
        const bool isp = (i == p);
        const uint32_t l = isp ? oldr_s[jj] : panel[jj * rows_c + q];
        Lm[(int64_t)t * ldl + i] = l;
        panel[jj * rows_c + q] = 0;
        if (l != 0 || isp) {
          for (int c = jj + 1; c < w; ++c) {
            const uint32_t a = isp ? oldr_s[c] : panel[c * rows_c + q];
            panel[c * rows_c + q] = submod_u32(a, mulmod_u32(l, u_s[c], mp), P);
          }
        }
      }
    }
    if (rank == 0) {
      // earlier L columns of THIS panel follow the row swap (reference swap_rows on d_L, pluq_kernels.jl:280-289)
      if (p != r && tid < r - rb) {
        uint32_t* col = Lm + (int64_t)(rb + tid) * ldl;
        const uint32_t x = col[r], y = col[p];
        col[r] = y;
        col[p] = x;
      }
      if (tid == 0) {
        b.pivcol[t] = j0 + jj;
        b.pinv[t] = pinv_s;
        b.swp[t] = p;
      }
    }
    ++r;
  }
  __syncthreads();
  for (int c = 0; c < w; ++c)
    for (int q = tid; q < my_n; q += PANEL_THREADS) W[(int64_t)(j0 + c) * ldw + my_lo + q] = panel[c * rows_c + q];

  // ---- compose this panel's transpositions (rb..r-1) into one gather list (warp 0 of CTA 0)
  if (rank == 0 && warp == 0) {
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
  cluster.sync();  // all CTAs read st->r long ago; order the final write after every CTA's last use
  if (rank == 0 && tid == 0) b.st->r = r;
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
      for (int c = tc; c < 32; c += 2) {
        if (c0 + c >= c_hi) break;
        uint32_t* dst = W + (int64_t)(c0 + c) * ldw + i0 + ti;
        uint64_t acc = 0;
        if (big_mod) {
          for (int t = 0; t < k; ++t) acc = mod_u64(acc + mod_u64((uint64_t)sL[t][ti] * sU[t][c], mp), mp);
        } else {
          for (int t = 0; t < k; ++t) acc += (uint64_t)sL[t][ti] * sU[t][c];
          acc = mod_u64(acc, mp);
        }
        *dst = submod_u32(*dst, (uint32_t)acc, P);
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
    uint32_t d = unit_diag ? 1u : (uint32_t)modinv_u64(sT[j][j], P);
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

int32_t launch_panel(gffm_ctx* ctx, gffm_mat* W, gffm_mat* L, int j0, int w, int rows_c_max, const PluqBufs& b, const ModP& mp) {
  static bool attr_set = false;
  if (!attr_set) {
    GFFM_CUDA(cudaFuncSetAttribute(pluq_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM_BUDGET));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(PANEL_CLUSTER);
  cfg.blockDim = dim3(PANEL_THREADS);
  cfg.dynamicSmemBytes = (size_t)w * rows_c_max * 4;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = PANEL_CLUSTER;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  GFFM_CUDA(cudaLaunchKernelEx(&cfg, pluq_panel_kernel, W->data, W->ld, (int)W->rows, j0, w, L->data, L->ld, b, mp));
  ctx->launches++;
  return GFFM_OK;
}

int32_t triinv_views(gffm_ctx* ctx, MatView T, MatView X, bool upper, bool unit_diag, uint64_t P, int* singular_dev);

// Row-echelon elimination of A (copy) with L; leaves everything on the device, returns rank/pivots on the host.
int32_t eliminate(gffm_mat* A, Elim* out) {
  gffm_ctx* ctx = A->ctx;
  const int m = (int)A->rows, n = (int)A->cols;
  const uint64_t N = A->N;
  if (N >= (1ull << 32)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "elimination needs N < 2^32");
  gffm_mat *W = nullptr, *L = nullptr;
  GFFM_TRY(gffm_mat_create(ctx, m, n, N, A->pad, &W));
  GFFM_TRY(gffm_mat_create(ctx, m, m, N, A->pad, &L));
  out->W = W;
  out->L = L;
  GFFM_TRY(gffm_copy_views(ctx, view_of(W), view_of(A)));
  const int maxr = std::min(m, n);
  out->rank = 0;
  if (maxr == 0) return GFFM_OK;
  // device bookkeeping
  const size_t need = sizeof(PluqState) + (size_t)maxr * (sizeof(int) * 2 + sizeof(uint32_t)) + 2 * 64 * sizeof(int) + 256;
  GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_misc2, need));
  char* base = (char*)ctx->ws_misc2.ptr;
  PluqBufs b;
  b.st = (PluqState*)base;
  b.pivcol = (int*)(base + 64);
  b.swp = b.pivcol + maxr;
  b.pinv = (uint32_t*)(b.swp + maxr);
  b.g_dst = (int*)(b.pinv + maxr);
  b.g_src = b.g_dst + 64;
  GFFM_CUDA(cudaMemsetAsync(base, 0, need, ctx->stream));
  const ModP mp = make_modp(N);
  const int big_mod = N > (1ull << 29) ? 1 : 0;
  const int NB = 256;
  int r0 = 0;
  for (int c0 = 0; c0 < n && r0 < m; c0 += NB) {
    const int nbc = std::min(NB, n - c0);
    const int rows_c_max = (int)ceil_div(m - r0, PANEL_CLUSTER);
    int w = PANEL_SMEM_BUDGET / (4 * std::max(rows_c_max, 1));
    w = std::min(w, 16);
    if (w < 1) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "matrix has too many rows (%d) for the panel kernel", m);
    for (int j0 = c0; j0 < c0 + nbc; j0 += w) {
      const int wj = std::min(w, c0 + nbc - j0);
      GFFM_TRY(launch_panel(ctx, W, L, j0, wj, rows_c_max, b, mp));
      // row swaps of this panel: W columns right of the panel, all earlier L columns [0, rb)
      if (j0 + wj < n) {
        const int ncol = n - (j0 + wj);
        gather_rows_kernel<<<(unsigned)std::min<int64_t>(ceil_div(ncol, 8), 4 * ctx->num_sms), 256, 0, ctx->stream>>>(
            W->data, W->ld, j0 + wj, n, nullptr, b);
        GFFM_LAUNCH_CHECK(ctx);
      }
      if (r0 > 0 || j0 > c0) {
        gather_rows_kernel<<<(unsigned)std::min<int64_t>(ceil_div(std::max(j0, 1), 8), 4 * ctx->num_sms), 256, 0, ctx->stream>>>(
            L->data, L->ld, 0, 0, &b.st->rb, b);
        GFFM_LAUNCH_CHECK(ctx);
      }
      // rest of the outer block: U12' = L11'^-1 W12', W22' -= L21' U12'
      const int c_lo = j0 + wj, c_hi = c0 + nbc;
      if (c_lo < c_hi) {
        trsm_small_kernel<<<(unsigned)ceil_div(c_hi - c_lo, 8), 256, 0, ctx->stream>>>(W->data, W->ld, c_lo, c_hi, L->data, L->ld, b, mp);
        GFFM_LAUNCH_CHECK(ctx);
        dim3 grid((unsigned)ceil_div(m - r0, 128), (unsigned)std::min<int64_t>(ceil_div(c_hi - c_lo, 32), 8));
        update_small_kernel<<<grid, 256, 0, ctx->stream>>>(W->data, W->ld, m, c_lo, c_hi, L->data, L->ld, b, big_mod, mp);
        GFFM_LAUNCH_CHECK(ctx);
      }
    }
    // pivots found in this outer block
    PluqState hs;
    GFFM_CUDA(cudaMemcpyAsync(&hs, b.st, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
    GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
    const int K = hs.r - r0;
    const int c1 = c0 + nbc;
    if (K > 0 && c1 < n) {
      // U12 = L11^-1 * W[r0..r0+K, c1..n)
      gffm_mat* Linv = nullptr;
      GFFM_TRY(gffm_mat_create(ctx, K, K, N, 0, &Linv));
      int32_t st = triinv_views(ctx, sub_view(view_of(L), r0, r0, K, K), view_of(Linv), /*upper=*/false, false, N, nullptr);
      gffm_mat* tmp = nullptr;
      if (st == GFFM_OK) st = gffm_mat_create(ctx, K, n - c1, N, 0, &tmp);
      if (st == GFFM_OK) st = gffm_gemm_views(ctx, view_of(tmp), view_of(Linv), sub_view(view_of(W), r0, c1, K, n - c1), N, N, GFFM_GEMM_STORE, GFFM_ALGO_AUTO);
      if (st == GFFM_OK) st = gffm_copy_views(ctx, sub_view(view_of(W), r0, c1, K, n - c1), view_of(tmp));
      // W22 -= L21 * U12
      if (st == GFFM_OK && hs.r < m)
        st = gffm_gemm_views(ctx, sub_view(view_of(W), hs.r, c1, m - hs.r, n - c1), sub_view(view_of(L), hs.r, r0, m - hs.r, K),
                             sub_view(view_of(W), r0, c1, K, n - c1), N, N, GFFM_GEMM_SUB, GFFM_ALGO_AUTO);
      gffm_mat_destroy(Linv);
      if (tmp) gffm_mat_destroy(tmp);
      GFFM_TRY(st);
    }
    r0 = hs.r;
  }
  out->rank = r0;
  out->pivcol.resize(r0);
  out->swp.resize(r0);
  if (r0 > 0) {
    GFFM_CUDA(cudaMemcpyAsync(out->pivcol.data(), b.pivcol, sizeof(int) * r0, cudaMemcpyDeviceToHost, ctx->stream));
    GFFM_CUDA(cudaMemcpyAsync(out->swp.data(), b.swp, sizeof(int) * r0, cudaMemcpyDeviceToHost, ctx->stream));
    GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return GFFM_OK;
}

// triangular inverse on views: X = T^-1 (T n x n lower or upper triangular; strict other half ignored, X's other
// half must be zero on entry).  Recursive halving at multiples of TRI_B; combine with two GEMMs:
//   lower: X21 = -X22 * T21 * X11        upper: X12 = -X11 * T12 * X22
// (reference triangular_inverse_no_copy.jl:153-157,182-186,406-410,435-439 -- there with unguarded cuBLAS float GEMM)
int32_t triinv_rec(gffm_ctx* ctx, MatView T, MatView X, bool upper, uint64_t P, int lo, int hi, gffm_mat* scratch) {
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
    MatView tmp = sub_view(view_of(scratch), 0, 0, n2, n1);
    GFFM_TRY(gffm_gemm_views(ctx, tmp, T21, X11, P, P, GFFM_GEMM_STORE, GFFM_ALGO_AUTO));
    GFFM_TRY(gffm_fill_view(ctx, X21, 0));
    GFFM_TRY(gffm_gemm_views(ctx, X21, X22, tmp, P, P, GFFM_GEMM_SUB, GFFM_ALGO_AUTO));
  } else {
    MatView T12 = sub_view(T, lo, mid, n1, n2), X11 = sub_view(X, lo, lo, n1, n1), X22 = sub_view(X, mid, mid, n2, n2),
            X12 = sub_view(X, lo, mid, n1, n2);
    MatView tmp = sub_view(view_of(scratch), 0, 0, n1, n2);
    GFFM_TRY(gffm_gemm_views(ctx, tmp, T12, X22, P, P, GFFM_GEMM_STORE, GFFM_ALGO_AUTO));
    GFFM_TRY(gffm_fill_view(ctx, X12, 0));
    GFFM_TRY(gffm_gemm_views(ctx, X12, X11, tmp, P, P, GFFM_GEMM_SUB, GFFM_ALGO_AUTO));
  }
  return GFFM_OK;
}

int32_t triinv_views(gffm_ctx* ctx, MatView T, MatView X, bool upper, bool unit_diag, uint64_t P, int* singular_dev) {
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
  gffm_mat* scratch = nullptr;
  const int half = (int)round_up((n + 1) / 2, TRI_B);
  GFFM_TRY(gffm_mat_create(ctx, half, half, P, 0, &scratch));
  int32_t st = triinv_rec(ctx, T, X, upper, P, 0, n, scratch);
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
  if (!A || !out) GFFM_FAIL(GFFM_ERR_INVALID, "null");
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
  if (!A || (n_pairs > 0 && !pairs)) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (n_pairs <= 0) return GFFM_OK;
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
