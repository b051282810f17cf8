// karatsuba.cu -- two-limb ("Karatsuba") matrices x = d1 + N1*d2 modulo M = N1*N2 (N2 | N1), and the literal
// reference-quirk PLUQ.
//
// KMatMul! (reference src/KaratsubaMatrix/KaratsubaMatrix.jl:133-204): three modular sub-products and a recombination
//   P1 = A1*B1            (reference: mod N1^2)        -> here reduced mod N1*N2 and split  C1 = . mod N1, carry = . div N1
//   P2 = (A1+A2)*(B1+B2)  (reference: mod (4N1)^2)     -> here mod N2; the limb add is FUSED into the GEMM prologue (the
//                                                         plane-split kernel reads both addends; reference kernel_1, KaratsubaKernels.jl:129-139)
//   P3 = A2*B2            (reference: mod N1^2)        -> here mod N2
//   C2 = (P2 - P1 - P3 + carry) mod N2                 (reference kernel_2, KaratsubaKernels.jl:141-158)
// The result equals (A*B) mod N1*N2 exactly as in the reference whenever N2 | N1 (its documented validity domain).
#include <algorithm>
#include "common.cuh"

int32_t gffm_gemm_tc_limb_ex(gffm_ctx* ctx, MatView Cv, MatView A, const MatView* A2, MatView B, const MatView* B2, uint64_t R,
                             uint64_t P, int mode);
int32_t gffm_gemm_tc_rns_ex(gffm_ctx* ctx, MatView Cv, MatView A, const MatView* A2, MatView B, const MatView* B2, uint64_t R,
                            uint64_t P, int mode, bool balanced, uint32_t* kara_hi, int64_t ldhi, uint64_t kara_N1);

namespace {

__global__ void __launch_bounds__(256)
kara_recombine_kernel(uint32_t* __restrict__ C2, int64_t ld2, const uint32_t* __restrict__ C1, int64_t ld1,
                      const uint32_t* __restrict__ P2, const uint32_t* __restrict__ P3, const uint32_t* __restrict__ carry, int64_t ldt,
                      int64_t rows, int64_t cols, const __grid_constant__ ModP m2) {
  const int64_t total = rows * cols;
  const uint32_t N2 = (uint32_t)m2.P;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx / rows, i = idx - j * rows;
    const uint32_t p1 = (uint32_t)mod_u64(C1[j * ld1 + i], m2);  // P1 mod N2 == C1 mod N2 because N2 | N1
    const uint32_t p2 = P2[j * ldt + i], p3 = P3[j * ldt + i];
    const uint32_t cp = (uint32_t)mod_u64(carry[j * ldt + i], m2);
    uint32_t r = submod_u32(p2, p1, N2);
    r = submod_u32(r, p3, N2);
    r = addmod_u32(r, cp, N2);
    C2[j * ld2 + i] = r;
  }
}

// Karatsuba elementwise ops with carry between the limbs (reference KaratsubaKernels.jl:2-125)
__global__ void __launch_bounds__(256)
kara_ewise_kernel(int op, uint32_t* __restrict__ C1, uint32_t* __restrict__ C2, int64_t ldc, const uint32_t* __restrict__ A1,
                  const uint32_t* __restrict__ A2, int64_t lda, const uint32_t* __restrict__ B1, const uint32_t* __restrict__ B2,
                  int64_t ldb, int64_t rows, int64_t cols, unsigned long long s, unsigned long long N1, unsigned long long N2) {
  const int64_t total = rows * cols;
  const unsigned long long M = N1 * N2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx / rows, i = idx - j * rows;
    const unsigned long long a = (unsigned long long)A1[j * lda + i] + N1 * (unsigned long long)A2[j * lda + i];
    unsigned long long b = 0;
    if (B1) b = (unsigned long long)B1[j * ldb + i] + N1 * (unsigned long long)B2[j * ldb + i];
    unsigned long long r;
    switch (op) {
      case GFFM_EW_ADD: r = (a % M + b % M) % M; break;
      case GFFM_EW_SUB: r = (a % M + M - b % M) % M; break;
      case GFFM_EW_RSSUB: r = (M - a % M) % M; break;  // negate
      default: {  // SMUL: (a * s) mod M with a,s < M <= 2^52 -> 128-bit product
        const unsigned long long x = a % M, y = s % M;
        unsigned long long hi = __umul64hi(x, y), lo = x * y;
        // reduce (hi:lo) mod M bit-serially on the high word (hi < 2^40)
        unsigned long long acc = hi % M;
        for (int k = 0; k < 64; k += 8) {
          acc = ((acc << 8) | ((lo >> (56 - k)) & 0xFF)) % M;  // acc < 2^52 -> acc<<8 < 2^60
        }
        r = acc;
      }
    }
    C1[j * ldc + i] = (uint32_t)(r % N1);
    C2[j * ldc + i] = (uint32_t)(r / N1);
  }
}

}  // namespace

// C2 = (P2 - P1 - P3 + carry) mod N2 on raw buffers (used by the multi-GPU Karatsuba product, mg.cu)
int32_t gffm_kara_recombine(gffm_ctx* ctx, MatView vC2, MatView vC1, const uint32_t* P2, const uint32_t* P3, const uint32_t* carry, int64_t ldt,
                            uint64_t N2) {
  const int64_t m = vC1.rows, n = vC1.cols;
  if (m == 0 || n == 0) return GFFM_OK;
  int64_t blocks = ceil_div(m * n, 1024);
  if (blocks > (int64_t)ctx->num_sms * 16) blocks = (int64_t)ctx->num_sms * 16;
  kara_recombine_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(vC2.p, vC2.ld, vC1.p, vC1.ld, P2, P3, carry, ldt, m, n, make_modp(N2));
  GFFM_LAUNCH_CHECK(ctx);
  return GFFM_OK;
}

// one K-chunk (k <= 65536) of the product on views: (C1, C2) = (A1 + N1*A2) * (B1 + N1*B2) mod N1*N2
static int32_t kmat_mul_chunk(gffm_ctx* ctx, MatView vC1, MatView vC2, MatView vA1, MatView vA2, MatView vB1, MatView vB2, uint64_t N1,
                              uint64_t N2) {
  const int64_t m = vA1.rows, n = vB1.cols;
  // temporaries: carry, P2, P3 (m x n, same leading dimension)
  const int64_t ldt = round_up(m, 32);
  GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_misc, (size_t)3 * ldt * n * 4));
  uint32_t* carry = (uint32_t*)ctx->ws_misc.ptr;
  uint32_t* P2 = carry + ldt * n;
  uint32_t* P3 = P2 + ldt * n;
  MatView vP2{P2, ldt, m, n}, vP3{P3, ldt, m, n};
  // P1: exact integer product, reduced mod N1*N2 and split into (C1, carry)
  GFFM_TRY(gffm_gemm_tc_rns_ex(ctx, vC1, vA1, nullptr, vB1, nullptr, N1, N1 * N2, GFFM_GEMM_STORE, /*balanced=*/false, carry, ldt, N1));
  // P2: limb add fused into the plane split; inputs < N1 + N2 <= 2*N1
  const uint64_t R2 = N1 + N2;
  if (R2 <= 65536) GFFM_TRY(gffm_gemm_tc_limb_ex(ctx, vP2, vA1, &vA2, vB1, &vB2, R2, N2, GFFM_GEMM_STORE));
  else GFFM_TRY(gffm_gemm_tc_rns_ex(ctx, vP2, vA1, &vA2, vB1, &vB2, R2, N2, GFFM_GEMM_STORE, false, nullptr, 0, 0));
  // P3: inputs < N2, result mod N2 -> balanced residues allowed
  if (N2 <= 65536) GFFM_TRY(gffm_gemm_tc_limb_ex(ctx, vP3, vA2, nullptr, vB2, nullptr, N2, N2, GFFM_GEMM_STORE));
  else GFFM_TRY(gffm_gemm_tc_rns_ex(ctx, vP3, vA2, nullptr, vB2, nullptr, N2, N2, GFFM_GEMM_STORE, true, nullptr, 0, 0));
  int64_t blocks = ceil_div(m * n, 1024);
  if (blocks > (int64_t)ctx->num_sms * 16) blocks = (int64_t)ctx->num_sms * 16;
  kara_recombine_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(vC2.p, vC2.ld, vC1.p, vC1.ld, P2, P3, carry, ldt, m, n, make_modp(N2));
  GFFM_LAUNCH_CHECK(ctx);
  return GFFM_OK;
}

extern "C" int32_t gffm_kmat_mul(gffm_mat* C1, gffm_mat* C2, gffm_mat* A1, gffm_mat* A2, gffm_mat* B1, gffm_mat* B2, uint64_t N1,
                                 uint64_t N2) {
  GFFM_ENTER_MAT(C1);
  if (!C1 || !C2 || !A1 || !A2 || !B1 || !B2) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  GFFM_NARROW_ONLY(C1, C2, A1, A2, B1, B2);
  if (N1 == 0 || N2 == 0 || N1 % N2 != 0) GFFM_FAIL(GFFM_ERR_INVALID, "Karatsuba product requires N2 | N1");
  if (N1 > (1ull << 26) || N1 * N2 > (1ull << 52)) GFFM_FAIL(GFFM_ERR_MODULUS_TOO_LARGE, "N1 <= 2^26 and N1*N2 <= 2^52 required");
  const int64_t m = A1->rows, k = A1->cols, n = B1->cols;
  if (A2->rows != m || A2->cols != k || B1->rows != k || B2->rows != k || B2->cols != n || C1->rows != m || C1->cols != n ||
      C2->rows != m || C2->cols != n)
    GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "KMatMul!: inconsistent sizes");
  for (gffm_mat* c : {C1, C2})
    for (gffm_mat* x : {A1, A2, B1, B2})
      if (c == x) GFFM_FAIL(GFFM_ERR_INVALID, "KMatMul!: C must not alias an operand");
  gffm_ctx* ctx = A1->ctx;
  if (m == 0 || n == 0) return GFFM_OK;
  gffm_touch(C1);
  gffm_touch(C2);
  // mat x vec (the only Karatsuba product the reference's tests run, KMatMul_gemv! KaratsubaMatrix.jl:238-300): one HBM pass
  if (n == 1 && A1->ld == A2->ld) return gffm_kmat_gemv(ctx, C1, C2, A1, A2, B1, B2, N1, N2);
  const int64_t kmax = 65536;  // one int32 accumulation chunk of the RNS sub-products
  if (k <= kmax) return kmat_mul_chunk(ctx, view_of(C1), view_of(C2), view_of(A1), view_of(A2), view_of(B1), view_of(B2), N1, N2);
  // K-chunking: every chunk is a complete two-limb product mod N1*N2; chunks are summed with the two-limb carry add
  // (reference KaratsubaKernels.jl:2-31 add kernel)
  const int64_t ldt = round_up(m, 32);
  uint32_t* T = nullptr;
  if (gffm_dev_alloc(ctx, (void**)&T, (size_t)2 * ldt * n * 4) != cudaSuccess) GFFM_FAIL(GFFM_ERR_OOM, "KMatMul!: chunk temporaries");
  int32_t st = GFFM_OK;
  for (int64_t k0 = 0; k0 < k && st == GFFM_OK; k0 += kmax) {
    const int64_t kc = std::min(kmax, k - k0);
    MatView a1 = sub_view(view_of(A1), 0, k0, m, kc), a2 = sub_view(view_of(A2), 0, k0, m, kc);
    MatView b1 = sub_view(view_of(B1), k0, 0, kc, n), b2 = sub_view(view_of(B2), k0, 0, kc, n);
    if (k0 == 0) {
      st = kmat_mul_chunk(ctx, view_of(C1), view_of(C2), a1, a2, b1, b2, N1, N2);
    } else {
      MatView t1{T, ldt, m, n}, t2{T + ldt * n, ldt, m, n};
      st = kmat_mul_chunk(ctx, t1, t2, a1, a2, b1, b2, N1, N2);
      if (st != GFFM_OK) break;
      int64_t blocks = ceil_div(m * n, 1024);
      if (blocks > (int64_t)ctx->num_sms * 16) blocks = (int64_t)ctx->num_sms * 16;
      if (C1->ld != C2->ld) {
        st = GFFM_ERR_UNSUPPORTED;
        gffm_set_error("limb pairs must share a leading dimension");
        break;
      }
      kara_ewise_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(GFFM_EW_ADD, C1->data, C2->data, C1->ld, C1->data, C2->data, C1->ld, t1.p, t2.p, ldt,
                                                                  m, n, 0ull, N1, N2);
      ctx->launches++;
      if (cudaGetLastError() != cudaSuccess) {
        st = GFFM_ERR_CUDA;
        gffm_set_error("kara_ewise_kernel launch failed");
      }
    }
  }
  gffm_dev_free(ctx, T);
  return st;
}

extern "C" int32_t gffm_kmat_ewise(int32_t op, gffm_mat* C1, gffm_mat* C2, gffm_mat* A1, gffm_mat* A2, gffm_mat* B1, gffm_mat* B2,
                                   int64_t scalar, uint64_t N1, uint64_t N2) {
  GFFM_ENTER_MAT(C1);
  if (!C1 || !C2 || !A1 || !A2) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  GFFM_NARROW_ONLY(C1, C2, A1, A2, B1, B2);
  if (op != GFFM_EW_ADD && op != GFFM_EW_SUB && op != GFFM_EW_SMUL && op != GFFM_EW_RSSUB) GFFM_FAIL(GFFM_ERR_INVALID, "bad Karatsuba op");
  if ((op == GFFM_EW_ADD || op == GFFM_EW_SUB) && (!B1 || !B2)) GFFM_FAIL(GFFM_ERR_INVALID, "binary op needs B");
  if (N1 == 0 || N2 == 0 || N1 * N2 > (1ull << 52) || N1 > (1ull << 32) || N2 > (1ull << 32)) GFFM_FAIL(GFFM_ERR_MODULUS_TOO_LARGE, "bad Karatsuba moduli");
  const int64_t m = A1->rows, n = A1->cols;
  if (A2->rows != m || A2->cols != n || C1->rows != m || C1->cols != n || C2->rows != m || C2->cols != n ||
      (B1 && (B1->rows != m || B1->cols != n || B2->rows != m || B2->cols != n)))
    GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "Karatsuba elementwise: sizes differ");
  if (C1->ld != C2->ld || A1->ld != A2->ld || (B1 && B1->ld != B2->ld)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "limb pairs must share a leading dimension");
  if (m * n == 0) return GFFM_OK;
  gffm_touch(C1);
  gffm_touch(C2);
  gffm_ctx* ctx = A1->ctx;
  const unsigned long long M = N1 * N2;
  long long sv = scalar % (long long)M;
  if (sv < 0) sv += (long long)M;
  int64_t blocks = ceil_div(m * n, 1024);
  if (blocks > (int64_t)ctx->num_sms * 16) blocks = (int64_t)ctx->num_sms * 16;
  kara_ewise_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(op, C1->data, C2->data, C1->ld, A1->data, A2->data, A1->ld,
                                                              B1 ? B1->data : nullptr, B2 ? B2->data : nullptr, B1 ? B1->ld : 0, m, n,
                                                              (unsigned long long)sv, N1, N2);
  GFFM_LAUNCH_CHECK(ctx);
  return GFFM_OK;
}

// ---------------------------------------------------------------------------------------------------
// literal reference PLUQ (GFFM_PIVOT_REFERENCE_QUIRK): the reference's own loop, pivot by pivot, including the
// swap-with-fixed-last-column-then-skip behaviour for an all-zero pivot column (pluq_kernels.jl:88,103,193) and L
// written at column `col` (:343,:389).  One persistent CTA, no host round trips.  Kept for bit-for-bit parity
// experiments on rank-deficient inputs; the blocked path (pluq.cu) is the product.
// ---------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ bool better(uint32_t v, int i, uint32_t bv, int bi) { return v > bv || (v == bv && i < bi); }
struct QuirkOut {
  int n_prow, n_pcol, rank, pad;
};

__global__ void __launch_bounds__(1024, 1)
pluq_quirk_kernel(uint32_t* __restrict__ W, int64_t ldw, uint32_t* __restrict__ Lm, int64_t ldl, int rows, int cols, int lcols,
                  long long* __restrict__ prow, long long* __restrict__ pcol, QuirkOut* __restrict__ out, const __grid_constant__ ModP mp) {
  __shared__ uint32_t red_val[32];
  __shared__ int red_idx[32];
  __shared__ uint32_t s_pv;
  __shared__ int s_pi;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t P = (uint32_t)mp.P;
  int row = 0, col = 0, n_prow = 0, n_pcol = 0;
  const int last = cols - 1;
  while (row < rows && col < cols) {
    // find_pivot
    uint32_t bv = 0;
    int bi = 0x7fffffff;
    for (int i = row + tid; i < rows; i += 1024) {
      const uint32_t v = W[(int64_t)col * ldw + i];
      if (v > bv) {
        bv = v;
        bi = i;
      }
    }
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
      for (int o = 16; o > 0; o >>= 1) {
        const uint32_t ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (lane == 0) {
        s_pv = bv;
        s_pi = bi;
      }
    }
    __syncthreads();
    const uint32_t pv = s_pv;
    const int prow_i = s_pi;
    __syncthreads();
    if (pv == 0) {
      // swap_cols(col, last) over all rows, record, skip the swapped-in column
      for (int i = tid; i < rows; i += 1024) {
        const uint32_t a = W[(int64_t)col * ldw + i], c = W[(int64_t)last * ldw + i];
        W[(int64_t)col * ldw + i] = c;
        W[(int64_t)last * ldw + i] = a;
      }
      if (tid == 0) {
        pcol[2 * n_pcol] = col + 1;
        pcol[2 * n_pcol + 1] = last + 1;
      }
      ++n_pcol;
      ++col;
      __syncthreads();
      continue;
    }
    const uint32_t pinv = modinv_u32(pv, P);
    // swap_rows_and_mod over all columns
    for (int c = tid; c < cols; c += 1024) {
      const uint32_t tmp = W[(int64_t)c * ldw + prow_i];
      W[(int64_t)c * ldw + prow_i] = W[(int64_t)c * ldw + row];
      W[(int64_t)c * ldw + row] = mulmod_u32(tmp, pinv, mp);
    }
    // swap_rows on L
    for (int c = tid; c < lcols; c += 1024) {
      const uint32_t a = Lm[(int64_t)c * ldl + row], d = Lm[(int64_t)c * ldl + prow_i];
      Lm[(int64_t)c * ldl + row] = d;
      Lm[(int64_t)c * ldl + prow_i] = a;
    }
    if (row != prow_i) {
      if (tid == 0) {
        prow[2 * n_prow] = row + 1;
        prow[2 * n_prow + 1] = prow_i + 1;
      }
      ++n_prow;
    }
    __syncthreads();
    // move_and_zero_out: L written at column `col`
    if (col < lcols) {
      if (tid == 0) Lm[(int64_t)col * ldl + row] = pv;
      for (int i = row + 1 + tid; i < rows; i += 1024) Lm[(int64_t)col * ldl + i] = W[(int64_t)col * ldw + i];
    }
    __syncthreads();
    // rank-1 update; multipliers come from the pivot column of W (== what was just copied to L)
    for (int c = col + 1; c < cols; ++c) {
      const uint32_t u = W[(int64_t)c * ldw + row];
      if (u == 0) continue;
      for (int i = row + 1 + tid; i < rows; i += 1024) {
        const uint32_t l = W[(int64_t)col * ldw + i];
        uint32_t* d = W + (int64_t)c * ldw + i;
        *d = submod_u32(*d, mulmod_u32(l, u, mp), P);
      }
    }
    __syncthreads();
    for (int i = row + 1 + tid; i < rows; i += 1024) W[(int64_t)col * ldw + i] = 0;
    __syncthreads();
    ++row;
    ++col;
  }
  if (tid == 0) {
    out->n_prow = n_prow;
    out->n_pcol = n_pcol;
    out->rank = row;
  }
}
}  // namespace

int32_t gffm_pluq_quirk(gffm_mat* A, gffm_mat** U, gffm_mat** L, int64_t* prow_pairs, int64_t* n_prow, int64_t* pcol_pairs,
                        int64_t* n_pcol, int64_t* rank) {
  gffm_ctx* ctx = A->ctx;
  const int m = (int)A->rows, n = (int)A->cols;
  GFFM_NARROW_ONLY(A);
  if (A->N >= (1ull << 32)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "N < 2^32 required");
  gffm_mat *W = nullptr, *Lm = nullptr;
  GFFM_TRY(gffm_mat_create(ctx, m, n, A->N, A->pad, &W));
  GFFM_TRY(gffm_mat_create(ctx, m, m, A->N, A->pad, &Lm));
  GFFM_TRY(gffm_copy_views(ctx, view_of(W), view_of(A)));
  const size_t np = (size_t)std::max(1, std::max(m, n));
  GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_misc2, 64 + 4 * np * sizeof(long long)));
  QuirkOut* dout = (QuirkOut*)ctx->ws_misc2.ptr;
  long long* dprow = (long long*)((char*)ctx->ws_misc2.ptr + 64);
  long long* dpcol = dprow + 2 * np;
  // L is rows x rows; a pivot whose column index is >= rows (only possible after skipped columns) has no L column
  const int lcols = m;
  if (m > 0 && n > 0) {
    pluq_quirk_kernel<<<1, 1024, 0, ctx->stream>>>(W->data, W->ld, Lm->data, Lm->ld, m, n, lcols, dprow, dpcol, dout, make_modp(A->N));
    GFFM_LAUNCH_CHECK(ctx);
  } else {
    GFFM_CUDA(cudaMemsetAsync(dout, 0, sizeof(QuirkOut), ctx->stream));
  }
  QuirkOut ho;
  GFFM_CUDA(cudaMemcpyAsync(&ho, dout, sizeof(ho), cudaMemcpyDeviceToHost, ctx->stream));
  GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (prow_pairs && ho.n_prow) GFFM_CUDA(cudaMemcpy(prow_pairs, dprow, sizeof(long long) * 2 * ho.n_prow, cudaMemcpyDeviceToHost));
  if (pcol_pairs && ho.n_pcol) GFFM_CUDA(cudaMemcpy(pcol_pairs, dpcol, sizeof(long long) * 2 * ho.n_pcol, cudaMemcpyDeviceToHost));
  if (n_prow) *n_prow = ho.n_prow;
  if (n_pcol) *n_pcol = ho.n_pcol;
  if (rank) *rank = ho.rank;
  *U = W;
  *L = Lm;
  return GFFM_OK;
}
