// gemv.cu -- modular matrix x vector products at the HBM roofline.
//
//   gffm_gemv       z = A*x mod P                       reference mul!(z,A,x;R,P): src/CuModMatrix/CuModMatrix.jl:816-836 and
//                                                       stripe_mul!(z,A,x;...): src/CuModMatrix/kernel_mul/stripe_mul.jl:82-168
//   gffm_kmat_gemv  (c1 + N1*c2) = (A1 + N1*A2) * (b1 + N1*b2) mod N1*N2
//                                                       reference KMatMul_gemv!: src/KaratsubaMatrix/KaratsubaMatrix.jl:238-300
//                                                       (kernel_1 + three batched cuBLAS GEMVs + kernel_2,
//                                                       src/KaratsubaMatrix/KaratsubaKernels.jl:129-158)
//
// The reference runs cuBLAS float GEMVs over K-stripes with a `mod` pass per stripe; the Karatsuba form first WRITES the limb
// sum A1+A2 (one extra matrix) and then reads three matrices.  Here the matrix is read exactly once with 128-bit loads (the
// Karatsuba form reads A1 and A2 once and forms A1+A2 in registers: 8 B per element pair instead of 4 passes), products are
// accumulated exactly in uint64 with a reduction every T terms (T from the input bound R: T*(R-1)^2 < 2^63), the K range is
// split across CTAs so that every SM has several CTAs with >= 8 independent 16-byte loads per thread in flight, and a tiny
// second stage sums the K-slices.  Algorithmic bytes: 4*m*k (plain), 8*m*k (Karatsuba).
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int GV_THREADS = 256;  // 8 warps: the warps of a CTA share a 128-row tile and interleave its columns
constexpr int GV_ROWS = 128;     // rows per CTA (4 per lane)
constexpr int GV_UNROLL = 8;     // columns in flight per warp

// MODE 0: accumulate raw products, reduce every T columns; MODE 1: (R-1)^2 does not fit 63 bits -> reduce every product
template <int MODE>
__global__ void __launch_bounds__(GV_THREADS)
gemv_kernel(uint32_t* __restrict__ out, int64_t out_stride, const uint32_t* __restrict__ A, int64_t lda, const uint32_t* __restrict__ x,
            int m, int k, int kper, int T, int vec_ok, const __grid_constant__ ModP mp) {
  __shared__ unsigned long long part[GV_THREADS / 32][GV_ROWS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i0 = blockIdx.x * GV_ROWS + lane * 4;
  const int kbeg = blockIdx.y * kper, kend = min(k, kbeg + kper);
  unsigned long long acc[4] = {0, 0, 0, 0};
  const bool full = vec_ok && (i0 + 3 < m);
  int since = 0;
  for (int kk = kbeg + w; kk < kend; kk += 8 * GV_UNROLL) {
    uint4 v[GV_UNROLL];
    uint32_t xv[GV_UNROLL];
#pragma unroll
    for (int u = 0; u < GV_UNROLL; ++u) {
      const int kc = kk + 8 * u;
      v[u] = make_uint4(0, 0, 0, 0);
      xv[u] = 0;
      if (kc < kend) {
        const uint32_t* col = A + (int64_t)kc * lda + i0;
        xv[u] = x[kc];
        if (full) {
          v[u] = *reinterpret_cast<const uint4*>(col);
        } else {
          if (i0 < m) v[u].x = col[0];
          if (i0 + 1 < m) v[u].y = col[1];
          if (i0 + 2 < m) v[u].z = col[2];
          if (i0 + 3 < m) v[u].w = col[3];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < GV_UNROLL; ++u) {
      const unsigned long long xs = xv[u];
      if (MODE == 0) {
        acc[0] += v[u].x * xs; acc[1] += v[u].y * xs; acc[2] += v[u].z * xs; acc[3] += v[u].w * xs;
      } else {
        acc[0] += mod_u64(v[u].x * xs, mp); acc[1] += mod_u64(v[u].y * xs, mp);
        acc[2] += mod_u64(v[u].z * xs, mp); acc[3] += mod_u64(v[u].w * xs, mp);
      }
    }
    since += GV_UNROLL;
    if (since + GV_UNROLL > T) {  // the next batch could overflow the exactness budget
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c] = mod_u64(acc[c], mp);
      since = 1;  // the reduced value counts as one term (< P <= (R-1)^2 whenever that matters)
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) part[w][lane * 4 + c] = mod_u64(acc[c], mp);
  __syncthreads();
  if (threadIdx.x < GV_ROWS) {
    unsigned long long s = 0;
#pragma unroll
    for (int ww = 0; ww < GV_THREADS / 32; ++ww) s += part[ww][threadIdx.x];
    const int i = blockIdx.x * GV_ROWS + threadIdx.x;
    if (i < m) out[(int64_t)blockIdx.y * out_stride + i] = (uint32_t)mod_u64(s, mp);
  }
}

__global__ void __launch_bounds__(256)
gemv_reduce_kernel(uint32_t* __restrict__ z, const uint32_t* __restrict__ part, int64_t stride, int ks, int m, const __grid_constant__ ModP mp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  unsigned long long s = 0;
  for (int t = 0; t < ks; ++t) s += part[(int64_t)t * stride + i];
  z[i] = (uint32_t)mod_u64(s, mp);
}

// Karatsuba mat x vec: one pass over A1 and A2, three exact dot products per row
//   s1 = sum A1*b1 (terms < N1^2)   s2 = sum (A1+A2)*(b1+b2) (terms < (N1+N2)^2)   s3 = sum A2*b2 (terms < N2^2)
// K-slice partials: s1 mod N1*N2 (uint64), s2 mod N2, s3 mod N2 (stored as uint64 for one layout).
__global__ void __launch_bounds__(GV_THREADS)
kgemv_kernel(unsigned long long* __restrict__ out, int64_t out_stride, const uint32_t* __restrict__ A1, const uint32_t* __restrict__ A2,
             int64_t lda, const uint32_t* __restrict__ b1, const uint32_t* __restrict__ b2, int m, int k, int kper, int T, int vec_ok,
             const __grid_constant__ ModP mM, const __grid_constant__ ModP m2) {
  __shared__ unsigned long long part[GV_THREADS / 32][GV_ROWS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i0 = blockIdx.x * GV_ROWS + lane * 4;
  const int kbeg = blockIdx.y * kper, kend = min(k, kbeg + kper);
  unsigned long long s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0}, s3[4] = {0, 0, 0, 0};
  const bool full = vec_ok && (i0 + 3 < m);
  constexpr int U = 4;
  int since = 0;
  for (int kk = kbeg + w; kk < kend; kk += 8 * U) {
    uint4 p[U], q[U];
    uint32_t x1[U], x2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int kc = kk + 8 * u;
      p[u] = q[u] = make_uint4(0, 0, 0, 0);
      x1[u] = x2[u] = 0;
      if (kc < kend) {
        const uint32_t* c1 = A1 + (int64_t)kc * lda + i0;
        const uint32_t* c2 = A2 + (int64_t)kc * lda + i0;
        x1[u] = b1[kc];
        x2[u] = b2[kc];
        if (full) {
          p[u] = *reinterpret_cast<const uint4*>(c1);
          q[u] = *reinterpret_cast<const uint4*>(c2);
        } else {
          if (i0 < m) { p[u].x = c1[0]; q[u].x = c2[0]; }
          if (i0 + 1 < m) { p[u].y = c1[1]; q[u].y = c2[1]; }
          if (i0 + 2 < m) { p[u].z = c1[2]; q[u].z = c2[2]; }
          if (i0 + 3 < m) { p[u].w = c1[3]; q[u].w = c2[3]; }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long y1 = x1[u], y3 = x2[u], y2 = (unsigned long long)x1[u] + x2[u];
      const uint32_t pa[4] = {p[u].x, p[u].y, p[u].z, p[u].w}, qa[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s1[c] += pa[c] * y1;
        s3[c] += qa[c] * y3;
        s2[c] += ((unsigned long long)pa[c] + qa[c]) * y2;
      }
    }
    since += U;
    if (since + U > T) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s1[c] = mod_u64(s1[c], mM);
        s2[c] = mod_u64(s2[c], m2);
        s3[c] = mod_u64(s3[c], m2);
      }
      since = 1;
    }
  }
  // cross-warp sums of the three accumulators (one shared tile, reused)
  for (int which = 0; which < 3; ++which) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      part[w][lane * 4 + c] = which == 0 ? mod_u64(s1[c], mM) : which == 1 ? mod_u64(s2[c], m2) : mod_u64(s3[c], m2);
    __syncthreads();
    if (threadIdx.x < GV_ROWS) {
      unsigned long long s = 0;
#pragma unroll
      for (int ww = 0; ww < GV_THREADS / 32; ++ww) s += part[ww][threadIdx.x];
      const int i = blockIdx.x * GV_ROWS + threadIdx.x;
      if (i < m) out[((int64_t)blockIdx.y * 3 + which) * out_stride + i] = which == 0 ? mod_u64(s, mM) : mod_u64(s, m2);
    }
    __syncthreads();
  }
}

// second stage: sum the K-slices and recombine (reference kernel_2, KaratsubaKernels.jl:141-158):
//   P1 = s1 mod N1*N2 -> c1 = P1 mod N1, carry = P1 div N1;  c2 = (P2 - P1 - P3 + carry) mod N2
__global__ void __launch_bounds__(256)
kgemv_finish_kernel(uint32_t* __restrict__ c1, uint32_t* __restrict__ c2, const unsigned long long* __restrict__ part, int64_t stride, int ks,
                    int m, unsigned long long N1, const __grid_constant__ ModP mM, const __grid_constant__ ModP m2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  unsigned long long t1 = 0, t2 = 0, t3 = 0;
  for (int t = 0; t < ks; ++t) {
    t1 += part[((int64_t)t * 3 + 0) * stride + i];
    t2 += part[((int64_t)t * 3 + 1) * stride + i];
    t3 += part[((int64_t)t * 3 + 2) * stride + i];
  }
  const unsigned long long P1 = mod_u64(t1, mM);
  const uint32_t N2 = (uint32_t)m2.P;
  const uint32_t P2 = (uint32_t)mod_u64(t2, m2), P3 = (uint32_t)mod_u64(t3, m2);
  const unsigned long long lo = P1 % N1, carry = P1 / N1;
  uint32_t r = submod_u32(P2, (uint32_t)mod_u64(lo, m2), N2);
  r = submod_u32(r, P3, N2);
  r = addmod_u32(r, (uint32_t)mod_u64(carry, m2), N2);
  c1[i] = (uint32_t)lo;
  c2[i] = r;
}

// how many columns one CTA covers: enough CTAs for >= 4 per SM, slices of at least 256 columns (8 warps x 32)
void gemv_shape(const gffm_ctx* ctx, int64_t m, int64_t k, int* ks, int* kper) {
  const int64_t row_tiles = ceil_div(m, GV_ROWS);
  int64_t want = ceil_div((int64_t)ctx->num_sms * 6, row_tiles);
  const int64_t max_slices = std::max<int64_t>(1, k / 256);
  want = std::max<int64_t>(1, std::min<int64_t>(want, std::min<int64_t>(max_slices, 64)));
  int64_t per = round_up(ceil_div(k, want), 8);
  *kper = (int)per;
  *ks = (int)ceil_div(k, per);
}

// terms of size < bound2 that a uint64 accumulator holding one reduced value can take
int terms_budget(long double bound2) {
  if (bound2 < 1) return 1 << 20;
  long double t = 9.0e18L / bound2;
  if (t > (1 << 20)) return 1 << 20;
  return (int)t;
}

}  // namespace

int32_t gffm_gemv_raw(gffm_ctx* ctx, uint32_t* z, const uint32_t* A, int64_t lda, const uint32_t* x, int64_t m, int64_t k, uint64_t R, uint64_t P) {
  if (m == 0) return GFFM_OK;
  if (P == 0 || P >= (1ull << 32)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "gemv needs 0 < P < 2^32");
  if (m >= (1ll << 31) || k >= (1ll << 31)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "gemv: dimension above 2^31");
  if (k == 0) {
    GFFM_CUDA(cudaMemsetAsync(z, 0, (size_t)m * 4, ctx->stream));
    return GFFM_OK;
  }
  if (R == 0 || R > (1ull << 32)) R = 1ull << 32;
  // exactness budget: a reduced accumulator (< P) plus T raw products (< (R-1)^2 each) must stay below 2^64
  const long double b2 = (long double)(R - 1) * (long double)(R - 1);
  const int budget = terms_budget(b2 > (long double)P ? b2 : (long double)P);
  const int mode = budget < 2 * GV_UNROLL ? 1 : 0;
  const int T = mode ? (1 << 20) : budget;  // MODE 1 adds values < P < 2^32: 2^20 of them are safe
  int ks, kper;
  gemv_shape(ctx, m, k, &ks, &kper);
  const int vec_ok = ((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (lda % 4) == 0) ? 1 : 0;
  dim3 grid((unsigned)ceil_div(m, GV_ROWS), (unsigned)ks);
  uint32_t* out = z;
  int64_t stride = 0;
  if (ks > 1) {
    stride = round_up(m, 32);
    GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_gemv, (size_t)ks * stride * 4));
    out = (uint32_t*)ctx->ws_gemv.ptr;
  }
  const ModP mp = make_modp(P);
  if (mode == 0) gemv_kernel<0><<<grid, GV_THREADS, 0, ctx->stream>>>(out, stride, A, lda, x, (int)m, (int)k, kper, T, vec_ok, mp);
  else gemv_kernel<1><<<grid, GV_THREADS, 0, ctx->stream>>>(out, stride, A, lda, x, (int)m, (int)k, kper, T, vec_ok, mp);
  GFFM_LAUNCH_CHECK(ctx);
  if (ks > 1) {
    gemv_reduce_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, ctx->stream>>>(z, out, stride, ks, (int)m, mp);
    GFFM_LAUNCH_CHECK(ctx);
  }
  return GFFM_OK;
}

extern "C" int32_t gffm_gemv(gffm_mat* z, gffm_mat* A, gffm_mat* x, uint64_t R, uint64_t P) {
  GFFM_ENTER_MAT(z);
  if (!z || !A || !x) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  GFFM_NARROW_ONLY(z, A, x);
  if (x->cols != 1 || z->cols != 1 || A->cols != x->rows || A->rows != z->rows)
    GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "gemv: A is %lldx%lld, x has %lld rows, z has %lld rows", (long long)A->rows, (long long)A->cols,
              (long long)x->rows, (long long)z->rows);
  if (!P) {
    if (A->N != x->N || A->N != z->N) GFFM_FAIL(GFFM_ERR_MODULUS_MISMATCH, "gemv operands have different moduli");
    P = z->N;
  }
  if (P >= (1ull << 32)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "gemv needs P < 2^32");
  // R bounds the STORED entries (reference kwarg R, stripe_mul.jl:96-114).  With a mod_P override the entries can be far larger
  // than P, so the default is the operands' own moduli, never P.
  if (!R) R = A->N > x->N ? A->N : x->N;
  if (z == x || z == A) GFFM_FAIL(GFFM_ERR_INVALID, "gemv: z must not alias an operand");
  gffm_touch(z);
  return gffm_gemv_raw(A->ctx, z->data, A->data, A->ld, x->data, A->rows, A->cols, R, P);
}

// Karatsuba mat x vec in one pass (called by gffm_kmat_mul for n == 1)
int32_t gffm_kmat_gemv(gffm_ctx* ctx, gffm_mat* C1, gffm_mat* C2, gffm_mat* A1, gffm_mat* A2, gffm_mat* B1, gffm_mat* B2, uint64_t N1,
                       uint64_t N2) {
  const int64_t m = A1->rows, k = A1->cols;
  if (m == 0) return GFFM_OK;
  if (A1->ld != A2->ld) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "limb pairs must share a leading dimension");
  if (m >= (1ll << 31) || k >= (1ll << 31)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "gemv: dimension above 2^31");
  const uint64_t M = N1 * N2;
  // largest term: (A1+A2)*(b1+b2) < (N1+N2)^2 <= 2^54; a reduced s1 is < M <= 2^52
  const long double b2 = (long double)(N1 + N2) * (long double)(N1 + N2);
  const int T = terms_budget(b2 > (long double)M ? b2 : (long double)M);
  if (T < 8) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "Karatsuba gemv: moduli too large for the uint64 accumulation budget");
  int ks, kper;
  gemv_shape(ctx, m, std::max<int64_t>(k, 1), &ks, &kper);
  const int64_t stride = round_up(m, 32);
  GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_gemv, (size_t)ks * 3 * stride * 8));
  unsigned long long* part = (unsigned long long*)ctx->ws_gemv.ptr;
  const int vec_ok = ((reinterpret_cast<uintptr_t>(A1->data) & 15) == 0 && (reinterpret_cast<uintptr_t>(A2->data) & 15) == 0 && (A1->ld % 4) == 0) ? 1 : 0;
  const ModP mM = make_modp(M), m2 = make_modp(N2);
  dim3 grid((unsigned)ceil_div(m, GV_ROWS), (unsigned)ks);
  kgemv_kernel<<<grid, GV_THREADS, 0, ctx->stream>>>(part, stride, A1->data, A2->data, A1->ld, B1->data, B2->data, (int)m, (int)k, kper, T, vec_ok, mM, m2);
  GFFM_LAUNCH_CHECK(ctx);
  kgemv_finish_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, ctx->stream>>>(C1->data, C2->data, part, stride, ks, (int)m, N1, mM, m2);
  GFFM_LAUNCH_CHECK(ctx);
  return GFFM_OK;
}
