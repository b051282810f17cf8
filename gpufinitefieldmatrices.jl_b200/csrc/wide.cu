// wide.cu -- uint64 residue storage for moduli 2^32 < N <= 2^52 (the reference container accepts N <= 2^52,
// src/CuModMatrix/CuModMatrix.jl:55-59; it stores float64-encoded integers, here exact uint64).  Covers the container
// (constructor with floored mod and exactness check, Array / unsafe_Array, zeros / eye / fill! / copy! / getindex / setindex! /
// change_modulus, rand) and the elementwise API (kernel_ops/{add,sub,mul,div,mod}_ops.jl with the mod_N override) with 128-bit
// intermediate products.  Products and eliminations above 2^32 go through KaratsubaMatrix (two 32-bit limbs) as in the reference.
#include <math.h>
#include <algorithm>
#include <type_traits>
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned long long mulmod_u64(unsigned long long a, unsigned long long b, unsigned long long P) {
  // a, b < P <= 2^52: the 104-bit product is reduced through its high and low words (hi < 2^40)
  const unsigned long long hi = __umul64hi(a, b), lo = a * b;
  unsigned long long acc = hi % P;
#pragma unroll
  for (int k = 0; k < 64; k += 8) acc = ((acc << 8) | ((lo >> (56 - k)) & 0xFFull)) % P;  // acc < 2^52 -> acc << 8 < 2^60
  return acc;
}

template <typename T>
__global__ void wide_upload_kernel(const T* __restrict__ src, int64_t lds, unsigned long long* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols,
                                   unsigned long long N, int do_mod, int* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i >= rows || j >= cols) return;
  const T x = src[j * lds + i];
  long long v;
  bool ok = true;
  if constexpr (std::is_integral<T>::value) {
    v = (long long)x;
  } else {
    const double d = (double)x;
    ok = (d == floor(d)) && fabs(d) <= 9.0e18;
    v = ok ? (long long)d : 0;
  }
  unsigned long long r;
  if (do_mod) {
    long long t = v % (long long)N;
    if (t < 0) t += (long long)N;
    r = (unsigned long long)t;
  } else {
    ok = ok && v >= 0;
    r = (unsigned long long)v;
  }
  if (!ok) atomicExch(bad, 1);
  dst[j * ldd + i] = r;
}

template <typename T>
__global__ void wide_download_kernel(const unsigned long long* __restrict__ src, int64_t lds, T* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols,
                                     int64_t src_rows, int64_t src_cols) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i >= rows || j >= cols) return;
  dst[j * ldd + i] = (i < src_rows && j < src_cols) ? (T)src[j * lds + i] : (T)0;
}

__global__ void __launch_bounds__(256)
wide_ewise_kernel(int op, unsigned long long* __restrict__ C, int64_t ldc, const unsigned long long* __restrict__ A, int64_t lda,
                  const unsigned long long* __restrict__ B, int64_t ldb, int64_t rows, int64_t cols, unsigned long long s, unsigned long long P) {
  const int64_t total = rows * cols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx / rows, i = idx - j * rows;
    const unsigned long long a = A[j * lda + i] % P;
    const unsigned long long b = B ? B[j * ldb + i] % P : 0ull;
    unsigned long long r;
    switch (op) {
      case GFFM_EW_MOD: r = a; break;
      case GFFM_EW_ADD: r = a + b; if (r >= P) r -= P; break;
      case GFFM_EW_SUB: r = a >= b ? a - b : a + P - b; break;
      case GFFM_EW_MUL: r = mulmod_u64(a, b, P); break;
      case GFFM_EW_SADD: r = a + s; if (r >= P) r -= P; break;
      case GFFM_EW_SSUB: r = a >= s ? a - s : a + P - s; break;
      case GFFM_EW_RSSUB: r = s >= a ? s - a : s + P - a; break;
      default: r = mulmod_u64(a, s, P); break;  // SMUL, SDIV (s already inverted)
    }
    C[j * ldc + i] = r;
  }
}

__global__ void wide_fill_kernel(unsigned long long* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols, unsigned long long v, int eye) {
  const int64_t total = rows * cols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx / rows, i = idx - j * rows;
    dst[j * ldd + i] = eye ? (i == j ? v : 0ull) : v;
  }
}

__global__ void wide_synth_kernel(unsigned long long* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols, uint64_t seed, unsigned long long N) {
  const int64_t total = rows * cols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx / rows, i = idx - j * rows;
    dst[j * ldd + i] = splitmix64(seed ^ (uint64_t)idx) % N;
  }
}

__global__ void wide_checksum_kernel(const unsigned long long* __restrict__ a, int64_t lda, const unsigned long long* __restrict__ b, int64_t ldb, int64_t rows,
                                     int64_t cols, unsigned long long* __restrict__ out) {
  unsigned long long sum = 0, diff = 0;
  const int64_t total = rows * cols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx / rows, i = idx - j * rows;
    const unsigned long long va = a[j * lda + i];
    sum += va * splitmix64((uint64_t)idx);
    if (b && b[j * ldb + i] != va) diff++;
  }
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    diff += __shfl_xor_sync(0xffffffffu, diff, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out[0], sum);
    if (diff) atomicAdd(&out[1], diff);
  }
}

unsigned wide_grid(gffm_ctx* ctx, int64_t total) {
  int64_t blocks = ceil_div(total, 256 * 4);
  const int64_t cap = (int64_t)ctx->num_sms * 16;
  return (unsigned)std::max<int64_t>(1, std::min(blocks, cap));
}

size_t host_size(int dt) { return (dt == GFFM_F64 || dt == GFFM_I64) ? 8 : ((dt == GFFM_F32 || dt == GFFM_U32 || dt == GFFM_I32) ? 4 : 0); }

unsigned long long wide_scalar(int64_t s, uint64_t P) {
  long long r = (long long)(s % (long long)P);
  if (r < 0) r += (long long)P;
  return (unsigned long long)r;
}

}  // namespace

int32_t gffm_wide_upload(gffm_mat* m, const void* host, int32_t dtype, int64_t ld, int32_t do_mod) {
  gffm_ctx* ctx = m->ctx;
  const size_t es = host_size(dtype);
  if (!es) GFFM_FAIL(GFFM_ERR_INVALID, "bad dtype %d", dtype);
  const int64_t slab_cols = std::max<int64_t>(1, std::min<int64_t>(m->cols, (int64_t)((256ull << 20) / (es * (size_t)m->rows))));
  GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_misc, (size_t)slab_cols * m->rows * es + 256));
  int* bad = (int*)((((uintptr_t)ctx->ws_misc.ptr + (size_t)slab_cols * m->rows * es) + 15) & ~(uintptr_t)15);
  GFFM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
  for (int64_t c0 = 0; c0 < m->cols; c0 += slab_cols) {
    const int64_t nc = std::min<int64_t>(slab_cols, m->cols - c0);
    GFFM_CUDA(cudaMemcpy2DAsync(ctx->ws_misc.ptr, (size_t)m->rows * es, (const char*)host + (size_t)c0 * ld * es, (size_t)ld * es, (size_t)m->rows * es, (size_t)nc,
                                cudaMemcpyHostToDevice, ctx->stream));
    for (int64_t s0 = 0; s0 < nc; s0 += 65535) {
      const int64_t snc = std::min<int64_t>(65535, nc - s0);
      dim3 grid((unsigned)ceil_div(m->rows, 256), (unsigned)snc);
      unsigned long long* dst = (unsigned long long*)m->data64 + (c0 + s0) * m->ld;
      const char* src = (const char*)ctx->ws_misc.ptr + (size_t)s0 * m->rows * es;
      switch (dtype) {
        case GFFM_F32: wide_upload_kernel<float><<<grid, 256, 0, ctx->stream>>>((const float*)src, m->rows, dst, m->ld, m->rows, snc, m->N, do_mod, bad); break;
        case GFFM_F64: wide_upload_kernel<double><<<grid, 256, 0, ctx->stream>>>((const double*)src, m->rows, dst, m->ld, m->rows, snc, m->N, do_mod, bad); break;
        case GFFM_I64: wide_upload_kernel<long long><<<grid, 256, 0, ctx->stream>>>((const long long*)src, m->rows, dst, m->ld, m->rows, snc, m->N, do_mod, bad); break;
        case GFFM_U32: wide_upload_kernel<unsigned int><<<grid, 256, 0, ctx->stream>>>((const unsigned int*)src, m->rows, dst, m->ld, m->rows, snc, m->N, do_mod, bad); break;
        default: wide_upload_kernel<int><<<grid, 256, 0, ctx->stream>>>((const int*)src, m->rows, dst, m->ld, m->rows, snc, m->N, do_mod, bad); break;
      }
      GFFM_LAUNCH_CHECK(ctx);
    }
    if (c0 + nc < m->cols) GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  int hbad = 0;
  GFFM_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (hbad) GFFM_FAIL(GFFM_ERR_INEXACT, "InexactError: entry is not an integer representable in the target range");
  return GFFM_OK;
}

int32_t gffm_wide_download(gffm_mat* m, void* host, int32_t dtype, int64_t ld, int32_t with_padding) {
  gffm_ctx* ctx = m->ctx;
  const size_t es = host_size(dtype);
  if (!es) GFFM_FAIL(GFFM_ERR_INVALID, "bad dtype %d", dtype);
  if (dtype == GFFM_F32 || dtype == GFFM_U32 || dtype == GFFM_I32)
    GFFM_FAIL(GFFM_ERR_INEXACT, "InexactError: residues modulo N > 2^32 do not fit a 32-bit host type (use Float64 or Int64)");
  const int64_t rows = with_padding ? m->rows + m->pad : m->rows, cols = with_padding ? m->cols + m->pad : m->cols;
  const int64_t slab_cols = std::max<int64_t>(1, std::min<int64_t>(cols, (int64_t)((256ull << 20) / (es * (size_t)rows))));
  GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_misc, (size_t)slab_cols * rows * es));
  for (int64_t c0 = 0; c0 < cols; c0 += slab_cols) {
    const int64_t nc = std::min<int64_t>(slab_cols, cols - c0);
    for (int64_t s0 = 0; s0 < nc; s0 += 65535) {
      const int64_t snc = std::min<int64_t>(65535, nc - s0);
      dim3 grid((unsigned)ceil_div(rows, 256), (unsigned)snc);
      const unsigned long long* src = (const unsigned long long*)m->data64 + (c0 + s0) * m->ld;
      char* dst = (char*)ctx->ws_misc.ptr + (size_t)s0 * rows * es;
      // physical storage holds (ld x pcols) elements: the padded image is read where it exists and is zero elsewhere
      const int64_t src_rows = std::min<int64_t>(rows, m->ld), src_cols = std::max<int64_t>(0, std::min<int64_t>(snc, m->pcols - (c0 + s0)));
      if (dtype == GFFM_F64) wide_download_kernel<double><<<grid, 256, 0, ctx->stream>>>(src, m->ld, (double*)dst, rows, rows, snc, src_rows, src_cols);
      else wide_download_kernel<long long><<<grid, 256, 0, ctx->stream>>>(src, m->ld, (long long*)dst, rows, rows, snc, src_rows, src_cols);
      GFFM_LAUNCH_CHECK(ctx);
    }
    GFFM_CUDA(cudaMemcpy2DAsync((char*)host + (size_t)c0 * ld * es, (size_t)ld * es, ctx->ws_misc.ptr, (size_t)rows * es, (size_t)rows * es, (size_t)nc,
                                cudaMemcpyDeviceToHost, ctx->stream));
    GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return GFFM_OK;
}

int32_t gffm_wide_ewise(int op, gffm_mat* C, gffm_mat* A, gffm_mat* B, int64_t scalar, uint64_t P) {
  gffm_ctx* ctx = C->ctx;
  if (!C->wide || !A->wide || (B && !B->wide))
    GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "elementwise operation mixing uint64 (N > 2^32) and uint32 storage: change_modulus one side first");
  if (C->rows != A->rows || C->cols != A->cols || (B && (B->rows != A->rows || B->cols != A->cols)))
    GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "elementwise operands differ in size");
  if (P == 0 || P > (1ull << 52)) GFFM_FAIL(GFFM_ERR_MODULUS_TOO_LARGE, "Modulus is bigger than 2^52");
  if ((op == GFFM_EW_ADD || op == GFFM_EW_SUB || op == GFFM_EW_MUL) && !B) GFFM_FAIL(GFFM_ERR_INVALID, "binary op needs B");
  const int64_t total = C->rows * C->cols;
  if (total == 0) return GFFM_OK;
  unsigned long long s = wide_scalar(scalar, P);
  if (op == GFFM_EW_SDIV) {
    s = modinv_u64(s, P);
    if (s == 0 && P != 1) GFFM_FAIL(GFFM_ERR_INVALID, "scalar is not invertible mod %llu", (unsigned long long)P);
  }
  wide_ewise_kernel<<<wide_grid(ctx, total), 256, 0, ctx->stream>>>(op, (unsigned long long*)C->data64, C->ld, (const unsigned long long*)A->data64, A->ld,
                                                                  B ? (const unsigned long long*)B->data64 : nullptr, B ? B->ld : 0, C->rows, C->cols, s, P);
  GFFM_LAUNCH_CHECK(ctx);
  return GFFM_OK;
}

int32_t gffm_wide_fill(gffm_mat* m, int64_t value, int eye, int64_t r0, int64_t c0, int64_t nr, int64_t nc) {
  if (nr * nc == 0) return GFFM_OK;
  wide_fill_kernel<<<wide_grid(m->ctx, nr * nc), 256, 0, m->ctx->stream>>>((unsigned long long*)m->data64 + c0 * m->ld + r0, m->ld, nr, nc, wide_scalar(value, m->N), eye);
  GFFM_LAUNCH_CHECK(m->ctx);
  return GFFM_OK;
}

int32_t gffm_wide_synth(gffm_mat* m, uint64_t seed) {
  if (m->rows * m->cols == 0) return GFFM_OK;
  wide_synth_kernel<<<wide_grid(m->ctx, m->rows * m->cols), 256, 0, m->ctx->stream>>>((unsigned long long*)m->data64, m->ld, m->rows, m->cols, seed, m->N);
  GFFM_LAUNCH_CHECK(m->ctx);
  return GFFM_OK;
}

int32_t gffm_wide_copy(gffm_mat* dst, gffm_mat* src) {
  if (!dst->wide || !src->wide) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "copy! between uint64 (N > 2^32) and uint32 storage");
  if (dst->rows * dst->cols == 0) return GFFM_OK;
  GFFM_CUDA(cudaMemcpy2DAsync(dst->data64, (size_t)dst->ld * 8, src->data64, (size_t)src->ld * 8, (size_t)src->rows * 8, (size_t)src->cols, cudaMemcpyDeviceToDevice,
                              dst->ctx->stream));
  return GFFM_OK;
}

int32_t gffm_wide_get(gffm_mat* m, int64_t i, int64_t j, int64_t* value) {
  unsigned long long v = 0;
  GFFM_CUDA(cudaMemcpyAsync(&v, (unsigned long long*)m->data64 + j * m->ld + i, 8, cudaMemcpyDeviceToHost, m->ctx->stream));
  GFFM_CUDA(cudaStreamSynchronize(m->ctx->stream));
  *value = (int64_t)v;
  return GFFM_OK;
}

int32_t gffm_wide_checksum(gffm_mat* a, gffm_mat* b, unsigned long long out[2]) {
  gffm_ctx* ctx = a->ctx;
  GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_misc2, 64));
  unsigned long long* d = (unsigned long long*)ctx->ws_misc2.ptr;
  GFFM_CUDA(cudaMemsetAsync(d, 0, 16, ctx->stream));
  if (a->rows * a->cols > 0) {
    wide_checksum_kernel<<<wide_grid(ctx, a->rows * a->cols), 256, 0, ctx->stream>>>((const unsigned long long*)a->data64, a->ld,
                                                                                   b ? (const unsigned long long*)b->data64 : nullptr, b ? b->ld : 0, a->rows, a->cols, d);
    GFFM_LAUNCH_CHECK(ctx);
  }
  GFFM_CUDA(cudaMemcpyAsync(out, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
  GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return GFFM_OK;
}
