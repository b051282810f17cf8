// api.cu -- context, container (CuModMatrix storage), host<->device conversion, elementwise kernels, the scalar
// SIMT GEMM/GEMV and the GEMM dispatcher of libgffm.  Reference counterparts are cited per function in
// include/gffm.h.
#include <stdarg.h>
#include <math.h>
#include <algorithm>
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void gffm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* gffm_last_error(void) { return g_err; }
extern "C" const char* gffm_version(void) { return "libgffm 0.1 (sm_100a; tcgen05 kind::i8 + TMA)"; }

extern "C" int32_t gffm_device_count(int32_t* count) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    cudaGetLastError();
    c = 0;
  }
  if (count) *count = c;
  return GFFM_OK;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
extern "C" int32_t gffm_create(int32_t device, gffm_ctx** out) {
  if (!out) GFFM_FAIL(GFFM_ERR_INVALID, "null out");
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess || c == 0) {
    cudaGetLastError();
    GFFM_FAIL(GFFM_ERR_NO_DEVICE, "no CUDA device: libgffm has no CPU fallback");
  }
  if (device < 0 || device >= c) GFFM_FAIL(GFFM_ERR_INVALID, "device %d out of range (%d devices)", device, c);
  GFFM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  GFFM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) GFFM_FAIL(GFFM_ERR_NO_DEVICE, "device %s is sm_%d%d; libgffm is built for sm_100a only", prop.name, prop.major, prop.minor);
  gffm_ctx* ctx = new gffm_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  GFFM_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->own_stream = true;
  {  // keep freed matrix / plane-cache blocks in the default pool instead of returning them to the driver at every sync
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  for (int i = 0; i < 8; ++i) GFFM_CUDA(cudaEventCreate(&ctx->ev[i]));
  ctx->trace = getenv("GFFM_TRACE") != nullptr && atoi(getenv("GFFM_TRACE")) != 0;
  *out = ctx;
  return GFFM_OK;
}

cudaError_t gffm_dev_alloc(gffm_ctx* ctx, void** ptr, size_t bytes) {
  ctx->alloc_bytes += (int64_t)bytes;
  ctx->alloc_calls++;
  cudaError_t e = cudaMallocAsync(ptr, bytes ? bytes : 4, ctx->stream);
  if (e != cudaSuccess) {  // pool fragmented or exhausted: hand cached blocks back to the driver and retry once
    cudaGetLastError();
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
      cudaStreamSynchronize(ctx->stream);
      cudaMemPoolTrimTo(pool, 0);
    }
    cudaGetLastError();
    e = cudaMallocAsync(ptr, bytes ? bytes : 4, ctx->stream);
    if (e != cudaSuccess) {
      cudaGetLastError();
      *ptr = nullptr;
    }
  }
  return e;
}
void gffm_dev_free(gffm_ctx* ctx, void* ptr) {
  if (!ptr) return;
  if (cudaFreeAsync(ptr, ctx->stream) != cudaSuccess) {
    cudaGetLastError();
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ptr);
  }
}

static void ws_free(gffm_workspace* ws) {
  if (ws->ptr) cudaFree(ws->ptr);
  ws->ptr = nullptr;
  ws->bytes = 0;
}

extern "C" int32_t gffm_destroy(gffm_ctx* ctx) {
  GFFM_ENTER_CTX(ctx);
  if (!ctx) return GFFM_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ws_free(&ctx->ws_planes_a);
  ws_free(&ctx->ws_planes_b);
  ws_free(&ctx->ws_eplanes);
  ws_free(&ctx->ws_misc);
  ws_free(&ctx->ws_misc2);
  ws_free(&ctx->ws_invtab);
  ws_free(&ctx->ws_scratch);
  ws_free(&ctx->ws_gemv);
  ws_free(&ctx->ws_host);
  if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
  if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
  if (ctx->s_aux) cudaStreamDestroy(ctx->s_aux);
  for (auto e : ctx->ev_pool) cudaEventDestroy(e);
  for (auto e : ctx->tile_events) cudaEventDestroy(e);
  if (ctx->ws_pinned.ptr) cudaFreeHost(ctx->ws_pinned.ptr);
  for (int i = 0; i < 8; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return GFFM_OK;
}

// prints the most recent trace records (GFFM_TRACE_LAST, default 160) as a timeline relative to the earliest of them and clears the trace
void gffm_trace_dump(gffm_ctx* ctx, int rank) {
  if (!ctx->trace || ctx->trace_recs.empty()) return;
  cudaDeviceSynchronize();
  const size_t last = getenv("GFFM_TRACE_LAST") ? (size_t)atoi(getenv("GFFM_TRACE_LAST")) : 160;
  const size_t n = ctx->trace_recs.size(), first = n > last ? n - last : 0;
  // reference: the earliest begin among the printed records
  size_t ref = first;
  for (size_t i = first; i < n; ++i) {
    float d = 0.f;
    if (cudaEventElapsedTime(&d, ctx->trace_recs[ref].a, ctx->trace_recs[i].a) == cudaSuccess && d < 0.f) ref = i;
  }
  static const char* snames[] = {"compute", "aux", "dist", "pull", "push", "comm"};
  for (size_t i = first; i < n; ++i) {
    const gffm_ctx::TraceRec& r = ctx->trace_recs[i];
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, ctx->trace_recs[ref].a, r.a);
    cudaEventElapsedTime(&b, ctx->trace_recs[ref].a, r.b);
    fprintf(stderr, "[gffm trace r%d] %-8s %-12s %4d  %9.3f -> %9.3f  (%.3f ms)\n", rank, snames[r.sid < 6 ? r.sid : 0], r.what, r.idx, a, b, b - a);
  }
  for (auto& r : ctx->trace_recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  ctx->trace_recs.clear();
  cudaGetLastError();
}

extern "C" int32_t gffm_sync(gffm_ctx* ctx) {
  GFFM_ENTER_CTX(ctx);
  if (!ctx) GFFM_FAIL(GFFM_ERR_INVALID, "null ctx");
  GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return GFFM_OK;
}

extern "C" int32_t gffm_set_stream(gffm_ctx* ctx, void* s) {
  GFFM_ENTER_CTX(ctx);
  if (!ctx) GFFM_FAIL(GFFM_ERR_INVALID, "null ctx");
  GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)s;
  ctx->own_stream = false;
  return GFFM_OK;
}
extern "C" int32_t gffm_get_stream(gffm_ctx* ctx, void** s) {
  GFFM_ENTER_CTX(ctx);
  if (!ctx || !s) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  *s = (void*)ctx->stream;
  return GFFM_OK;
}
extern "C" int32_t gffm_set_profiling(gffm_ctx* ctx, int32_t on) {
  GFFM_ENTER_CTX(ctx);
  if (!ctx) GFFM_FAIL(GFFM_ERR_INVALID, "null ctx");
  ctx->profile = on != 0;
  ctx->n_ev = 0;
  return GFFM_OK;
}
extern "C" int32_t gffm_last_timings(gffm_ctx* ctx, double* ms, int32_t cap, int32_t* n) {
  GFFM_ENTER_CTX(ctx);
  if (!ctx) GFFM_FAIL(GFFM_ERR_INVALID, "null ctx");
  ctx->timings.clear();
  if (ctx->n_ev == 0 && !ctx->elim_timings.empty()) {
    ctx->timings = ctx->elim_timings;
    ctx->elim_timings.clear();
  }
  if (ctx->n_ev == 0 && !ctx->tile_events.empty()) {
    // tiled product: {0, sum of the GEMM launch durations, number of launches}
    double sum = 0;
    GFFM_CUDA(cudaEventSynchronize(ctx->tile_events.back()));
    for (size_t i = 0; i + 1 < ctx->tile_events.size(); i += 2) {
      float t = 0.f;
      GFFM_CUDA(cudaEventElapsedTime(&t, ctx->tile_events[i], ctx->tile_events[i + 1]));
      sum += t;
    }
    ctx->timings = {0.0, sum, (double)(ctx->tile_events.size() / 2)};
  }
  if (ctx->n_ev >= 2) {
    GFFM_CUDA(cudaEventSynchronize(ctx->ev[ctx->n_ev - 1]));
    for (int i = 0; i + 1 < ctx->n_ev; ++i) {
      float t = 0.f;
      GFFM_CUDA(cudaEventElapsedTime(&t, ctx->ev[i], ctx->ev[i + 1]));
      ctx->timings.push_back((double)t);
    }
  }
  int32_t k = 0;
  for (; k < cap && k < (int32_t)ctx->timings.size(); ++k) ms[k] = ctx->timings[k];
  if (n) *n = k;
  return GFFM_OK;
}
extern "C" int32_t gffm_set_gemm_ctas(gffm_ctx* ctx, int32_t ctas) {
  GFFM_ENTER_CTX(ctx);
  if (!ctx || ctas < 0) GFFM_FAIL(GFFM_ERR_INVALID, "bad argument");
  ctx->gemm_ctas = ctas;
  return GFFM_OK;
}
extern "C" int32_t gffm_launch_count(gffm_ctx* ctx, int64_t* count) {
  GFFM_ENTER_CTX(ctx);
  if (!ctx || !count) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  *count = ctx->launches;
  return GFFM_OK;
}

extern "C" int32_t gffm_alloc_stats(gffm_ctx* ctx, int64_t* bytes, int64_t* calls) {
  GFFM_ENTER_CTX(ctx);
  if (!ctx) GFFM_FAIL(GFFM_ERR_INVALID, "null ctx");
  if (bytes) *bytes = ctx->alloc_bytes;
  if (calls) *calls = ctx->alloc_calls;
  return GFFM_OK;
}

int32_t gffm_ws_reserve(gffm_ctx* ctx, gffm_workspace* ws, size_t bytes) {
  if (ws->bytes >= bytes) return GFFM_OK;
  if (ws->ptr) {
    GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
    GFFM_CUDA(cudaFree(ws->ptr));
    ws->ptr = nullptr;
    ws->bytes = 0;
  }
  size_t want = bytes + bytes / 8 + 4096;
  ctx->alloc_bytes += (int64_t)want;
  ctx->alloc_calls++;
  cudaError_t e = cudaMalloc(&ws->ptr, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(&ws->ptr, want);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    ws->ptr = nullptr;
    GFFM_FAIL(GFFM_ERR_OOM, "workspace allocation of %zu bytes failed", bytes);
  }
  ws->bytes = want;
  return GFFM_OK;
}

int32_t gffm_pinned_reserve(gffm_ctx* ctx, size_t bytes) {
  gffm_workspace* ws = &ctx->ws_pinned;
  if (ws->bytes >= bytes) return GFFM_OK;
  if (ws->ptr) {
    GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFreeHost(ws->ptr);
    ws->ptr = nullptr;
    ws->bytes = 0;
  }
  GFFM_CUDA(cudaMallocHost(&ws->ptr, bytes));
  ws->bytes = bytes;
  return GFFM_OK;
}

// ---------------------------------------------------------------------------------------------
// container
// ---------------------------------------------------------------------------------------------
extern "C" int32_t gffm_mat_create(gffm_ctx* ctx, int64_t rows, int64_t cols, uint64_t N, int32_t pad, gffm_mat** out) {
  GFFM_ENTER_CTX(ctx);
  if (!ctx || !out) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (rows < 0 || cols < 0) GFFM_FAIL(GFFM_ERR_INVALID, "negative size");
  if (N > (1ull << 52)) GFFM_FAIL(GFFM_ERR_MODULUS_TOO_LARGE, "Modulus is bigger than 2^52");  // CuModMatrix.jl:55-59
  if (N == 0) GFFM_FAIL(GFFM_ERR_INVALID, "modulus must be positive");
  if (pad < 0) pad = GFFM_REF_PAD;
  gffm_mat* m = new gffm_mat();
  m->ctx = ctx;
  m->rows = rows;
  m->cols = cols;
  m->pad = pad;
  m->N = N;
  m->ld = round_up(rows + pad > 0 ? rows + pad : 1, 32);
  m->pcols = cols + pad > 0 ? cols + pad : 1;
  m->owned = true;
  cudaSetDevice(ctx->device);
  m->wide = N > (1ull << 32);  // uint64 residues (wide.cu)
  size_t bytes = (size_t)m->ld * m->pcols * (m->wide ? sizeof(uint64_t) : sizeof(uint32_t));
  void* mem = nullptr;
  cudaError_t e = gffm_dev_alloc(ctx, &mem, bytes);
  if (e != cudaSuccess) {
    delete m;
    GFFM_FAIL(GFFM_ERR_OOM, "device allocation of %zu bytes failed", bytes);
  }
  if (m->wide) m->data64 = mem;
  else m->data = (uint32_t*)mem;
  GFFM_CUDA(cudaMemsetAsync(mem, 0, bytes, ctx->stream));
  *out = m;
  return GFFM_OK;
}

extern "C" int32_t gffm_mat_wrap(gffm_ctx* ctx, void* dptr, int64_t rows, int64_t cols, int64_t ld, uint64_t N, gffm_mat** out) {
  GFFM_ENTER_CTX(ctx);
  if (!ctx || !out || !dptr) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (ld < rows) GFFM_FAIL(GFFM_ERR_INVALID, "ld < rows");
  if (N == 0 || N > (1ull << 32)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "modulus out of range for uint32 storage");
  gffm_mat* m = new gffm_mat();
  m->ctx = ctx;
  m->data = (uint32_t*)dptr;
  m->rows = rows;
  m->cols = cols;
  m->ld = ld;
  m->pcols = cols;
  m->pad = 0;
  m->N = N;
  m->owned = false;
  *out = m;
  return GFFM_OK;
}

static void free_plane_caches(gffm_mat* m) {
  for (int r = 0; r < 2; ++r) {
    if (m->cache[r].ptr) gffm_dev_free(m->ctx, m->cache[r].ptr);
    m->cache[r] = gffm_plane_cache();
  }
}

extern "C" int32_t gffm_mat_touch(gffm_mat* m) {
  GFFM_ENTER_MAT(m);
  if (!m) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  gffm_touch(m);
  return GFFM_OK;
}

extern "C" int32_t gffm_mat_drop_cache(gffm_mat* m) {
  GFFM_ENTER_MAT(m);
  if (!m) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  cudaSetDevice(m->ctx->device);
  free_plane_caches(m);
  return GFFM_OK;
}

extern "C" int32_t gffm_mat_destroy(gffm_mat* m) {
  GFFM_ENTER_MAT(m);
  if (!m) return GFFM_OK;
  // stream-ordered release: the blocks return to the pool after the work already queued on the context's stream (every
  // helper stream of the library is joined to it before an API call returns); safe from a finalizer thread
  cudaSetDevice(m->ctx->device);
  if (m->cache[0].ptr || m->cache[1].ptr) free_plane_caches(m);
  if (m->owned && m->data) gffm_dev_free(m->ctx, m->data);
  if (m->owned && m->data64) gffm_dev_free(m->ctx, m->data64);
  delete m;
  return GFFM_OK;
}

#define MAT_GETTER(name, type, expr)                          \
  extern "C" int32_t gffm_mat_##name(gffm_mat* m, type* v) {  \
    if (!m || !v) GFFM_FAIL(GFFM_ERR_INVALID, "null");        \
    *v = (expr);                                              \
    return GFFM_OK;                                           \
  }
MAT_GETTER(rows, int64_t, m->rows)
MAT_GETTER(cols, int64_t, m->cols)
MAT_GETTER(pad, int32_t, m->pad)
MAT_GETTER(modulus, uint64_t, m->N)
MAT_GETTER(ld, int64_t, m->ld)
MAT_GETTER(device_ptr, void*, m->wide ? m->data64 : (void*)m->data)

// ---- conversion kernels ----------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ bool to_residue(T x, uint64_t N, int do_mod, uint32_t& out) {
  // floored mod (reference mod_ops.jl:8); integers only (convert.(T,A) is exact or throws, CuModMatrix.jl:70-86)
  long long v;
  if constexpr (sizeof(T) == 8 && !std::is_integral<T>::value) {
    if (!(x == floor(x)) || fabs(x) > 9.0e18) return false;
    v = (long long)x;
  } else if constexpr (!std::is_integral<T>::value) {
    if (!(x == floorf(x)) || fabsf(x) > 9.0e18f) return false;
    v = (long long)x;
  } else {
    v = (long long)x;
  }
  if (do_mod) {
    long long r = v % (long long)N;
    if (r < 0) r += (long long)N;
    out = (uint32_t)r;
    return true;
  }
  if (v < 0 || v > 0xFFFFFFFFll) return false;
  out = (uint32_t)v;
  return true;
}

template <typename T>
__global__ void upload_kernel(const T* __restrict__ src, int64_t lds, uint32_t* __restrict__ dst, int64_t ldd, int64_t rows,
                              int64_t cols, uint64_t N, int do_mod, int* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i >= rows || j >= cols) return;
  uint32_t r = 0;
  if (!to_residue<T>(src[j * lds + i], N, do_mod, r)) atomicExch(bad, 1);
  dst[j * ldd + i] = r;
}

template <typename T>
__global__ void download_kernel(const uint32_t* __restrict__ src, int64_t lds, T* __restrict__ dst, int64_t ldd, int64_t rows,
                                int64_t cols) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i >= rows || j >= cols) return;
  dst[j * ldd + i] = (T)src[j * lds + i];
}

static size_t dtype_size(int dt) {
  switch (dt) {
    case GFFM_F32: return 4;
    case GFFM_F64: return 8;
    case GFFM_I64: return 8;
    case GFFM_U32: return 4;
    case GFFM_I32: return 4;
  }
  return 0;
}

// grid.y is limited to 65535 -> loop over column slabs
#define FOR_COL_SLABS(cols, ...)                                     \
  for (int64_t _c0 = 0; _c0 < (cols); _c0 += 65535) {                \
    const int64_t _nc = ((cols) - _c0) < 65535 ? ((cols) - _c0) : 65535; \
    __VA_ARGS__                                                      \
  }

extern "C" int32_t gffm_mat_upload(gffm_mat* m, const void* host, int32_t dtype, int64_t ld, int32_t do_mod) {
  GFFM_ENTER_MAT(m);
  if (!m || (!host && m->rows * m->cols > 0)) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  const size_t es = dtype_size(dtype);
  if (!es) GFFM_FAIL(GFFM_ERR_INVALID, "bad dtype %d", dtype);
  if (ld < m->rows) GFFM_FAIL(GFFM_ERR_INVALID, "ld < rows");
  if (m->rows == 0 || m->cols == 0) return GFFM_OK;
  gffm_touch(m);
  gffm_ctx* ctx = m->ctx;
  cudaSetDevice(ctx->device);
  if (m->wide) return gffm_wide_upload(m, host, dtype, ld, do_mod);
  if (dtype == GFFM_U32) {
    // residues already in storage format: one strided DMA straight into the matrix, reduction in place
    GFFM_CUDA(cudaMemcpy2DAsync(m->data, (size_t)m->ld * 4, host, (size_t)ld * 4, (size_t)m->rows * 4, (size_t)m->cols,
                                cudaMemcpyHostToDevice, ctx->stream));
    if (do_mod && m->N < (1ull << 32)) GFFM_TRY(gffm_ew_views(ctx, GFFM_EW_MOD, view_of(m), view_of(m), nullptr, 0, m->N));
    GFFM_CUDA(cudaStreamSynchronize(ctx->stream));  // the host buffer is caller-owned
    return GFFM_OK;
  }
  // staged through a device scratch buffer in column slabs (bounded memory, chunked for 8 GB operands)
  const int64_t slab_cols = std::max<int64_t>(1, std::min<int64_t>(m->cols, (int64_t)((256ull << 20) / (es * (size_t)m->rows))));
  GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_misc, (size_t)slab_cols * m->rows * es + 256));
  int* bad = (int*)((char*)ctx->ws_misc.ptr + (size_t)slab_cols * m->rows * es);
  bad = (int*)(((uintptr_t)bad + 15) & ~(uintptr_t)15);
  GFFM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
  for (int64_t c0 = 0; c0 < m->cols; c0 += slab_cols) {
    const int64_t nc = std::min<int64_t>(slab_cols, m->cols - c0);
    GFFM_CUDA(cudaMemcpy2DAsync(ctx->ws_misc.ptr, (size_t)m->rows * es, (const char*)host + (size_t)c0 * ld * es, (size_t)ld * es,
                                (size_t)m->rows * es, (size_t)nc, cudaMemcpyHostToDevice, ctx->stream));
    dim3 block(256);
    FOR_COL_SLABS(nc, {
      dim3 grid((unsigned)ceil_div(m->rows, 256), (unsigned)_nc);
      uint32_t* dst = m->data + (c0 + _c0) * m->ld;
      const char* src = (const char*)ctx->ws_misc.ptr + (size_t)_c0 * m->rows * es;
      switch (dtype) {
        case GFFM_F32: upload_kernel<float><<<grid, block, 0, ctx->stream>>>((const float*)src, m->rows, dst, m->ld, m->rows, _nc, m->N, do_mod, bad); break;
        case GFFM_F64: upload_kernel<double><<<grid, block, 0, ctx->stream>>>((const double*)src, m->rows, dst, m->ld, m->rows, _nc, m->N, do_mod, bad); break;
        case GFFM_I64: upload_kernel<long long><<<grid, block, 0, ctx->stream>>>((const long long*)src, m->rows, dst, m->ld, m->rows, _nc, m->N, do_mod, bad); break;
        case GFFM_U32: upload_kernel<unsigned int><<<grid, block, 0, ctx->stream>>>((const unsigned int*)src, m->rows, dst, m->ld, m->rows, _nc, m->N, do_mod, bad); break;
        case GFFM_I32: upload_kernel<int><<<grid, block, 0, ctx->stream>>>((const int*)src, m->rows, dst, m->ld, m->rows, _nc, m->N, do_mod, bad); break;
      }
      GFFM_LAUNCH_CHECK(ctx);
    })
    if (c0 + nc < m->cols) GFFM_CUDA(cudaStreamSynchronize(ctx->stream));  // scratch reuse
  }
  int hbad = 0;
  GFFM_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (hbad) GFFM_FAIL(GFFM_ERR_INEXACT, "InexactError: entry is not an integer representable in the target range");
  return GFFM_OK;
}

extern "C" int32_t gffm_mat_download(gffm_mat* m, void* host, int32_t dtype, int64_t ld, int32_t with_padding) {
  GFFM_ENTER_MAT(m);
  if (!m) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  const size_t es = dtype_size(dtype);
  if (!es) GFFM_FAIL(GFFM_ERR_INVALID, "bad dtype %d", dtype);
  const int64_t rows = with_padding ? m->rows + m->pad : m->rows;
  const int64_t cols = with_padding ? m->cols + m->pad : m->cols;
  if (ld < rows) GFFM_FAIL(GFFM_ERR_INVALID, "ld < rows");
  if (rows == 0 || cols == 0) return GFFM_OK;
  if (!host) GFFM_FAIL(GFFM_ERR_INVALID, "null host buffer");
  gffm_ctx* ctx = m->ctx;
  cudaSetDevice(ctx->device);
  if (m->wide) return gffm_wide_download(m, host, dtype, ld, with_padding);
  if (dtype == GFFM_U32 && (!with_padding || (rows <= m->ld && cols <= m->pcols))) {
    GFFM_CUDA(cudaMemcpy2DAsync(host, (size_t)ld * 4, m->data, (size_t)m->ld * 4, (size_t)rows * 4, (size_t)cols,
                                cudaMemcpyDeviceToHost, ctx->stream));
    GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
    return GFFM_OK;
  }
  const int64_t slab_cols = std::max<int64_t>(1, std::min<int64_t>(cols, (int64_t)((256ull << 20) / (es * (size_t)rows))));
  GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_misc, (size_t)slab_cols * rows * es));
  for (int64_t c0 = 0; c0 < cols; c0 += slab_cols) {
    const int64_t nc = std::min<int64_t>(slab_cols, cols - c0);
    dim3 block(256);
    FOR_COL_SLABS(nc, {
      dim3 grid((unsigned)ceil_div(rows, 256), (unsigned)_nc);
      const uint32_t* src = m->data + (c0 + _c0) * m->ld;
      char* dst = (char*)ctx->ws_misc.ptr + (size_t)_c0 * rows * es;
      switch (dtype) {
        case GFFM_F32: download_kernel<float><<<grid, block, 0, ctx->stream>>>(src, m->ld, (float*)dst, rows, rows, _nc); break;
        case GFFM_F64: download_kernel<double><<<grid, block, 0, ctx->stream>>>(src, m->ld, (double*)dst, rows, rows, _nc); break;
        case GFFM_I64: download_kernel<long long><<<grid, block, 0, ctx->stream>>>(src, m->ld, (long long*)dst, rows, rows, _nc); break;
        case GFFM_U32: download_kernel<unsigned int><<<grid, block, 0, ctx->stream>>>(src, m->ld, (unsigned int*)dst, rows, rows, _nc); break;
        case GFFM_I32: download_kernel<int><<<grid, block, 0, ctx->stream>>>(src, m->ld, (int*)dst, rows, rows, _nc); break;
      }
      GFFM_LAUNCH_CHECK(ctx);
    })
    GFFM_CUDA(cudaMemcpy2DAsync((char*)host + (size_t)c0 * ld * es, (size_t)ld * es, ctx->ws_misc.ptr, (size_t)rows * es,
                                (size_t)rows * es, (size_t)nc, cudaMemcpyDeviceToHost, ctx->stream));
    GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return GFFM_OK;
}

// ---- elementwise --------------------------------------------------------------------------------
// One grid-stride kernel for all ops (reference: 8 separate 256-thread kernels, kernel_ops/*.jl).  Operands are
// residues < 2^32; they are first reduced mod P when the override modulus is smaller than the stored range
// (mod_N override semantics, test/CuModMatrix/inplace_operations_test.jl:125-191).
__global__ void __launch_bounds__(256)
ewise_kernel(int op, uint32_t* __restrict__ C, int64_t ldc, const uint32_t* __restrict__ A, int64_t lda,
             const uint32_t* __restrict__ B, int64_t ldb, int64_t rows, int64_t cols, uint32_t s, const __grid_constant__ ModP mp) {
  const int64_t total = rows * cols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx / rows, i = idx - j * rows;
    const uint32_t P = (uint32_t)mp.P;
    uint32_t a = A[j * lda + i];
    if (a >= mp.P) a = (uint32_t)mod_u64(a, mp);
    uint32_t b = 0;
    if (B) {
      b = B[j * ldb + i];
      if (b >= mp.P) b = (uint32_t)mod_u64(b, mp);
    }
    uint32_t r;
    switch (op) {
      case GFFM_EW_MOD: r = a; break;
      case GFFM_EW_ADD: r = addmod_u32(a, b, P); break;
      case GFFM_EW_SUB: r = submod_u32(a, b, P); break;
      case GFFM_EW_MUL: r = mulmod_u32(a, b, mp); break;
      case GFFM_EW_SADD: r = addmod_u32(a, s, P); break;
      case GFFM_EW_SSUB: r = submod_u32(a, s, P); break;
      case GFFM_EW_RSSUB: r = submod_u32(s, a, P); break;
      default: r = mulmod_u32(a, s, mp); break;  // SMUL, SDIV (s already inverted)
    }
    C[j * ldc + i] = r;
  }
}

static uint32_t scalar_residue(int64_t s, uint64_t P) {
  long long r = (long long)(s % (long long)P);
  if (r < 0) r += (long long)P;
  return (uint32_t)r;
}

int32_t gffm_ew_views(gffm_ctx* ctx, int op, MatView C, MatView A, const MatView* B, int64_t scalar, uint64_t P) {
  if (C.rows != A.rows || C.cols != A.cols || (B && (B->rows != A.rows || B->cols != A.cols)))
    GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "elementwise operands differ in size");
  if (P == 0 || P > (1ull << 32)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "modulus out of range");
  if ((op == GFFM_EW_ADD || op == GFFM_EW_SUB || op == GFFM_EW_MUL) && !B) GFFM_FAIL(GFFM_ERR_INVALID, "binary op needs B");
  const int64_t total = C.rows * C.cols;
  if (total == 0) return GFFM_OK;
  uint32_t s = scalar_residue(scalar, P);
  if (op == GFFM_EW_SDIV) {
    s = (uint32_t)modinv_u64(s, P);
    if (s == 0 && P != 1) GFFM_FAIL(GFFM_ERR_INVALID, "scalar is not invertible mod %llu", (unsigned long long)P);
  }
  int64_t blocks = ceil_div(total, 256 * 4);
  const int64_t cap = (int64_t)ctx->num_sms * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  ModP mp = make_modp(P);
  if (P == (1ull << 32)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "P == 2^32");
  ewise_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(op, C.p, C.ld, A.p, A.ld, B ? B->p : nullptr, B ? B->ld : 0, C.rows, C.cols, s, mp);
  GFFM_LAUNCH_CHECK(ctx);
  return GFFM_OK;
}

extern "C" int32_t gffm_ewise(int32_t op, gffm_mat* C, gffm_mat* A, gffm_mat* B, int64_t scalar, uint64_t mod_override) {
  GFFM_ENTER_MAT(C);
  if (!C || !A) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (op < GFFM_EW_MOD || op > GFFM_EW_SDIV) GFFM_FAIL(GFFM_ERR_INVALID, "bad op");
  gffm_touch(C);
  const uint64_t P = mod_override ? mod_override : C->N;
  if (!mod_override) {  // reference checks moduli unless mod_N is given (CuModMatrix.jl add!/sub! preambles)
    if (A->N != C->N || (B && B->N != C->N)) GFFM_FAIL(GFFM_ERR_MODULUS_MISMATCH, "operands have different moduli");
  }
  if (gffm_any_wide({C, A, B})) return gffm_wide_ewise(op, C, A, B, scalar, P);
  MatView bv;
  if (B) bv = view_of(B);
  return gffm_ew_views(C->ctx, op, view_of(C), view_of(A), B ? &bv : nullptr, scalar, P);
}

// ---- small utility kernels ----------------------------------------------------------------------------
__global__ void copy_kernel(uint32_t* __restrict__ dst, int64_t ldd, const uint32_t* __restrict__ src, int64_t lds, int64_t rows, int64_t cols) {
  const int64_t total = rows * cols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx / rows, i = idx - j * rows;
    dst[j * ldd + i] = src[j * lds + i];
  }
}
__global__ void fill_kernel(uint32_t* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols, uint32_t v, int eye) {
  const int64_t total = rows * cols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx / rows, i = idx - j * rows;
    dst[j * ldd + i] = eye ? (i == j ? v : 0u) : v;
  }
}
__global__ void synth_kernel(uint32_t* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols, uint64_t seed, uint64_t N) {
  const int64_t total = rows * cols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx / rows, i = idx - j * rows;
    dst[j * ldd + i] = (uint32_t)(splitmix64(seed ^ (uint64_t)idx) % N);
  }
}
__global__ void transpose_kernel(uint32_t* __restrict__ dst, int64_t ldd, const uint32_t* __restrict__ src, int64_t lds, int64_t rows, int64_t cols) {
  __shared__ uint32_t t[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
  for (int jj = threadIdx.y; jj < 32; jj += 8) {
    const int64_t i = i0 + threadIdx.x, j = j0 + jj;
    t[jj][threadIdx.x] = (i < rows && j < cols) ? src[j * lds + i] : 0u;
  }
  __syncthreads();
  for (int ii = threadIdx.y; ii < 32; ii += 8) {
    const int64_t j = j0 + threadIdx.x, i = i0 + ii;  // dst is cols x rows: element (j,i)
    if (i < rows && j < cols) dst[i * ldd + j] = t[threadIdx.x][ii];
  }
}

static unsigned grid_for(gffm_ctx* ctx, int64_t total) {
  int64_t blocks = ceil_div(total, 256 * 4);
  const int64_t cap = (int64_t)ctx->num_sms * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

int32_t gffm_copy_views(gffm_ctx* ctx, MatView dst, MatView src) {
  if (dst.rows != src.rows || dst.cols != src.cols) GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "copy size mismatch");
  if (dst.rows * dst.cols == 0) return GFFM_OK;
  GFFM_CUDA(cudaMemcpy2DAsync(dst.p, (size_t)dst.ld * 4, src.p, (size_t)src.ld * 4, (size_t)src.rows * 4, (size_t)src.cols,
                              cudaMemcpyDeviceToDevice, ctx->stream));
  return GFFM_OK;
}
int32_t gffm_fill_view(gffm_ctx* ctx, MatView dst, uint32_t value) {
  if (dst.rows * dst.cols == 0) return GFFM_OK;
  fill_kernel<<<grid_for(ctx, dst.rows * dst.cols), 256, 0, ctx->stream>>>(dst.p, dst.ld, dst.rows, dst.cols, value, 0);
  GFFM_LAUNCH_CHECK(ctx);
  return GFFM_OK;
}

extern "C" int32_t gffm_mat_copy(gffm_mat* dst, gffm_mat* src) {
  GFFM_ENTER_MAT(dst);
  if (!dst || !src) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (dst->rows != src->rows || dst->cols != src->cols) GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "copy!: sizes differ");
  gffm_touch(dst);
  if (gffm_any_wide({dst, src})) return gffm_wide_copy(dst, src);
  return gffm_copy_views(dst->ctx, view_of(dst), view_of(src));
}
extern "C" int32_t gffm_mat_copy_block(gffm_mat* dst, int64_t dr0, int64_t dc0, gffm_mat* src, int64_t sr0, int64_t sc0, int64_t nr, int64_t nc) {
  GFFM_ENTER_MAT(dst);
  if (!dst || !src) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (dr0 < 0 || dc0 < 0 || sr0 < 0 || sc0 < 0 || nr < 0 || nc < 0 || dr0 + nr > dst->rows || dc0 + nc > dst->cols ||
      sr0 + nr > src->rows || sc0 + nc > src->cols)
    GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "block out of range");
  GFFM_NARROW_ONLY(dst, src);
  gffm_touch(dst);
  return gffm_copy_views(dst->ctx, sub_view(view_of(dst), dr0, dc0, nr, nc), sub_view(view_of(src), sr0, sc0, nr, nc));
}
extern "C" int32_t gffm_mat_fill(gffm_mat* m, int64_t value) {
  GFFM_ENTER_MAT(m);
  if (!m) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  gffm_touch(m);
  if (m->wide) return gffm_wide_fill(m, value, 0, 0, 0, m->rows, m->cols);
  return gffm_fill_view(m->ctx, view_of(m), scalar_residue(value, m->N));
}
extern "C" int32_t gffm_mat_zero(gffm_mat* m) { return gffm_mat_fill(m, 0); }
extern "C" int32_t gffm_mat_eye(gffm_mat* m) {
  GFFM_ENTER_MAT(m);
  if (!m) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (m->rows * m->cols == 0) return GFFM_OK;
  gffm_touch(m);
  if (m->wide) return gffm_wide_fill(m, 1, 1, 0, 0, m->rows, m->cols);
  fill_kernel<<<grid_for(m->ctx, m->rows * m->cols), 256, 0, m->ctx->stream>>>(m->data, m->ld, m->rows, m->cols, scalar_residue(1, m->N), 1);
  GFFM_LAUNCH_CHECK(m->ctx);
  return GFFM_OK;
}
extern "C" int32_t gffm_mat_synth(gffm_mat* m, uint64_t seed) {
  GFFM_ENTER_MAT(m);
  if (!m) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (m->rows * m->cols == 0) return GFFM_OK;
  gffm_touch(m);
  if (m->wide) return gffm_wide_synth(m, seed);
  synth_kernel<<<grid_for(m->ctx, m->rows * m->cols), 256, 0, m->ctx->stream>>>(m->data, m->ld, m->rows, m->cols, seed, m->N);
  GFFM_LAUNCH_CHECK(m->ctx);
  return GFFM_OK;
}
extern "C" int32_t gffm_mat_rand(gffm_mat* m, uint64_t seed) { return gffm_mat_synth(m, splitmix64(seed ^ 0x5DEECE66Dull)); }

extern "C" int32_t gffm_mat_set_modulus(gffm_mat* m, uint64_t N, int32_t reduce) {
  GFFM_ENTER_MAT(m);
  if (!m) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (N == 0) GFFM_FAIL(GFFM_ERR_INVALID, "modulus must be positive");
  if (N > (1ull << 52)) GFFM_FAIL(GFFM_ERR_MODULUS_TOO_LARGE, "Modulus is bigger than 2^52");
  if (m->wide) {  // stays in uint64 storage whatever the new modulus is
    m->N = N;
    gffm_touch(m);
    return reduce ? gffm_wide_ewise(GFFM_EW_MOD, m, m, nullptr, 0, N) : GFFM_OK;
  }
  if (N > (1ull << 32)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "a matrix created with N <= 2^32 keeps uint32 storage: create the matrix with the larger modulus instead");
  m->N = N;
  gffm_touch(m);
  if (reduce && N < (1ull << 32)) return gffm_ew_views(m->ctx, GFFM_EW_MOD, view_of(m), view_of(m), nullptr, 0, N);
  return GFFM_OK;
}

extern "C" int32_t gffm_mat_get_elem(gffm_mat* m, int64_t i, int64_t j, int64_t* value) {
  GFFM_ENTER_MAT(m);
  if (!m || !value) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (i < 0 || j < 0 || i >= m->rows || j >= m->cols) GFFM_FAIL(GFFM_ERR_INVALID, "BoundsError");
  if (m->wide) return gffm_wide_get(m, i, j, value);
  uint32_t v = 0;
  GFFM_CUDA(cudaMemcpyAsync(&v, m->data + j * m->ld + i, 4, cudaMemcpyDeviceToHost, m->ctx->stream));
  GFFM_CUDA(cudaStreamSynchronize(m->ctx->stream));
  *value = (int64_t)v;
  return GFFM_OK;
}
extern "C" int32_t gffm_mat_set_elem(gffm_mat* m, int64_t i, int64_t j, int64_t value) {
  GFFM_ENTER_MAT(m);
  if (!m) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (i < 0 || j < 0 || i >= m->rows || j >= m->cols) GFFM_FAIL(GFFM_ERR_INVALID, "BoundsError");
  gffm_touch(m);
  if (m->wide) return gffm_wide_fill(m, value, 0, i, j, 1, 1);
  return gffm_fill_view(m->ctx, sub_view(view_of(m), i, j, 1, 1), scalar_residue(value, m->N));
}
extern "C" int32_t gffm_mat_transpose(gffm_mat* dst, gffm_mat* src) {
  GFFM_ENTER_MAT(dst);
  if (!dst || !src) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (dst->rows != src->cols || dst->cols != src->rows) GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "transpose: sizes differ");
  GFFM_NARROW_ONLY(dst, src);
  if (src->rows * src->cols == 0) return GFFM_OK;
  gffm_touch(dst);
  dim3 grid((unsigned)ceil_div(src->rows, 32), (unsigned)ceil_div(src->cols, 32));
  if (grid.y > 65535) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "transpose: too many columns");
  transpose_kernel<<<grid, dim3(32, 8), 0, src->ctx->stream>>>(dst->data, dst->ld, src->data, src->ld, src->rows, src->cols);
  GFFM_LAUNCH_CHECK(src->ctx);
  return GFFM_OK;
}

// ---- equality / checksum ----------------------------------------------------------------------------
__global__ void checksum_kernel(const uint32_t* __restrict__ a, int64_t lda, const uint32_t* __restrict__ b, int64_t ldb, int64_t rows,
                                int64_t cols, unsigned long long* __restrict__ out /* [0]=sum, [1]=diff count */) {
  unsigned long long sum = 0, diff = 0;
  const int64_t total = rows * cols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = idx / rows, i = idx - j * rows;
    const uint32_t va = a[j * lda + i];
    sum += (unsigned long long)va * splitmix64((uint64_t)idx);
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
static int32_t checksum_impl(gffm_mat* a, gffm_mat* b, unsigned long long out[2]) {
  gffm_ctx* ctx = a->ctx;
  if (a->wide || (b && b->wide)) {
    if (!a->wide || (b && !b->wide)) {  // different storage kinds never compare equal
      out[0] = 0;
      out[1] = 1;
      return GFFM_OK;
    }
    return gffm_wide_checksum(a, b, out);
  }
  GFFM_TRY(gffm_ws_reserve(ctx, &ctx->ws_misc2, 64));
  unsigned long long* d = (unsigned long long*)ctx->ws_misc2.ptr;
  GFFM_CUDA(cudaMemsetAsync(d, 0, 16, ctx->stream));
  if (a->rows * a->cols > 0) {
    checksum_kernel<<<grid_for(ctx, a->rows * a->cols), 256, 0, ctx->stream>>>(a->data, a->ld, b ? b->data : nullptr, b ? b->ld : 0, a->rows, a->cols, d);
    GFFM_LAUNCH_CHECK(ctx);
  }
  GFFM_CUDA(cudaMemcpyAsync(out, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
  GFFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return GFFM_OK;
}
extern "C" int32_t gffm_mat_equal(gffm_mat* a, gffm_mat* b, int32_t* equal) {
  GFFM_ENTER_MAT(a);
  if (!a || !b || !equal) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (a->rows != b->rows || a->cols != b->cols) {
    *equal = 0;
    return GFFM_OK;
  }
  unsigned long long out[2];
  GFFM_TRY(checksum_impl(a, b, out));
  *equal = out[1] == 0;
  return GFFM_OK;
}
extern "C" int32_t gffm_mat_checksum(gffm_mat* a, uint64_t* sum) {
  GFFM_ENTER_MAT(a);
  if (!a || !sum) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  unsigned long long out[2];
  GFFM_TRY(checksum_impl(a, nullptr, out));
  *sum = out[0];
  return GFFM_OK;
}

// ---------------------------------------------------------------------------------------------
// scalar SIMT GEMM: exact uint64 accumulation, 32x32 output tile, any R,P < 2^32.  Used for tiny shapes
// (where a tensor-core tile would be mostly padding) and as the on-device cross-check of the tcgen05 paths.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gemm_simt_kernel(uint32_t* __restrict__ C, int64_t ldc, const uint32_t* __restrict__ A, int64_t lda, const uint32_t* __restrict__ B,
                 int64_t ldb, int m, int n, int k, int mode, int reduce_every_step, const __grid_constant__ ModP mp) {
  __shared__ uint32_t sA[32][33];  // [kk][i]
  __shared__ uint32_t sB[32][33];  // [j][kk]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty 0..7
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  uint64_t acc[4] = {0, 0, 0, 0};  // rows i0+tx, columns j0 + ty + 8*c
  for (int k0 = 0; k0 < k; k0 += 32) {
    for (int kk = ty; kk < 32; kk += 8) {
      const int i = i0 + tx, kg = k0 + kk;
      sA[kk][tx] = (i < m && kg < k) ? A[(int64_t)kg * lda + i] : 0u;
    }
    for (int jj = ty; jj < 32; jj += 8) {
      const int j = j0 + jj, kg = k0 + tx;
      sB[jj][tx] = (j < n && kg < k) ? B[(int64_t)j * ldb + kg] : 0u;
    }
    __syncthreads();
    if (reduce_every_step) {
      for (int kk = 0; kk < 32; ++kk) {
        const uint64_t a = sA[kk][tx];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] = mod_u64(acc[c] + mod_u64(a * sB[ty + 8 * c][kk], mp), mp);
      }
    } else {
#pragma unroll 8
      for (int kk = 0; kk < 32; ++kk) {
        const uint64_t a = sA[kk][tx];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] += a * sB[ty + 8 * c][kk];
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c] = mod_u64(acc[c], mp);
    }
    __syncthreads();
  }
  const int i = i0 + tx;
  if (i >= m) return;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int j = j0 + ty + 8 * c;
    if (j >= n) continue;
    uint32_t r = (uint32_t)acc[c];
    uint32_t* dst = C + (int64_t)j * ldc + i;
    if (mode == GFFM_GEMM_ADD) r = addmod_u32(*dst, r, (uint32_t)mp.P);
    else if (mode == GFFM_GEMM_SUB) r = submod_u32(*dst, r, (uint32_t)mp.P);
    *dst = r;
  }
}

int32_t gffm_gemm_simt(gffm_ctx* ctx, MatView C, MatView A, MatView B, uint64_t R, uint64_t P, int mode) {
  if (P == 0 || P >= (1ull << 32)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "SIMT GEMM needs P < 2^32");
  const int64_t m = A.rows, k = A.cols, n = B.cols;
  if (m == 0 || n == 0) return GFFM_OK;
  dim3 grid((unsigned)ceil_div(m, 32), (unsigned)ceil_div(n, 32));
  if (grid.y > 65535) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "SIMT GEMM: too many columns");
  // inputs are arbitrary uint32 (< 2^32): 32 products of < 2^64 cannot be summed safely -> reduce each step unless
  // the caller's modulus guarantees a,b < 2^29 (32 * 2^58 + 2^32 < 2^64)
  const int every = (P > (1ull << 29) || R > (1ull << 29)) ? 1 : 0;
  gemm_simt_kernel<<<grid, 256, 0, ctx->stream>>>(C.p, C.ld, A.p, A.ld, B.p, B.ld, (int)m, (int)n, (int)k, mode, every, make_modp(P));
  GFFM_LAUNCH_CHECK(ctx);
  return GFFM_OK;
}

// GEMV: gemv.cu

// ---------------------------------------------------------------------------------------------
// GEMM dispatcher
// ---------------------------------------------------------------------------------------------
int32_t gffm_gemm_views(gffm_ctx* ctx, MatView C, MatView A, MatView B, uint64_t R, uint64_t P, int mode, int algo) {
  if (A.cols != B.rows || C.rows != A.rows || C.cols != B.cols)
    GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "gemm: C %lldx%lld = A %lldx%lld * B %lldx%lld", (long long)C.rows, (long long)C.cols,
              (long long)A.rows, (long long)A.cols, (long long)B.rows, (long long)B.cols);
  if (mode < GFFM_GEMM_STORE || mode > GFFM_GEMM_SUB) GFFM_FAIL(GFFM_ERR_INVALID, "bad gemm mode");
  if (P == 0 || P >= (1ull << 32)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "gemm needs 0 < P < 2^32");
  const int64_t m = A.rows, k = A.cols, n = B.cols;
  if (algo == GFFM_ALGO_AUTO) {
    const double work = (double)m * (double)n * (double)(k > 0 ? k : 1);
    if (work < 2.0 * 128.0 * 128.0 * 128.0 || k < 16 || !gffm_tc_available(ctx)) algo = GFFM_ALGO_SIMT;
    else if (R <= 65536) algo = GFFM_ALGO_LIMB;
    else if (R < (1ull << 32)) algo = GFFM_ALGO_RNS;
    else algo = GFFM_ALGO_SIMT;
  }
  // large products whose K fits one accumulation chunk can take the tiled multi-stream path (split / CRT concurrent with the
  // GEMM).  Opt-in (GFFM_TILED=1): measured on B200 it gains nothing for resident operands -- the issue-bound helper kernels
  // take issue slots from the single MMA-issuing warp and the GEMM slows by what the helpers save (profiles/r01_notes.md).
  if ((algo == GFFM_ALGO_LIMB || algo == GFFM_ALGO_RNS) && getenv("GFFM_TILED") != nullptr) {
    const bool rns = algo == GFFM_ALGO_RNS;
    const bool ok_enc = rns ? (R > 65536 && R < (1ull << 32)) : (R <= 65536);
    const int64_t kchunk = gffm_gemm_kchunk(R, rns);
    if (ok_enc && k <= kchunk && m >= 4096 && n >= 4096 && k >= 1024)
      return gffm_gemm_tiled(ctx, C, A, B, R, P, mode, rns && (R % P) == 0);
  }
  switch (algo) {
    case GFFM_ALGO_SIMT: return gffm_gemm_simt(ctx, C, A, B, R, P, mode);
    case GFFM_ALGO_LIMB: return gffm_gemm_tc_limb(ctx, C, A, B, R, P, mode);
    case GFFM_ALGO_RNS: return gffm_gemm_tc_rns(ctx, C, A, B, R, P, mode, /*balanced=*/(R % P) == 0, nullptr, 0);
  }
  GFFM_FAIL(GFFM_ERR_INVALID, "bad algo %d", algo);
}

extern "C" int32_t gffm_gemm_block(gffm_mat* C, int64_t cr0, int64_t cc0, gffm_mat* A, int64_t ar0, int64_t ac0, gffm_mat* B,
                                   int64_t br0, int64_t bc0, int64_t m, int64_t n, int64_t k, uint64_t R, uint64_t P, int32_t mode,
                                   int32_t algo) {
  GFFM_ENTER_MAT(C);
  if (!C || !A || !B) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  GFFM_NARROW_ONLY(C, A, B);
  if (cr0 < 0 || cc0 < 0 || ar0 < 0 || ac0 < 0 || br0 < 0 || bc0 < 0 || m < 0 || n < 0 || k < 0 || cr0 + m > C->rows ||
      cc0 + n > C->cols || ar0 + m > A->rows || ac0 + k > A->cols || br0 + k > B->rows || bc0 + n > B->cols)
    GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "gemm_block: block out of range");
  if (!P) {
    if (A->N != B->N || A->N != C->N) GFFM_FAIL(GFFM_ERR_MODULUS_MISMATCH, "gemm operands have different moduli");
    P = C->N;
  }
  if (!R) R = A->N > B->N ? A->N : B->N;
  gffm_touch(C);
  // operands may reuse their cached 8-bit planes (C is excluded: it is being written, and may alias neither input)
  MatView av = (A != C) ? cached_view_of(A) : view_of(A), bv = (B != C) ? cached_view_of(B) : view_of(B);
  return gffm_gemm_views(C->ctx, sub_view(view_of(C), cr0, cc0, m, n), sub_view(av, ar0, ac0, m, k), sub_view(bv, br0, bc0, k, n), R, P,
                         mode, algo);
}

extern "C" int32_t gffm_gemm(gffm_mat* C, gffm_mat* A, gffm_mat* B, uint64_t R, uint64_t P, int32_t mode, int32_t algo) {
  GFFM_ENTER_MAT(C);
  if (!C || !A || !B) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  GFFM_NARROW_ONLY(C, A, B);
  // reference check order: modulus first, then sizes (CuModMatrix.jl:769-783)
  if (!P && (A->N != B->N || A->N != C->N)) GFFM_FAIL(GFFM_ERR_MODULUS_MISMATCH, "gemm operands have different moduli");
  if (A->cols != B->rows || C->rows != A->rows || C->cols != B->cols)
    GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "mul!: C %lldx%lld = A %lldx%lld * B %lldx%lld", (long long)C->rows, (long long)C->cols,
              (long long)A->rows, (long long)A->cols, (long long)B->rows, (long long)B->cols);
  return gffm_gemm_block(C, 0, 0, A, 0, 0, B, 0, 0, A->rows, B->cols, A->cols, R, P, mode, algo);
}
