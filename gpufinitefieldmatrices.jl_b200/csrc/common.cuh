// common.cuh -- internal types, error handling and exact modular-arithmetic device helpers of libgffm.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <initializer_list>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/gffm.h"

#define GFFM_REF_PAD 32  // TILE_WIDTH, reference src/CuModMatrix/CuModMatrix.jl:2,62

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
void gffm_set_error(const char* fmt, ...);
#define GFFM_FAIL(code, ...)            \
  do {                                  \
    gffm_set_error(__VA_ARGS__);        \
    return (code);                      \
  } while (0)
#define GFFM_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      gffm_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                     cudaGetErrorString(_e));                                             \
      return GFFM_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)
#define GFFM_TRY(expr)            \
  do {                            \
    int32_t _s = (expr);          \
    if (_s != GFFM_OK) return _s; \
  } while (0)

// ---------------------------------------------------------------------------------------------
// context / matrix
// ---------------------------------------------------------------------------------------------
struct gffm_workspace {
  void* ptr = nullptr;
  size_t bytes = 0;
};

// Every ABI entry makes its context's device current first: a process may drive several GPUs (one context each) and the
// calling thread's current device is whatever the host runtime left there.
#define GFFM_ENTER_CTX(c) do { if (c) cudaSetDevice((c)->device); } while (0)
#define GFFM_ENTER_MAT(m) do { if ((m) && (m)->ctx) cudaSetDevice((m)->ctx->device); } while (0)

// Kernel attributes (max dynamic shared memory, non-portable cluster size) are per DEVICE in the CUDA runtime: a process that
// drives several GPUs must set them once on each.  run(device, f) calls f the first time it sees a device ordinal.
struct PerDeviceOnce {
  std::mutex mu;
  unsigned long long done = 0;
  template <class F>
  void run(int device, F&& f) {
    const unsigned long long bit = 1ull << (device & 63);
    std::lock_guard<std::mutex> g(mu);
    if (!(done & bit)) {
      f();
      done |= bit;
    }
  }
};

struct gffm_ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int64_t launches = 0;
  int64_t alloc_bytes = 0, alloc_calls = 0;  // device memory requested by the library on this context since creation (gffm_alloc_stats)
  int gemm_ctas = 0;  // cap on the persistent GEMM grid (0 = one CTA per SM); leaves SMs to concurrent NCCL kernels (gffm_set_gemm_ctas)
  // grow-only scratch buffers (stream-ordered reuse)
  gffm_workspace ws_planes_a, ws_planes_b, ws_eplanes, ws_misc, ws_misc2, ws_pinned, ws_invtab, ws_scratch, ws_host, ws_gemv;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_aux = nullptr;  // copy / helper streams of the tiled GEMM
  std::vector<cudaEvent_t> ev_pool;
  std::vector<cudaEvent_t> tile_events;  // (start, end) pairs around the GEMM launches of the last profiled tiled product
  uint64_t inv_table_N = 0;
  std::vector<double> timings;
  std::vector<double> elim_timings;  // {inner panels, U12 = L11^-1 A12, trailing GEMM} ms of the last profiled elimination
  bool profile = false;
  // GFFM_TRACE=1: timing events around the kernels / copies of the multi-stream pipelines (external-plane GEMM, multi-GPU layer);
  // dumped as a timeline by gffm_trace_dump (called from gffm_mg_barrier / gffm_sync).  Diagnostics only.
  bool trace = false;
  struct TraceRec {
    const char* what;
    int idx, sid;
    cudaEvent_t a, b;
  };
  std::vector<TraceRec> trace_recs;
  int n_ev = 0;  // events recorded by the last profiled call
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// Cached 8-bit operand planes of a matrix (built lazily by the first GEMM that needs them, reused while the
// matrix is unchanged -- every writer bumps `version`).  One entry per operand role.
struct gffm_plane_cache {
  void* ptr = nullptr;
  size_t bytes = 0;
  bool valid = false;
  uint64_t version = 0;
  int mode = 0, nplanes = 0, balanced = 0;
  uint64_t R = 0;
  int64_t r0 = 0, c0 = 0, rows = 0, cols = 0, k0 = 0, kc = 0, Kp = 0, rowsP = 0;
};

struct gffm_mat {
  gffm_ctx* ctx = nullptr;
  uint32_t* data = nullptr;  // column-major, element (i,j) at data[j*ld + i]
  int64_t rows = 0, cols = 0;
  int64_t ld = 0;     // physical leading dimension (elements), >= rows + pad, multiple of 32
  int64_t pcols = 0;  // physical columns allocated (>= cols + pad)
  int32_t pad = GFFM_REF_PAD;
  uint64_t N = 0;
  bool owned = true;
  // moduli 2^32 < N <= 2^52 (CuModMatrix.jl:55-59): uint64 residues in data64, `data` stays null.  Container + elementwise API only
  // (wide.cu); every other entry point rejects such matrices (GFFM_NARROW_ONLY).
  bool wide = false;
  void* data64 = nullptr;
  uint64_t version = 1;         // bumped by every API call that writes the matrix
  gffm_plane_cache cache[2];    // [0] = as A operand (transposed planes), [1] = as B operand
};
static inline void gffm_touch(gffm_mat* m) { m->version++; }

static inline bool gffm_any_wide(std::initializer_list<const gffm_mat*> ms) {
  for (const gffm_mat* m : ms)
    if (m && m->wide) return true;
  return false;
}
#define GFFM_NARROW_ONLY(...)                                                                                                     \
  do {                                                                                                                            \
    if (gffm_any_wide({__VA_ARGS__}))                                                                                             \
      GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "matrices with a modulus above 2^32 (uint64 storage) support the container and the elementwise API only; " \
                                      "products above 2^32 go through KaratsubaMatrix");                                            \
  } while (0)
int32_t gffm_wide_upload(gffm_mat* m, const void* host, int32_t dtype, int64_t ld, int32_t do_mod);
int32_t gffm_wide_download(gffm_mat* m, void* host, int32_t dtype, int64_t ld, int32_t with_padding);
int32_t gffm_wide_ewise(int op, gffm_mat* C, gffm_mat* A, gffm_mat* B, int64_t scalar, uint64_t P);
int32_t gffm_wide_fill(gffm_mat* m, int64_t value, int eye, int64_t r0, int64_t c0, int64_t nr, int64_t nc);
int32_t gffm_wide_synth(gffm_mat* m, uint64_t seed);
int32_t gffm_wide_copy(gffm_mat* dst, gffm_mat* src);
int32_t gffm_wide_get(gffm_mat* m, int64_t i, int64_t j, int64_t* value);
int32_t gffm_wide_checksum(gffm_mat* a, gffm_mat* b, unsigned long long out[2]);

static inline cudaEvent_t gffm_trace_begin(gffm_ctx* ctx, cudaStream_t st) {
  if (!ctx->trace) return nullptr;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, st);
  return e;
}
static inline void gffm_trace_end(gffm_ctx* ctx, const char* what, int idx, int sid, cudaEvent_t a, cudaStream_t st) {
  if (!a) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, st);
  ctx->trace_recs.push_back(gffm_ctx::TraceRec{what, idx, sid, a, e});
}
void gffm_trace_dump(gffm_ctx* ctx, int rank);

int32_t gffm_ws_reserve(gffm_ctx* ctx, gffm_workspace* ws, size_t bytes);
// Stream-ordered device memory for matrices and plane caches (cudaMallocAsync on the device's default pool, which is told to
// keep freed blocks): temporaries such as the W / L factors of an elimination cost microseconds instead of a device-wide
// cudaMalloc / cudaFree.  gffm_dev_free orders the release after the work queued on the context's stream.
cudaError_t gffm_dev_alloc(gffm_ctx* ctx, void** ptr, size_t bytes);
void gffm_dev_free(gffm_ctx* ctx, void* ptr);
int32_t gffm_pinned_reserve(gffm_ctx* ctx, size_t bytes);

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// ---------------------------------------------------------------------------------------------
// exact modular arithmetic
// ---------------------------------------------------------------------------------------------
// Barrett constants for a modulus P < 2^63: mu = floor((2^64-1)/P).  q = mulhi(v,mu) underestimates
// floor(v/P) by at most 2 (mu may be one less than floor(2^64/P) when P | 2^64), so up to two conditional
// subtractions restore r in [0,P).
struct ModP {
  uint64_t P;
  uint64_t mu;
};
static inline ModP make_modp(uint64_t P) {
  ModP m;
  m.P = P;
  m.mu = P ? (~0ull) / P : 0;
  return m;
}
__host__ __device__ __forceinline__ uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
// v arbitrary u64 -> v mod P
__host__ __device__ __forceinline__ uint64_t mod_u64(uint64_t v, const ModP& m) {
  uint64_t q = mulhi64(v, m.mu);
  uint64_t r = v - q * m.P;
  if (r >= m.P) r -= m.P;
  if (r >= m.P) r -= m.P;
  return r;
}
// (a*b) mod P for a,b < P < 2^32
__host__ __device__ __forceinline__ uint32_t mulmod_u32(uint32_t a, uint32_t b, const ModP& m) {
  return (uint32_t)mod_u64((uint64_t)a * b, m);
}
__host__ __device__ __forceinline__ uint32_t addmod_u32(uint32_t a, uint32_t b, uint32_t P) {
  uint64_t s = (uint64_t)a + b;
  return (uint32_t)(s >= P ? s - P : s);
}
__host__ __device__ __forceinline__ uint32_t submod_u32(uint32_t a, uint32_t b, uint32_t P) {
  return a >= b ? a - b : (uint32_t)((uint64_t)a + P - b);
}
// extended Euclid (reference pluq_kernels.jl:11-31); returns 0 when gcd != 1
__host__ __device__ inline uint64_t modinv_u64(uint64_t p, uint64_t P) {
  int64_t inv = 0, new_inv = 1;
  int64_t rem = (int64_t)P, new_rem = (int64_t)(p % P);
  while (new_rem != 0) {
    int64_t q = rem / new_rem;
    int64_t t = inv - q * new_inv;
    inv = new_inv;
    new_inv = t;
    t = rem - q * new_rem;
    rem = new_rem;
    new_rem = t;
  }
  if (rem != 1) return 0;
  if (inv < 0) inv += (int64_t)P;
  return (uint64_t)inv;
}
// same algorithm on 32-bit remainders (hardware-friendly division); P < 2^32.  ~10x faster than the 64-bit form on
// the GPU, where 64-bit integer division is emulated.
__host__ __device__ inline uint32_t modinv_u32(uint32_t p, uint32_t P) {
  uint32_t r0 = P, r1 = p % P;
  long long t0 = 0, t1 = 1;
  while (r1 != 0) {
    const uint32_t q = r0 / r1;
    const uint32_t r2 = r0 - q * r1;
    const long long t2 = t0 - (long long)q * t1;
    r0 = r1;
    r1 = r2;
    t0 = t1;
    t1 = t2;
  }
  if (r0 != 1) return 0;
  if (t0 < 0) t0 += (long long)P;
  return (uint32_t)t0;
}
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

// ---------------------------------------------------------------------------------------------
// internal entry points shared between translation units
// ---------------------------------------------------------------------------------------------
struct MatView {  // a sub-block of a column-major uint32 matrix
  uint32_t* p;
  int64_t ld;
  int64_t rows, cols;
  gffm_mat* owner = nullptr;  // set only by the public GEMM entry points: enables the operand-plane cache
  int64_t r0 = 0, c0 = 0;     // offset of this view inside owner
};
static inline MatView view_of(gffm_mat* m) { return MatView{m->data, m->ld, m->rows, m->cols}; }
static inline MatView sub_view(const MatView& v, int64_t r0, int64_t c0, int64_t nr, int64_t nc) {
  return MatView{v.p + c0 * v.ld + r0, v.ld, nr, nc, v.owner, v.r0 + r0, v.c0 + c0};
}
// view that may use the plane cache of m (owned matrices only: external memory can change behind our back)
static inline MatView cached_view_of(gffm_mat* m) {
  MatView v = view_of(m);
  if (m->owned) v.owner = m;
  return v;
}

// C (op)= A*B mod P on views; inputs < R.  algo as in gffm.h.
int32_t gffm_gemm_views(gffm_ctx* ctx, MatView C, MatView A, MatView B, uint64_t R, uint64_t P, int mode, int algo);
// scalar SIMT kernel (exact, any R,P < 2^32)
int32_t gffm_gemm_simt(gffm_ctx* ctx, MatView C, MatView A, MatView B, uint64_t R, uint64_t P, int mode);
// tensor-core paths
int32_t gffm_gemm_tc_limb(gffm_ctx* ctx, MatView C, MatView A, MatView B, uint64_t R, uint64_t P, int mode);
int32_t gffm_gemm_tc_rns(gffm_ctx* ctx, MatView C, MatView A, MatView B, uint64_t R, uint64_t P, int mode,
                         bool balanced, uint32_t* kara_hi, uint64_t kara_N1);
bool gffm_tc_available(gffm_ctx* ctx);
int64_t gffm_gemm_kchunk(uint64_t R, bool rns);
int32_t gffm_gemm_tiled(gffm_ctx* ctx, MatView C, MatView A, MatView B, uint64_t R, uint64_t P, int mode, bool balanced);
// z = A*x mod P on raw device pointers (gemv.cu); Karatsuba mat x vec in one pass over A1, A2
int32_t gffm_gemv_raw(gffm_ctx* ctx, uint32_t* z, const uint32_t* A, int64_t lda, const uint32_t* x, int64_t m, int64_t k, uint64_t R, uint64_t P);
int32_t gffm_kmat_gemv(gffm_ctx* ctx, gffm_mat* C1, gffm_mat* C2, gffm_mat* A1, gffm_mat* A2, gffm_mat* B1, gffm_mat* B2, uint64_t N1, uint64_t N2);

int32_t gffm_ew_views(gffm_ctx* ctx, int op, MatView C, MatView A, const MatView* B, int64_t scalar, uint64_t P);
int32_t gffm_copy_views(gffm_ctx* ctx, MatView dst, MatView src);
int32_t gffm_fill_view(gffm_ctx* ctx, MatView dst, uint32_t value);

#define GFFM_LAUNCH_CHECK(ctx)                 \
  do {                                         \
    (ctx)->launches++;                         \
    GFFM_CUDA(cudaGetLastError());             \
  } while (0)
