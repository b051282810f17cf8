/* oracle_c.c -- C restatement of the reference's CPU ground truth for the hot path.  TEST INFRASTRUCTURE ONLY:
 * used by tests/ as a second, faster checker and by bench.py as the `cpu_baseline` / `--impl reference` arm.  Never
 * linked into or called by the product (gpufinitefieldmatrices.jl_b200/).
 *
 * The reference's tests compute their expected values with plain host integers, `mod.(A*B, N)` on Matrix{Int}
 * (/root/reference/test/CuModMatrix/stripe_mul_test.jl:11,26,48; matmul_operations_test.jl:65,107); Julia is not
 * available in this image, so that computation is restated here: exact uint64 accumulation with a reduction every
 * `chunk` terms so that no partial sum can overflow, pthreads over row blocks (all host cores).
 * The elimination follows /root/reference/src/CuModMatrix/rref_lu_pluq/pluq_kernels.jl:46-157 (pivot = maximum
 * residue, first index; pivot row scaled to 1; L keeps the pivot and the un-normalised sub-column) in its
 * well-defined echelon form (see oracle.py:echelon).
 * Parity status: pinned against oracle.py (which is pinned against the reference's literal fixtures) in
 * tests/test_oracle_golden.py::test_c_oracle_matches_numpy_oracle.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* libgomp is not installed in this image, so the row-block parallelism uses plain pthreads */
static int g_threads = 0;
int oracle_num_threads(void) {
  if (g_threads <= 0) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    g_threads = n < 1 ? 1 : (n > 256 ? 256 : (int)n);
  }
  return g_threads;
}
void oracle_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

typedef struct {
  const uint32_t *A, *B;
  uint32_t* C;
  uint32_t* Ap; /* A repacked: [row block of 32][k][32 rows], contiguous, so the inner loop streams 128-byte lines */
  int64_t lda, ldb, ldc, m, k, n, chunk;
  uint64_t N;
  volatile int64_t* next;
  volatile int64_t* next_pack;
} mm_job;

#define MM_IB 32
#define MM_JB 64

static void mm_pack(mm_job* J) {
  const int64_t nib = (J->m + MM_IB - 1) / MM_IB;
  for (;;) {
    const int64_t t = __sync_fetch_and_add(J->next_pack, 1);
    if (t >= nib) break;
    const int64_t i0 = t * MM_IB, ib = (J->m - i0) < MM_IB ? (J->m - i0) : MM_IB;
    uint32_t* dst = J->Ap + t * J->k * MM_IB;
    for (int64_t kk = 0; kk < J->k; ++kk) {
      const uint32_t* a = J->A + kk * J->lda + i0;
      for (int64_t i = 0; i < ib; ++i) dst[kk * MM_IB + i] = a[i];
      for (int64_t i = ib; i < MM_IB; ++i) dst[kk * MM_IB + i] = 0;
    }
  }
}

/* 32 running sums of one column of C over kk in [k0, k1): the loop over i vectorises (zero-extend, 32x32->64 multiply, 64-bit add) */
static inline void mm_kernel(const uint32_t* restrict ap, const uint32_t* restrict bcol, int64_t k0, int64_t k1, uint64_t* restrict ac) {
  uint64_t r[MM_IB];
  for (int i = 0; i < MM_IB; ++i) r[i] = ac[i];
  for (int64_t kk = k0; kk < k1; ++kk) {
    const uint64_t b = bcol[kk];
    const uint32_t* a = ap + kk * MM_IB;
    for (int i = 0; i < MM_IB; ++i) r[i] += (uint64_t)a[i] * b;
  }
  for (int i = 0; i < MM_IB; ++i) ac[i] = r[i];
}

/* Cache-blocked exact product: a task = 32 rows x 64 columns of C.  The packed 32 x chunk block of A (128 KiB at chunk = 1024) is
 * re-read from L2 for each of the 64 columns; `chunk` is also the number of terms a uint64 sum may take before a reduction. */
static void* mm_worker(void* arg) {
  mm_job* J = (mm_job*)arg;
  uint64_t acc[MM_JB][MM_IB];
  const int64_t nib = (J->m + MM_IB - 1) / MM_IB, njb = (J->n + MM_JB - 1) / MM_JB;
  for (;;) {
    const int64_t t = __sync_fetch_and_add(J->next, 1);
    if (t >= nib * njb) break;
    const int64_t bi = t % nib, i0 = bi * MM_IB, j0 = (t / nib) * MM_JB;
    const int64_t ib = (J->m - i0) < MM_IB ? (J->m - i0) : MM_IB, jb = (J->n - j0) < MM_JB ? (J->n - j0) : MM_JB;
    const uint32_t* ap = J->Ap + bi * J->k * MM_IB;
    for (int64_t j = 0; j < jb; ++j)
      for (int i = 0; i < MM_IB; ++i) acc[j][i] = 0;
    for (int64_t k0 = 0; k0 < J->k; k0 += J->chunk) {
      const int64_t k1 = (k0 + J->chunk < J->k) ? k0 + J->chunk : J->k;
      for (int64_t j = 0; j < jb; ++j) {
        mm_kernel(ap, J->B + (j0 + j) * J->ldb, k0, k1, acc[j]);
        for (int i = 0; i < MM_IB; ++i) acc[j][i] %= J->N;
      }
    }
    for (int64_t j = 0; j < jb; ++j)
      for (int64_t i = 0; i < ib; ++i) J->C[(j0 + j) * J->ldc + i0 + i] = (uint32_t)acc[j][i];
  }
  return NULL;
}

static void* mm_pack_worker(void* arg) {
  mm_pack((mm_job*)arg);
  return NULL;
}

/* C = A*B mod N; column-major, A m x k (lda), B k x n (ldb), C m x n (ldc); entries < 2^32, N < 2^32 */
void oracle_matmul_mod(const uint32_t* A, int64_t lda, const uint32_t* B, int64_t ldb, uint32_t* C, int64_t ldc, int64_t m,
                       int64_t k, int64_t n, uint64_t N, uint64_t in_bound) {
  if (m <= 0 || n <= 0) return;
  /* terms are < in_bound^2; a reduced sum (< N) plus `chunk` terms must stay below 2^64 */
  uint64_t R = in_bound ? in_bound : N;
  long double r2 = (long double)(R - 1) * (long double)(R - 1);
  int64_t chunk = r2 < 1 ? (k > 0 ? k : 1) : (int64_t)(9.0e18L / r2);
  if (chunk < 1) chunk = 1;
  if (chunk > 1024) chunk = 1024; /* also the K blocking: 32 rows x 1024 terms x 4 B = 128 KiB of A per block */
  const int64_t nib = (m + MM_IB - 1) / MM_IB;
  uint32_t* Ap = (uint32_t*)malloc((size_t)(nib * (k > 0 ? k : 1) * MM_IB) * sizeof(uint32_t));
  volatile int64_t next = 0, next_pack = 0;
  mm_job job = {A, B, C, Ap, lda, ldb, ldc, m, k, n, chunk, N, &next, &next_pack};
  const int nt = oracle_num_threads();
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nt);
  for (int t = 1; t < nt; ++t) pthread_create(&th[t], NULL, mm_pack_worker, &job);
  mm_pack(&job);
  for (int t = 1; t < nt; ++t) pthread_join(th[t], NULL);
  for (int t = 1; t < nt; ++t) pthread_create(&th[t], NULL, mm_worker, &job);
  mm_worker(&job);
  for (int t = 1; t < nt; ++t) pthread_join(th[t], NULL);
  free(th);
  free(Ap);
}

static uint64_t mod_inv_u64(uint64_t p, uint64_t P) { /* pluq_kernels.jl:11-31 */
  int64_t inv = 0, new_inv = 1, rem = (int64_t)P, new_rem = (int64_t)(p % P);
  while (new_rem != 0) {
    int64_t q = rem / new_rem, t = inv - q * new_inv;
    inv = new_inv; new_inv = t;
    t = rem - q * new_rem; rem = new_rem; new_rem = t;
  }
  if (inv < 0) inv += (int64_t)P;
  return (uint64_t)inv;
}

/* Row echelon elimination in place (W m x n column-major, ld), L m x m (ldl, zero on entry).  Returns the rank;
 * pivcol[t], swp[t] (0-based row exchanged with row t).  Same conventions as oracle.py:echelon. */
int64_t oracle_echelon(uint32_t* W, int64_t ld, uint32_t* L, int64_t ldl, int64_t m, int64_t n, uint64_t N, int64_t* pivcol,
                       int64_t* swp) {
  int64_t row = 0;
  for (int64_t col = 0; col < n && row < m; ++col) {
    uint32_t best = 0; int64_t bi = -1;
    for (int64_t i = row; i < m; ++i) { uint32_t v = W[col * ld + i]; if (v > best) { best = v; bi = i; } }
    if (best == 0) continue;
    const uint64_t pinv = mod_inv_u64(best, N);
    /* swap rows row <-> bi over all columns of W and L; scale the pivot row */
    for (int64_t c = 0; c < n; ++c) {
      uint32_t t = W[c * ld + bi]; W[c * ld + bi] = W[c * ld + row]; W[c * ld + row] = (uint32_t)(((uint64_t)t * pinv) % N);
    }
    for (int64_t c = 0; c < m; ++c) { uint32_t t = L[c * ldl + bi]; L[c * ldl + bi] = L[c * ldl + row]; L[c * ldl + row] = t; }
    L[row * ldl + row] = best;
    for (int64_t i = row + 1; i < m; ++i) { L[row * ldl + i] = W[col * ld + i]; W[col * ld + i] = 0; }
    /* w[i] -= l[i] * u  ==  w[i] += l[i] * (N - u)  (mod N): x = w + l * (N - u) <= (N - 1) * N < 2^64 for N <= 2^32, reduced with a
     * precomputed reciprocal (quotient estimate at most 2 short) instead of two hardware divisions per element */
    const uint64_t mu = N > 1 ? (uint64_t)(~(uint64_t)0 / N) : 0;
    for (int64_t c = col + 1; c < n; ++c) {
      const uint64_t u = W[c * ld + row];
      if (u == 0) continue;
      const uint64_t nu = N - u;
      uint32_t* wc = W + c * ld;
      const uint32_t* lc = L + row * ldl;
      for (int64_t i = row + 1; i < m; ++i) {
        const uint64_t x = (uint64_t)wc[i] + (uint64_t)lc[i] * nu;
        const uint64_t q = (uint64_t)(((unsigned __int128)x * mu) >> 64);
        uint64_t r = x - q * N;
        if (r >= N) r -= N;
        if (r >= N) r -= N;
        wc[i] = (uint32_t)r;
      }
    }
    pivcol[row] = col; swp[row] = bi;
    ++row;
  }
  return row;
}
