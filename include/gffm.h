/*
 * gffm.h -- C ABI of libgffm.so, the B200-native (sm_100a) hot path behind the
 * GPUFiniteFieldMatrices.jl API (CuModMatrix / KaratsubaMatrix).
 *
 * The reference has no FFI boundary of its own (it is 100% Julia + CUDA.jl); the boundary this
 * library sits behind is the package's exported Julia API (src/GPUFiniteFieldMatrices.jl:36-60).
 * Each entry point below names the reference method(s) it replaces (paths relative to the reference
 * repository root).  The Julia shim that binds these with `ccall` is in
 * gpufinitefieldmatrices.jl_b200/julia/GPUFiniteFieldMatricesB200.jl; the ctypes mirror used by the
 * tests/bench is gpufinitefieldmatrices.jl_b200/capi.py.  See INTEGRATION.md.
 *
 * Conventions
 *  - every function returns an int32 status (GFFM_OK == 0); gffm_last_error() gives the message of the
 *    last failing call on this thread.
 *  - matrices are opaque handles; the library owns device memory.  Host buffers are caller-owned,
 *    COLUMN-MAJOR (Julia layout) with leading dimension `ld` in elements, only touched during the call.
 *  - device storage: canonical residues in [0,N) as uint32, column-major, logical size rows x cols plus the
 *    reference's +32 zero padding per dimension (src/CuModMatrix/CuModMatrix.jl:61-67); padding is always 0.
 *  - calls are enqueued on the context's CUDA stream; download / get_elem / pluq outputs / gffm_sync block.
 *  - indices and permutation tuples crossing this ABI are 1-based (Julia convention) where noted.
 *  - there is NO CPU fallback: without a CUDA device gffm_create fails with GFFM_ERR_NO_DEVICE.
 */
#ifndef GFFM_H
#define GFFM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gffm_ctx gffm_ctx;
typedef struct gffm_mat gffm_mat;

/* status codes -> Julia exceptions (src/CuModMatrix/CuModMatrix.jl:5-31) */
enum {
  GFFM_OK = 0,
  GFFM_ERR_INVALID = 1,            /* ArgumentError */
  GFFM_ERR_SIZE_MISMATCH = 2,      /* CuModArraySizeMismatchException */
  GFFM_ERR_MODULUS_MISMATCH = 3,   /* CuModArrayModulusMismatchException */
  GFFM_ERR_MODULUS_TOO_LARGE = 4,  /* CuModArrayModulusMismatchException (N > 2^52), CuModMatrix.jl:55-59 */
  GFFM_ERR_NOT_SQUARE = 5,         /* CuModMatrixNotSquareException */
  GFFM_ERR_NOT_INVERTIBLE = 6,     /* MatrixNotInvertibleException (undefined in the reference, :485) */
  GFFM_ERR_INVERSE_NOT_DEFINED = 7,/* InverseNotDefinedException, triangular_inverse_no_copy.jl:463 */
  GFFM_ERR_CUDA = 8,
  GFFM_ERR_NO_DEVICE = 9,
  GFFM_ERR_UNSUPPORTED = 10,
  GFFM_ERR_INEXACT = 11,           /* InexactError from convert.(T,A), CuModMatrix.jl:70-86 */
  GFFM_ERR_OOM = 12,
  GFFM_ERR_MODULUS_NOT_PRIME = 13  /* CuModMatrixModulusNotPrimeException (CuModMatrix.jl:23-25): a pivot is a zero divisor */
};

/* host element types for upload/download */
enum { GFFM_F32 = 0, GFFM_F64 = 1, GFFM_I64 = 2, GFFM_U32 = 3, GFFM_I32 = 4 };

/* elementwise ops, gffm_ewise (src/CuModMatrix/kernel_ops/{add,sub,mul,div,mod}_ops.jl) */
enum {
  GFFM_EW_MOD = 0,        /* C = mod(A, m)            mod_ops.jl:3-27  (mod_elements!) */
  GFFM_EW_ADD = 1,        /* C = mod(A + B, m)        add_ops.jl:23-30 */
  GFFM_EW_SUB = 2,        /* C = mod(A - B, m)        sub_ops.jl:33-40 */
  GFFM_EW_MUL = 3,        /* C = mod(A .* B, m)       mul_ops.jl:23-30 (elementwise_multiply!) */
  GFFM_EW_SADD = 4,       /* C = mod(A + s, m)        add_ops.jl:33-40 */
  GFFM_EW_SSUB = 5,       /* C = mod(A - s, m)        sub_ops.jl:43-50 */
  GFFM_EW_RSSUB = 6,      /* C = mod(s - A, m)        sub_ops.jl:53-69 (rscalar_sub!, negate!) */
  GFFM_EW_SMUL = 7,       /* C = mod(A * s, m)        mul_ops.jl:33-40 */
  GFFM_EW_SDIV = 8        /* C = mod(A * s^-1, m)     div_ops.jl:13-22 */
};

/* gffm_gemm accumulate modes */
enum { GFFM_GEMM_STORE = 0, GFFM_GEMM_ADD = 1, GFFM_GEMM_SUB = 2 };

/* gffm_gemm algorithm selector (0 = automatic) */
enum {
  GFFM_ALGO_AUTO = 0,
  GFFM_ALGO_SIMT = 1,   /* scalar uint64 kernel (small shapes, cross-check) */
  GFFM_ALGO_LIMB = 2,   /* positional 8-bit limbs, tcgen05 kind::i8, inputs < 2^16 */
  GFFM_ALGO_RNS = 3     /* residue limbs mod 8-bit moduli, tcgen05 kind::i8, CRT epilogue kernel */
};

/* column-pivot behaviour of gffm_pluq */
enum {
  GFFM_PIVOT_CORRECT = 0,         /* rank-revealing, valid factorisation P*A*Q = L*U */
  GFFM_PIVOT_REFERENCE_QUIRK = 1  /* literal pluq_kernels.jl:88,103,193 behaviour (swap with fixed last column, skip) */
};

/* ---- library / context ------------------------------------------------------------------ */
const char* gffm_version(void);
const char* gffm_last_error(void);
int32_t gffm_device_count(int32_t* count);

/* one context per Julia task/thread; owns a stream + workspaces.  Replaces CUDA.jl's task-local state. */
int32_t gffm_create(int32_t device, gffm_ctx** out);
int32_t gffm_destroy(gffm_ctx* ctx);
int32_t gffm_sync(gffm_ctx* ctx);                         /* CUDA.synchronize / CUDA.@sync */
int32_t gffm_set_stream(gffm_ctx* ctx, void* cuda_stream); /* adopt an external cudaStream_t (e.g. torch's) */
int32_t gffm_get_stream(gffm_ctx* ctx, void** cuda_stream);
/* phase profiling: when on, gemm calls record CUDA events around their phases on the context stream */
int32_t gffm_set_profiling(gffm_ctx* ctx, int32_t on);
/* CUDA-event timings (ms) of the phases of the last profiled gemm call -- RNS: {plane split, tcgen05 GEMM kernel, CRT
 * kernel}; LIMB: {plane split, tcgen05 GEMM kernel}; tiled multi-stream products (large shapes): {0, sum of the GEMM
 * launch durations, number of GEMM launches}; eliminations: {panel phase, triangular solves, Schur GEMMs}.  Blocks until
 * the events have completed. */
int32_t gffm_last_timings(gffm_ctx* ctx, double* ms, int32_t capacity, int32_t* n_written);
/* Cap on the number of persistent CTAs of the tensor-core GEMM (0 = one per SM, the default).  The GEMM CTA owns its SM's
 * shared memory, so kernels of OTHER libraries (NCCL's broadcast in the multi-GPU layer) can only run on SMs it leaves free. */
int32_t gffm_set_gemm_ctas(gffm_ctx* ctx, int32_t ctas);
/* number of library kernels launched on this context since creation (bench.py's gpu_launches) */
int32_t gffm_launch_count(gffm_ctx* ctx, int64_t* count);
/* device memory the library has requested on this context since creation (matrices, plane caches, workspaces): cumulative
 * bytes and number of requests.  The difference across a call is what CUDA.@timed's gpu_bytes reports for the reference
 * (test/CuModMatrix/allocations_test.jl:21-52: 0 for the in-place elementwise methods, < 20 for a warm mul!) */
int32_t gffm_alloc_stats(gffm_ctx* ctx, int64_t* bytes, int64_t* calls);

/* ---- container: struct CuModArray + ctors, CuModMatrix.jl:42-181, :510-556 ------------------------ */
/* zeros(T, rows, cols, N) with +pad zero slack (CuModMatrix.jl:530-534); pad < 0 selects the reference's 32 */
int32_t gffm_mat_create(gffm_ctx* ctx, int64_t rows, int64_t cols, uint64_t N, int32_t pad, gffm_mat** out);
/* device-wrapper ctor CuModMatrix(::CuArray, N) (CuModMatrix.jl:113-121): adopt external uint32 column-major
 * device memory (not owned, no padding assumed) */
int32_t gffm_mat_wrap(gffm_ctx* ctx, void* device_u32, int64_t rows, int64_t cols, int64_t ld, uint64_t N,
                      gffm_mat** out);
/* A matrix used as a GEMM operand keeps its 8-bit operand planes cached until it is written; this frees them early
 * (and makes the next product rebuild them -- bench.py uses it so that every timed step is a fresh product). */
int32_t gffm_mat_drop_cache(gffm_mat* m);
/* declare that the matrix was modified through its raw device pointer (gffm_mat_device_ptr): invalidates cached planes */
int32_t gffm_mat_touch(gffm_mat* m);
int32_t gffm_mat_destroy(gffm_mat* m); /* safe from a Julia finalizer thread: stream-ordered free, no callbacks */
/* host ctor CuModMatrix(A, N; mod) (CuModMatrix.jl:53-99): convert (exactness checked -> GFFM_ERR_INEXACT),
 * copy into the top-left corner, floored mod when do_mod != 0 */
int32_t gffm_mat_upload(gffm_mat* m, const void* host, int32_t dtype, int64_t ld, int32_t do_mod);
/* Array(A) (:256-261) when with_padding == 0, unsafe_Array(A) (:251-253) otherwise ((rows+pad) x (cols+pad)) */
int32_t gffm_mat_download(gffm_mat* m, void* host, int32_t dtype, int64_t ld, int32_t with_padding);
int32_t gffm_mat_rows(gffm_mat* m, int64_t* rows);
int32_t gffm_mat_cols(gffm_mat* m, int64_t* cols);
int32_t gffm_mat_pad(gffm_mat* m, int32_t* pad);
int32_t gffm_mat_modulus(gffm_mat* m, uint64_t* N);
int32_t gffm_mat_ld(gffm_mat* m, int64_t* ld);
int32_t gffm_mat_device_ptr(gffm_mat* m, void** ptr);
/* change_modulus / change_modulus_no_alloc! (CuModMatrix.jl:726-760): new N, entries reduced when reduce != 0 */
int32_t gffm_mat_set_modulus(gffm_mat* m, uint64_t N, int32_t reduce);
int32_t gffm_mat_copy(gffm_mat* dst, gffm_mat* src);          /* copy!/copyto! (:640-680) */
int32_t gffm_mat_fill(gffm_mat* m, int64_t value);            /* fill! (:690-700): mod(value, N) everywhere */
int32_t gffm_mat_zero(gffm_mat* m);                           /* zero! */
int32_t gffm_mat_eye(gffm_mat* m);                            /* eye (:510-523) */
int32_t gffm_mat_rand(gffm_mat* m, uint64_t seed);            /* rand (:540-556); counter-based, padding stays 0 */
/* SURVEY 8(d) generator: val = splitmix64(seed ^ (j*rows+i)) mod N (0-based, column-major) */
int32_t gffm_mat_synth(gffm_mat* m, uint64_t seed);
int32_t gffm_mat_get_elem(gffm_mat* m, int64_t i, int64_t j, int64_t* value); /* 0-based */
int32_t gffm_mat_set_elem(gffm_mat* m, int64_t i, int64_t j, int64_t value);  /* 0-based, stored mod N */
int32_t gffm_mat_transpose(gffm_mat* dst, gffm_mat* src);      /* transpose (:300-305) */
/* dst[0:nr,0:nc] = src[r0:r0+nr, c0:c0+nc] (0-based) -- @view / getindex with ranges */
int32_t gffm_mat_copy_block(gffm_mat* dst, int64_t dr0, int64_t dc0, gffm_mat* src, int64_t sr0, int64_t sc0,
                            int64_t nr, int64_t nc);
/* bit-exact comparison on the device (used by tests at sizes the host cannot hold cheaply) */
int32_t gffm_mat_equal(gffm_mat* a, gffm_mat* b, int32_t* equal);
/* 64-bit checksum sum_{ij} (val * splitmix64(i + j*rows)) mod 2^64 -- size-independent parity property */
int32_t gffm_mat_checksum(gffm_mat* a, uint64_t* sum);

/* ---- modular GEMM / GEMV -------------------------------------------------------------------------- */
/* mul!(C,A,B) / A*B / mulN! (CuModMatrix.jl:767-809, kernel_ops/mul_ops.jl:54-58) and stripe_mul!
 * (kernel_mul/stripe_mul.jl:175-244): C (op)= A*B mod P.  in_bound_R = exclusive bound on the inputs
 * (0 -> A's modulus), mod_P = modulus applied (0 -> C's modulus; the `mod_N` override of mat_mul_gpu_type,
 * kernel_mul/mat_mul_gpu_direct.jl:8-42).  mode: GFFM_GEMM_*; algo: GFFM_ALGO_*. */
int32_t gffm_gemm(gffm_mat* C, gffm_mat* A, gffm_mat* B, uint64_t in_bound_R, uint64_t mod_P, int32_t mode,
                  int32_t algo);
/* same on 0-based sub-blocks: C[cr0.., cc0..] (m x n) (op)= A[ar0.., ac0..] (m x k) * B[br0.., bc0..] (k x n) */
int32_t gffm_gemm_block(gffm_mat* C, int64_t cr0, int64_t cc0, gffm_mat* A, int64_t ar0, int64_t ac0,
                        gffm_mat* B, int64_t br0, int64_t bc0, int64_t m, int64_t n, int64_t k,
                        uint64_t in_bound_R, uint64_t mod_P, int32_t mode, int32_t algo);
/* C_host = A_host * B_host mod N directly between HOST buffers (column-major uint32 residues, leading dimensions in
 * elements; pinned memory gives full overlap).  One call replaces CuModMatrix(A); CuModMatrix(B); mul!(C,A,B); Array(C)
 * (CuModMatrix.jl:53-99, :767-787, :256-261) and pipelines H2D copies, plane split, tcgen05 GEMM tiles and D2H copies on
 * three streams.  Blocks until C_host is complete. */
int32_t gffm_gemm_host(gffm_ctx* ctx, void* C_host, int64_t ldc, const void* A_host, int64_t lda, const void* B_host, int64_t ldb,
                       int64_t m, int64_t n, int64_t k, int32_t dtype, uint64_t N);
/* Multi-GPU layer (new -- the reference is single-GPU; north_star (4): products shard by row blocks of A, B is broadcast).
 * One sharded product step on this rank: C = A * B mod P where A, C are this rank's row blocks and B ARRIVES in `npanels`
 * column panels [col_off[p], col_off[p+1]) (0-based, col_off has npanels+1 entries covering [0, cols(B)); interior offsets
 * that are multiples of 256 take the pipelined path).  ready[p] (cudaEvent_t or NULL; the array may be NULL) has been
 * recorded by the caller on the stream that delivers panel p (e.g. after ncclBroadcast); the plane split of panel p+1 and the
 * CRT of panel p run on an internal stream while the tensor-core GEMM of panel p runs on the context stream.  consumed[p]
 * (optional) is recorded once panel p has been read for the last time, so the caller may already overwrite it with the next
 * step's data.  Semantics of the result are those of mul!(C,A,B) (CuModMatrix.jl:767-787). */
int32_t gffm_gemm_panels(gffm_mat* C, gffm_mat* A, gffm_mat* B, int32_t npanels, const int64_t* col_off, void* const* ready,
                         void* const* consumed, uint64_t in_bound_R, uint64_t mod_P);
/* ---- multi-GPU layer through the C ABI (new): one rank per GPU, NCCL for the bootstrap and as a transport,
 * copy engines over NVLink peer memory as the default transport.  Products shard by row blocks of A and C; B lives on `root`.
 * What travels is B's 8-bit operand planes: the column range owned by rank q is split by rank q only and collected by the
 * others while their GEMMs already run (csrc/mg.cu).  All calls are collective (every rank, same order, same shapes). ---- */
typedef struct gffm_mg gffm_mg;
enum {
  GFFM_MG_AUTO = 0,         /* peer-memory planes where CUDA IPC / peer access works between all ranks, else NCCL planes */
  GFFM_MG_NCCL_BCAST = 1,   /* ncclBroadcast of B's uint32 column ranges, every rank splits all of B */
  GFFM_MG_NCCL_PLANES = 2,  /* grouped ncclSend/ncclRecv scatter of the ranges, owner split, grouped in-place ncclAllGather of the planes */
  GFFM_MG_P2P_PLANES = 3,   /* copy-engine push of the ranges / pull of the planes through peer memory, epoch flags, no SM used */
  GFFM_MG_P2P_PUSH = 4,     /* fused split + push: the owner's split kernel stores its planes into every rank's buffer (NVLink stores) */
  GFFM_MG_P2P_RAW = 5       /* copy-engine scatter + forward of the uint32 ranges (4 bytes per element), every rank splits every range locally */
};
/* root argument of gffm_mg_gemm for a B that is ALREADY distributed: every rank passes its own column range
 * [off[rank], off[rank+1]) (gffm_mg_owner_ranges) as a k x width matrix -- e.g. uploaded from the host over that GPU's own PCIe link */
enum { GFFM_MG_DISTRIBUTED = -1 };
/* 128-byte NCCL unique id: create on one rank, hand to the others by any channel (MPI, a file, Distributed.jl, torch.distributed) */
int32_t gffm_mg_unique_id(void* id128);
/* ctx = this rank's context (its device).  NCCL is dlopen'ed (libnccl.so.2, or $GFFM_NCCL_LIB) */
int32_t gffm_mg_create(gffm_ctx* ctx, const void* id128, int32_t nranks, int32_t rank, gffm_mg** out);
int32_t gffm_mg_destroy(gffm_mg* mg);
/* peer_memory: bit 0 = peer mappings established, bit 1 = flag waits are stream memory operations, bit 2 = flag writes are too */
int32_t gffm_mg_info(gffm_mg* mg, int32_t* rank, int32_t* nranks, int32_t* transport, int32_t* peer_memory);
int32_t gffm_mg_set_transport(gffm_mg* mg, int32_t transport);
/* drains this rank's streams, meets the other ranks, reports a peer-flag time-out of an earlier product */
int32_t gffm_mg_barrier(gffm_mg* mg);
/* column ranges of B owned by the ranks: off[0..nranks] (equal widths, multiples of 256) -- pure function, no GPU needed */
int32_t gffm_mg_owner_ranges(int64_t n, int32_t nranks, int64_t* off);
/* the ranges the peer-memory transports use from GFFM_MG_ROOT_FREE_MIN (default 6) ranks on when B lives on `root`: the root owns
 * nothing (its NVLink egress already carries everybody else's uint32 ranges), the 256-column blocks are dealt to the other ranks in
 * rank order; off[root + 1] == off[root] -- pure function, no GPU needed */
int32_t gffm_mg_owner_ranges_root_free(int64_t n, int32_t nranks, int32_t root, int64_t* off);
/* C_shard = A_shard * B mod P: mul!(C,A,B) (CuModMatrix.jl:767-787) on row blocks.  B holds the matrix on `root`; on the other
 * ranks it is a same-shape matrix created the same way (receive buffer of the broadcast transport, otherwise untouched).
 * b_ready_event: cudaEvent_t recorded after B's last modification on root, or NULL = B is ready in context-stream order.  With an
 * event the distribution of this product may run under the GEMMs of the previous one.  Asynchronous like gffm_gemm. */
int32_t gffm_mg_gemm(gffm_mg* mg, gffm_mat* C, gffm_mat* A, gffm_mat* B, int32_t root, void* b_ready_event, uint64_t in_bound_R, uint64_t mod_P);
/* KMatMul!(C,A,B) (KaratsubaMatrix.jl:133-204) on row blocks: A1, A2, C1, C2 are this rank's row blocks, B1, B2 live on root;
 * inner dimension at most 65536 (and within one limb chunk for the P2 / P3 sub-products) */
int32_t gffm_mg_kmat_mul(gffm_mg* mg, gffm_mat* C1, gffm_mat* C2, gffm_mat* A1, gffm_mat* A2, gffm_mat* B1, gffm_mat* B2, uint64_t N1,
                         uint64_t N2, int32_t root, void* b_ready_event);
/* z_shard = A_shard * x mod P: mul!(z,A,x) (CuModMatrix.jl:816-836) on row blocks; x (n x 1 on every rank) is broadcast from root */
int32_t gffm_mg_gemv(gffm_mg* mg, gffm_mat* z, gffm_mat* A, gffm_mat* x, int32_t root, uint64_t in_bound_R, uint64_t mod_P);

/* mul!(z,A,x;R,P) (CuModMatrix.jl:816-836, stripe_mul.jl:82-168): z = A*x mod P, x and z are n x 1 matrices */
int32_t gffm_gemv(gffm_mat* z, gffm_mat* A, gffm_mat* x, uint64_t in_bound_R, uint64_t mod_P);

/* ---- elementwise (kernel_ops/{add,sub,mul,div,mod}_ops.jl; mod_N override semantics of inplace_operations_test.jl:125-191) ----- */
int32_t gffm_ewise(int32_t op, gffm_mat* C, gffm_mat* A, gffm_mat* B_or_null, int64_t scalar, uint64_t mod_override);

/* ---- elimination ------------------------------------------------------------------------------------ */
/* pluq_gpu_kernel(A) -> (U, L, Perm_rows, Perm_cols) (rref_lu_pluq/pluq_kernels.jl:46-157).
 * U: rows x cols, L: rows x rows are created by the call (caller destroys).  perm buffers are caller-allocated
 * with capacity >= min(rows,cols) (rows) / cols (cols) PAIRS of int64 (1-based tuples, in application order). */
int32_t gffm_pluq(gffm_mat* A, gffm_mat** U, gffm_mat** L, int64_t* prow_pairs, int64_t* n_prow,
                  int64_t* pcol_pairs, int64_t* n_pcol, int64_t* rank, int32_t col_pivot_mode);
/* lu(A) -> (U, L, Perm): intended lu_gpu_type (test/Experiments/rref_gpu_type.jl:60-103); U = row echelon form,
 * pivcols (capacity min(rows,cols), 0-based) lists the pivot columns */
int32_t gffm_lu(gffm_mat* A, gffm_mat** U, gffm_mat** L, int64_t* prow_pairs, int64_t* n_prow,
                int64_t* pivcols, int64_t* rank);
/* rref(A): intended rref_gpu_type (test/Experiments/rref_gpu_type.jl:8-51); unique reduced row echelon form */
int32_t gffm_rref(gffm_mat* A, gffm_mat** R, int64_t* pivcols, int64_t* rank);
int32_t gffm_rank(gffm_mat* A, int64_t* rank);
/* inverse / is_invertible_with_inverse (CuModMatrix.jl:356-422, :480-502): *invertible = 0 and *Ainv = NULL for a
 * singular matrix (status stays GFFM_OK; the Julia shim turns that into (false, nothing) or the exception) */
int32_t gffm_inverse(gffm_mat* A, gffm_mat** Ainv, int32_t* invertible);
/* upper_/lower_triangular_inverse_no_copy (triangular/triangular_inverse_no_copy.jl:197-228, :450-478) */
int32_t gffm_triinv(gffm_mat* A, int32_t upper, gffm_mat** out);
/* apply_{row,col}_{,inv_}perm! (rref_lu_pluq/permutations.jl:11-27, :73-89): ordered 1-based transposition list */
int32_t gffm_apply_perm(gffm_mat* A, const int64_t* pairs, int64_t n_pairs, int32_t on_cols, int32_t inverse);
/* mod_inv (pluq_kernels.jl:11-31), batched on the device: out[i] = in[i]^-1 mod N (0 where not invertible) */
int32_t gffm_modinv_batch(gffm_ctx* ctx, const uint64_t* in_host, uint64_t* out_host, int64_t n, uint64_t N);

/* ---- Karatsuba two-limb product (KaratsubaMatrix/KaratsubaMatrix.jl:133-204, KaratsubaKernels.jl:129-158) ----
 * (C1 + N1*C2) = (A1 + N1*A2) * (B1 + N1*B2) mod N1*N2, requires N2 | N1.  B may be n x 1 (KMatMul! on vectors,
 * KMatMul_gemv! :238-300). */
int32_t gffm_kmat_mul(gffm_mat* C1, gffm_mat* C2, gffm_mat* A1, gffm_mat* A2, gffm_mat* B1, gffm_mat* B2,
                      uint64_t N1, uint64_t N2);
/* Karatsuba elementwise with carry (KaratsubaKernels.jl:2-125): op = GFFM_EW_ADD / SUB / SMUL / RSSUB(negate) */
int32_t gffm_kmat_ewise(int32_t op, gffm_mat* C1, gffm_mat* C2, gffm_mat* A1, gffm_mat* A2, gffm_mat* B1_or_null,
                        gffm_mat* B2_or_null, int64_t scalar, uint64_t N1, uint64_t N2);

#ifdef __cplusplus
}
#endif
#endif /* GFFM_H */
