// gemm_internal.cuh -- narrow internal interface of the tensor-core GEMM (gemm_tc.cu) used by the multi-GPU layer (mg.cu):
// the 8-bit operand planes of B are produced somewhere else (another GPU) and handed to the GEMM as a finished buffer.
#pragma once
#include "common.cuh"

// Everything that fixes the encoding of one product chunk: limb count / RNS moduli for inputs < R over an inner dimension kc,
// reduction modulus P, the epilogue mode.  All ranks derive the same plan from the same arguments, so planes split on one GPU can
// be multiplied on another.
struct GemmBPlan;

struct BPlaneSpec {
  int nplanes;   // 8-bit planes per operand
  int BN;        // GEMM tile width in columns of B: panel boundaries must be multiples of it
  int64_t Kp;    // padded inner dimension (bytes per plane row)
  int rns;       // 1: residue planes + CRT kernel, 0: positional limbs
};

// crt_mode: GFFM_GEMM_STORE / ADD / SUB applied when C is written.  kara_N1 != 0: P = N1*N2 with the Karatsuba carry split
// (C <- x mod N1, hi <- x div N1).  balanced: inputs may be shifted into (-R/2, R/2] (needs P | R).
int32_t gffm_bplan_create(int64_t kc, uint64_t R, uint64_t P, bool balanced, int crt_mode, uint64_t kara_N1, GemmBPlan** out);
void gffm_bplan_destroy(GemmBPlan* plan);
const BPlaneSpec* gffm_bplan_spec(const GemmBPlan* plan);

// Split columns [0, B.cols) of the view B (kc x nc, + optional second addend: planes of B + B2) into rows [row0, row0 + nc) of the
// plane buffer `planes` laid out [nplanes][rowsPB][Kp], on stream st.
int32_t gffm_bplan_split(gffm_ctx* ctx, const GemmBPlan* plan, MatView B, const MatView* B2, uint8_t* planes, int64_t rowsPB, int64_t row0,
                         cudaStream_t st);

// Fused split + push: the same split, every 16-byte chunk stored to `nbufs` plane buffers (local and peer memory) of identical layout.
// The kernel runs on `push_sms` dedicated SMs (one 1024-thread CTA each that no GEMM CTA can share an SM with).
int32_t gffm_bplan_split_push(gffm_ctx* ctx, const GemmBPlan* plan, MatView B, const MatView* B2, uint8_t* const* plane_bufs, int nbufs, int64_t rowsPB,
                              int64_t row0, cudaStream_t st, int push_sms);

struct ExtBPlanes {
  const uint8_t* planes = nullptr;  // [nplanes][rowsPB][Kp]
  int64_t rowsPB = 0;
  int npanels = 0;
  const int64_t* off = nullptr;         // npanels + 1 column offsets; off[p] is a multiple of BN, off[npanels] == n
  const int* order = nullptr;           // processing order of the panels (nullptr: 0, 1, ...)
  cudaEvent_t const* ready = nullptr;   // ready[p] (may be null): the planes of panel p are complete
};

// C (crt_mode)= A * B mod P with B given as finished planes.  A (m x kc view, optional second addend A2) is split locally (plane
// cache applies).  Per panel, in `order`: wait ready[p], tensor-core GEMM on the context stream, CRT (RNS) on the auxiliary stream.
// On return the context stream has been made to wait for the auxiliary stream.  kara_hi / ldhi: carry output of the Karatsuba split.
int32_t gffm_bplan_gemm(gffm_ctx* ctx, const GemmBPlan* plan, MatView C, MatView A, const MatView* A2, int64_t n, const ExtBPlanes& ext,
                        uint32_t* kara_hi, int64_t ldhi);

// profiling hook: forget the GEMM launch events of earlier products (gffm_last_timings then reports the launches since)
void gffm_profile_reset_tiles(gffm_ctx* ctx);
// Karatsuba recombination on raw buffers (karatsuba.cu): C2 = (P2 - P1 - P3 + carry) mod N2, P1 mod N2 taken from C1
int32_t gffm_kara_recombine(gffm_ctx* ctx, MatView C2, MatView C1, const uint32_t* P2, const uint32_t* P3, const uint32_t* carry, int64_t ldt,
                            uint64_t N2);
