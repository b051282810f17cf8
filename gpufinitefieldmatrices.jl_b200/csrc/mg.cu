// mg.cu -- multi-GPU layer of libgffm (new: the reference is single-GPU; BASELINE north_star (4)).
//
// One rank per GPU (one process per GPU, or several contexts in one process).  Products shard by ROW BLOCKS of A and C; B lives
// on a root rank.  What travels between GPUs is not B but its 8-BIT OPERAND PLANES: the column range of B owned by rank q is
// split into planes by rank q only (1/G of the split work per GPU instead of all of it on every GPU), and every rank collects the
// other ranges while its tensor-core GEMM already multiplies the ranges that have arrived.  Three transports:
//
//   GFFM_MG_P2P_PLANES  (default where peer access works)  copy engines over NVLink peer memory, no SM of any GPU is used for
//                       communication: the root PUSHES each rank's uint32 column range into that rank's staging buffer, every
//                       rank splits its range into planes, every rank PULLS the other ranges' planes from their owners.  Ranks
//                       synchronise through 32-bit epoch flags in each other's memory, written and awaited with STREAM MEMORY
//                       OPERATIONS (cuStreamWriteValue32 / cuStreamWaitValue32: executed by the GPU front end, not by an SM) --
//                       stream-ordered on both sides, no host round trip, no NCCL or polling kernel competing with the persistent
//                       GEMM (measured: a one-warp polling kernel sharing an SM with a GEMM CTA doubled that launch's duration
//                       through the static tile schedule; profiles/r02_notes.md).  Kernel-based flags (bounded polling) remain as
//                       the fallback where the memory operations are not available (GFFM_MG_FLAGS=kernel forces them).
//                       Transfers are spread over several copy streams (one copy engine each).
//   GFFM_MG_P2P_PUSH    the owner's split kernel stores every plane chunk to ALL ranks' plane buffers at once (fused split + push,
//                       posted NVLink stores from the kernel that creates the planes): no staging copy of the planes, no pull, the
//                       copy engines only carry the root's uint32 ranges.  Consumers wait for the owner's epoch flag, owners wait
//                       for the consumers' "buffer free" flags.
//   GFFM_MG_P2P_RAW     what travels is the uint32 residues (4 bytes per element instead of 8 one-byte planes): the root's copy engines
//                       push every owner's range into that owner's staging buffer, every owner's copy engines FORWARD the range to all
//                       other ranks' staging buffers (an all-gather by posted peer writes; the root needs nothing, it holds B), and
//                       every rank splits every range locally as it arrives, under its own GEMMs.  No SM of any GPU is used for
//                       communication and the GEMM keeps all SMs; the price is the replicated split (HBM-bound, 0.6 ms per 16384^2
//                       operand when alone).  Half the NVLink bytes of the plane transports.
//   GFFM_MG_NCCL_PLANES the same data flow with NCCL: grouped ncclSend/ncclRecv scatter of the uint32 ranges, one grouped
//                       in-place ncclAllGather of the planes.
//   GFFM_MG_NCCL_BCAST  ncclBroadcast of B's uint32 column ranges, every rank splits all of B (round-1 data flow).
//
// Plane buffers and staging are triple-buffered by epoch, so the distribution of product e+1 runs under the GEMMs of products
// e-1 and e when the caller says that B is ready (b_ready event).  NCCL is loaded with dlopen (libnccl.so.2): the
// library has no link-time dependency on it and shares the copy a host runtime (e.g. torch) has already loaded.
#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <unistd.h>
#include <algorithm>
#include "gemm_internal.cuh"

namespace {

constexpr int MG_MAX_RANKS = 32;
// Plane / staging buffers per rank: the distribution of product e+1 may start while product e-1 is still being multiplied and has two
// whole products' worth of time to complete (measured on 8 B200: with two buffers it could only start when EVERY rank had finished
// product e-1, and its ~3 ms did not fit under one ~3.4 ms product once the ranks' skew was added; profiles/r02_notes.md)
constexpr int MG_NBUF = 3;
constexpr size_t MG_CTL_BYTES = 4096;
// control words (uint32) at the start of every rank's arena
enum { F_STAGED = 0, F_READY = 64, F_PULLED = 128, F_SPLIT_DONE = 192, F_ERROR = 256, F_PROBE = 320, F_FREE = 384 };

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  if (api.h) return &api;
  const char* names[] = {getenv("GFFM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return nullptr;
#define MG_SYM(field, name)                                     \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name)); \
  if (!api.field) return nullptr;
  MG_SYM(GetVersion, "ncclGetVersion")
  MG_SYM(GetUniqueId, "ncclGetUniqueId")
  MG_SYM(CommInitRank, "ncclCommInitRank")
  MG_SYM(CommDestroy, "ncclCommDestroy")
  MG_SYM(Broadcast, "ncclBroadcast")
  MG_SYM(AllGather, "ncclAllGather")
  MG_SYM(AllReduce, "ncclAllReduce")
  MG_SYM(Send, "ncclSend")
  MG_SYM(Recv, "ncclRecv")
  MG_SYM(GroupStart, "ncclGroupStart")
  MG_SYM(GroupEnd, "ncclGroupEnd")
  MG_SYM(GetErrorString, "ncclGetErrorString")
#undef MG_SYM
  api.h = h;
  return &api;
}

// stream memory operations of the driver API, resolved through the runtime (no link-time dependency on libcuda)
struct MemOps {
  CUresult (*wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
  CUresult (*write32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
};
MemOps* mem_ops() {
  static MemOps ops;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *w = nullptr, *v = nullptr;
    cudaDriverEntryPointQueryResult q1, q2;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &w, cudaEnableDefault, &q1) == cudaSuccess && q1 == cudaDriverEntryPointSuccess &&
        cudaGetDriverEntryPoint("cuStreamWriteValue32", &v, cudaEnableDefault, &q2) == cudaSuccess && q2 == cudaDriverEntryPointSuccess) {
      ops.wait32 = reinterpret_cast<decltype(ops.wait32)>(w);
      ops.write32 = reinterpret_cast<decltype(ops.write32)>(v);
    }
    cudaGetLastError();
  }
  return (ops.wait32 && ops.write32) ? &ops : nullptr;
}

#define MG_NCCL(expr)                                                                                              \
  do {                                                                                                             \
    ncclResult_t _r = (expr);                                                                                      \
    if (_r != ncclSuccess) {                                                                                       \
      gffm_set_error("NCCL error %d at %s:%d: %s", (int)_r, __FILE__, __LINE__, nccl_api()->GetErrorString(_r)); \
      return GFFM_ERR_CUDA;                                                                                        \
    }                                                                                                              \
  } while (0)

// ---- device-side flag primitives ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// one warp: lane i (if mask bit i is set) polls flags[i] until (int32)(flags[i] - target) >= 0.  Bounded: after timeout_ns the
// kernel raises *err and returns, so a lost peer turns into a reported error instead of a hung GPU.
__global__ void mg_wait_kernel(const uint32_t* __restrict__ flags, uint32_t mask, uint32_t target, uint32_t* err, unsigned long long timeout_ns) {
  const int lane = threadIdx.x;
  bool ok = ((mask >> lane) & 1u) == 0;
  const unsigned long long t0 = global_timer_ns();
  const volatile uint32_t* vf = flags;
  for (;;) {
    if (!ok) ok = (int32_t)(vf[lane] - target) >= 0;  // relaxed polling; one acquire fence at the end
    if (__all_sync(0xffffffffu, ok)) break;
    if (global_timer_ns() - t0 > timeout_ns) {
      if (!ok) atomicExch(err, 1u + (uint32_t)lane);
      break;
    }
    __nanosleep(2000);
  }
  (void)ld_acquire_sys(flags + lane);
  __threadfence_system();
}

struct MgTargets {
  uint32_t* p[MG_MAX_RANKS + 1];
};
// writes `value` to every target (local or peer memory) after a system-wide fence: everything this stream did before the kernel
// (split kernels, copy-engine transfers) is visible to whoever observes the flag
__global__ void mg_signal_kernel(MgTargets t, int n, uint32_t value) {
  const int i = threadIdx.x;
  __threadfence_system();
  if (i < n && t.p[i]) st_release_sys(t.p[i], value);
}

struct MgXchg {  // what the ranks tell each other about their arenas
  cudaIpcMemHandle_t handle;
  uint64_t ptr;
  uint64_t bytes;
  int32_t pid;
  int32_t dev;
  int32_t ok;
  int32_t pad[9];
};
static_assert(sizeof(MgXchg) == 128, "exchange record is 128 bytes");

}  // namespace

struct gffm_mg {
  gffm_ctx* ctx = nullptr;
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
  int transport = 0;  // resolved GFFM_MG_*
  int requested = 0;
  bool p2p_ok = false;
  static constexpr int NCOPY = 4;  // parallel copy streams (one copy engine each) for the pulls and for the root's pushes
  cudaStream_t s_comm = nullptr, s_dist = nullptr, s_pull[NCOPY] = {}, s_push[NCOPY] = {};
  cudaEvent_t copy_ev[2][NCOPY] = {};  // joins of the copy streams (0: pulls, 1: pushes)
  int ncopy = NCOPY;
  int push_sms = 20;      // GFFM_MG_P2P_PUSH: SMs dedicated to the fused split + push kernel (GFFM_MG_PUSH_SMS)
  int push_ce_peers = 0;  // ... and how many of the peers get their copy from the copy engines instead (GFFM_MG_PUSH_CE_PEERS)
  int root_free_min = 6;  // peer-memory transports: from this many ranks on the root owns no column range of B (0: never; GFFM_MG_ROOT_FREE_MIN)
  int saved_gemm_ctas = -1;
  cudaEvent_t split_ev[MG_NBUF] = {}, ce_done[MG_NBUF] = {};
  bool wait_memops = false, signal_memops = false;  // flags through stream memory operations instead of one-warp kernels
  // arena: [control words | MG_NBUF staging buffers | MG_NBUF plane buffers], one cudaMalloc, exported through CUDA IPC
  char* base = nullptr;
  size_t arena_bytes = 0, stage_bytes = 0, planes_bytes = 0;
  char* peer_base[MG_MAX_RANKS] = {};
  bool peer_ipc[MG_MAX_RANKS] = {};
  uint32_t epoch = 0;
  unsigned long long timeout_ns = 30ull * 1000000000ull;
  // events
  cudaEvent_t gemm_done[MG_NBUF] = {}, stage_free[MG_NBUF] = {}, staged[MG_NBUF] = {}, gathered[MG_NBUF] = {};
  cudaEvent_t ready_ev[MG_NBUF][MG_MAX_RANKS] = {};
  cudaEvent_t bc_ev[MG_MAX_RANKS] = {}, bc_consumed[MG_MAX_RANKS] = {};
  cudaEvent_t ev_call = nullptr, ev_push = nullptr, ev_comm = nullptr;
  void* xchg_dev = nullptr;  // (nranks + 1) * 128 bytes + barrier word
  int64_t rounds = 0;
  char* stage(int b) const { return base + MG_CTL_BYTES + (size_t)b * stage_bytes; }
  char* planes(int b) const { return base + MG_CTL_BYTES + MG_NBUF * stage_bytes + (size_t)b * planes_bytes; }
  size_t planes_off(int b) const { return MG_CTL_BYTES + MG_NBUF * stage_bytes + (size_t)b * planes_bytes; }
  size_t stage_off(int b) const { return MG_CTL_BYTES + (size_t)b * stage_bytes; }
  uint32_t* ctl(int q) const { return reinterpret_cast<uint32_t*>(q == rank ? base : peer_base[q]); }
};

namespace {

int32_t mg_barrier_impl(gffm_mg* mg) {
  NcclApi* nc = nccl_api();
  int* word = reinterpret_cast<int*>((char*)mg->xchg_dev + (size_t)(mg->nranks + 1) * 128);
  MG_NCCL(nc->AllReduce(word, word, 1, ncclInt32, ncclSum, mg->comm, mg->s_comm));
  GFFM_CUDA(cudaStreamSynchronize(mg->s_comm));
  return GFFM_OK;
}

int32_t mg_sync_streams(gffm_mg* mg) {
  GFFM_CUDA(cudaStreamSynchronize(mg->ctx->stream));
  if (mg->ctx->s_aux) GFFM_CUDA(cudaStreamSynchronize(mg->ctx->s_aux));
  GFFM_CUDA(cudaStreamSynchronize(mg->s_dist));
  for (int j = 0; j < gffm_mg::NCOPY; ++j) {
    GFFM_CUDA(cudaStreamSynchronize(mg->s_pull[j]));
    GFFM_CUDA(cudaStreamSynchronize(mg->s_push[j]));
  }
  GFFM_CUDA(cudaStreamSynchronize(mg->s_comm));
  return GFFM_OK;
}

void mg_close_peers(gffm_mg* mg) {
  for (int q = 0; q < mg->nranks; ++q) {
    if (q != mg->rank && mg->peer_base[q] && mg->peer_ipc[q]) cudaIpcCloseMemHandle(mg->peer_base[q]);
    mg->peer_base[q] = nullptr;
    mg->peer_ipc[q] = false;
  }
  cudaGetLastError();
}

// (Re)allocates the arena so that it holds 2 x stage_bytes + 2 x planes_bytes and re-establishes the peer mappings.  Collective:
// every rank reaches the same decision because the sizes derive from arguments all ranks share.
int32_t mg_ensure_arena(gffm_mg* mg, size_t stage_bytes, size_t planes_bytes) {
  stage_bytes = (stage_bytes + 4095) & ~(size_t)4095;
  planes_bytes = (planes_bytes + 4095) & ~(size_t)4095;
  if (mg->base && mg->stage_bytes >= stage_bytes && mg->planes_bytes >= planes_bytes) return GFFM_OK;
  NcclApi* nc = nccl_api();
  GFFM_TRY(mg_sync_streams(mg));
  GFFM_TRY(mg_barrier_impl(mg));  // nobody reads or writes anybody's arena any more
  mg_close_peers(mg);
  GFFM_TRY(mg_barrier_impl(mg));  // every importer has closed its mappings
  if (mg->base) {
    GFFM_CUDA(cudaFree(mg->base));
    mg->base = nullptr;
  }
  mg->stage_bytes = std::max(mg->stage_bytes, stage_bytes);
  mg->planes_bytes = std::max(mg->planes_bytes, planes_bytes);
  mg->arena_bytes = MG_CTL_BYTES + MG_NBUF * (mg->stage_bytes + mg->planes_bytes);
  MgXchg mine;
  memset(&mine, 0, sizeof(mine));
  mine.ok = 1;
  if (cudaMalloc((void**)&mg->base, mg->arena_bytes) != cudaSuccess) {
    cudaGetLastError();
    mg->base = nullptr;
    mine.ok = 0;
  } else {
    GFFM_CUDA(cudaMemset(mg->base, 0, MG_CTL_BYTES));
    if (cudaIpcGetMemHandle(&mine.handle, mg->base) != cudaSuccess) {
      cudaGetLastError();
      memset(&mine.handle, 0, sizeof(mine.handle));
      mine.ok = 2;  // allocated, not exportable
    }
  }
  mine.ptr = (uint64_t)(uintptr_t)mg->base;
  mine.bytes = mg->arena_bytes;
  mine.pid = (int32_t)getpid();
  mine.dev = mg->ctx->device;
  // all-gather of the records
  char* xs = (char*)mg->xchg_dev;
  GFFM_CUDA(cudaMemcpy(xs, &mine, sizeof(mine), cudaMemcpyHostToDevice));
  MG_NCCL(nc->AllGather(xs, xs + 128, 128, ncclUint8, mg->comm, mg->s_comm));
  GFFM_CUDA(cudaStreamSynchronize(mg->s_comm));
  std::vector<MgXchg> all(mg->nranks);
  GFFM_CUDA(cudaMemcpy(all.data(), xs + 128, (size_t)mg->nranks * 128, cudaMemcpyDeviceToHost));
  bool alloc_ok = true, p2p = mg->requested != GFFM_MG_NCCL_BCAST && mg->requested != GFFM_MG_NCCL_PLANES;
  for (int q = 0; q < mg->nranks; ++q) {
    if (all[q].ok == 0) alloc_ok = false;
    if (all[q].ok != 1) p2p = false;
  }
  if (!alloc_ok) GFFM_FAIL(GFFM_ERR_OOM, "multi-GPU arena of %zu bytes could not be allocated on every rank", mg->arena_bytes);
  int my_ok = 1;
  if (p2p) {
    for (int q = 0; q < mg->nranks && my_ok; ++q) {
      if (q == mg->rank) continue;
      if (all[q].pid == mine.pid) {  // same process, another context: plain peer access
        int can = 0;
        cudaDeviceCanAccessPeer(&can, mg->ctx->device, all[q].dev);
        if (!can && all[q].dev != mg->ctx->device) {
          my_ok = 0;
          break;
        }
        if (all[q].dev != mg->ctx->device) {
          cudaError_t e = cudaDeviceEnablePeerAccess(all[q].dev, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) my_ok = 0;
          cudaGetLastError();
        }
        mg->peer_base[q] = (char*)(uintptr_t)all[q].ptr;
        mg->peer_ipc[q] = false;
      } else {
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[q].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          my_ok = 0;
          break;
        }
        mg->peer_base[q] = (char*)p;
        mg->peer_ipc[q] = true;
      }
    }
  } else {
    my_ok = 0;
  }
  // every rank must have every mapping, or nobody uses peer memory
  int* word = reinterpret_cast<int*>(xs + (size_t)(mg->nranks + 1) * 128);
  GFFM_CUDA(cudaMemcpy(word, &my_ok, sizeof(int), cudaMemcpyHostToDevice));
  MG_NCCL(nc->AllReduce(word, word, 1, ncclInt32, ncclMin, mg->comm, mg->s_comm));
  GFFM_CUDA(cudaStreamSynchronize(mg->s_comm));
  int all_ok = 0;
  GFFM_CUDA(cudaMemcpy(&all_ok, word, sizeof(int), cudaMemcpyDeviceToHost));
  int zero = 0;
  GFFM_CUDA(cudaMemcpy(word, &zero, sizeof(int), cudaMemcpyHostToDevice));
  mg->p2p_ok = all_ok == 1;
  if (!mg->p2p_ok) mg_close_peers(mg);
  // can a stream memory operation write a flag into PEER memory?  Every rank writes a token into a test word of every peer, all
  // ranks meet, every rank checks its own test words; one failure anywhere -> everybody signals with the one-warp kernel instead.
  mg->signal_memops = false;
  {
    const char* f = getenv("GFFM_MG_FLAGS");
    int mine_ok = (mg->p2p_ok && mg->wait_memops && mem_ops() && !(f && !strcmp(f, "kernel"))) ? 1 : 0;
    const uint32_t token = 0xC0FFEE00u + (uint32_t)(mg->arena_bytes >> 20);
    if (mine_ok) {
      for (int q = 0; q < mg->nranks; ++q) {
        if (q == mg->rank) continue;
        if (mem_ops()->write32((CUstream)mg->s_comm, (CUdeviceptr)(uintptr_t)(mg->ctl(q) + F_PROBE + mg->rank), token, CU_STREAM_WRITE_VALUE_DEFAULT) != CUDA_SUCCESS) mine_ok = 0;
      }
      if (cudaStreamSynchronize(mg->s_comm) != cudaSuccess) mine_ok = 0;
      cudaGetLastError();
    }
    GFFM_TRY(mg_barrier_impl(mg));
    if (mine_ok) {
      uint32_t got[MG_MAX_RANKS] = {};
      GFFM_CUDA(cudaMemcpy(got, mg->ctl(mg->rank) + F_PROBE, sizeof(uint32_t) * mg->nranks, cudaMemcpyDeviceToHost));
      for (int q = 0; q < mg->nranks; ++q)
        if (q != mg->rank && got[q] != token) mine_ok = 0;
    }
    GFFM_CUDA(cudaMemcpy(word, &mine_ok, sizeof(int), cudaMemcpyHostToDevice));
    MG_NCCL(nc->AllReduce(word, word, 1, ncclInt32, ncclMin, mg->comm, mg->s_comm));
    GFFM_CUDA(cudaStreamSynchronize(mg->s_comm));
    int all_sig = 0;
    GFFM_CUDA(cudaMemcpy(&all_sig, word, sizeof(int), cudaMemcpyDeviceToHost));
    GFFM_CUDA(cudaMemcpy(word, &zero, sizeof(int), cudaMemcpyHostToDevice));
    mg->signal_memops = all_sig == 1;
  }
  if ((mg->requested == GFFM_MG_P2P_PLANES || mg->requested == GFFM_MG_P2P_PUSH || mg->requested == GFFM_MG_P2P_RAW) && !mg->p2p_ok)
    GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "a peer-memory transport was requested but peer memory (CUDA IPC / peer access) is not available between all ranks");
  mg->transport = mg->requested != GFFM_MG_AUTO ? mg->requested : (mg->p2p_ok ? GFFM_MG_P2P_PUSH : GFFM_MG_NCCL_PLANES);
  mg->epoch = 0;
  // the fused split + push kernel owns `push_sms` SMs; the persistent GEMM gets the others (unless the caller set its own cap)
  if (mg->transport == GFFM_MG_P2P_PUSH && mg->nranks > 1) {
    if (mg->saved_gemm_ctas < 0) mg->saved_gemm_ctas = mg->ctx->gemm_ctas;
    if (mg->saved_gemm_ctas == 0) mg->ctx->gemm_ctas = mg->ctx->num_sms - mg->push_sms;
  } else if (mg->saved_gemm_ctas >= 0) {
    mg->ctx->gemm_ctas = mg->saved_gemm_ctas;
    mg->saved_gemm_ctas = -1;
  }
  return GFFM_OK;
}

int mg_sid(gffm_mg* mg, cudaStream_t st) {
  if (st == mg->s_dist) return 2;
  if (st == mg->s_comm) return 5;
  for (int j = 0; j < gffm_mg::NCOPY; ++j) {
    if (st == mg->s_pull[j]) return 3;
    if (st == mg->s_push[j]) return 4;
  }
  return 0;
}

int32_t mg_wait(gffm_mg* mg, cudaStream_t st, int word0, uint32_t mask, uint32_t target) {
  if (!mask) return GFFM_OK;
  uint32_t* ctl = mg->ctl(mg->rank);
  cudaEvent_t tw = gffm_trace_begin(mg->ctx, st);
  if (mg->wait_memops) {
    MemOps* mo = mem_ops();
    for (int i = 0; i < 32; ++i) {
      if (!((mask >> i) & 1u)) continue;
      const CUresult cr = mo->wait32((CUstream)st, (CUdeviceptr)(uintptr_t)(ctl + word0 + i), target, CU_STREAM_WAIT_VALUE_GEQ);
      if (cr != CUDA_SUCCESS) GFFM_FAIL(GFFM_ERR_CUDA, "cuStreamWaitValue32 failed with CUresult %d", (int)cr);
    }
  } else {
    mg_wait_kernel<<<1, 32, 0, st>>>(ctl + word0, mask, target, ctl + F_ERROR, mg->timeout_ns);
    GFFM_LAUNCH_CHECK(mg->ctx);
  }
  gffm_trace_end(mg->ctx, word0 == F_STAGED ? "wait:staged" : word0 == F_READY ? "wait:ready" : word0 == F_PULLED ? "wait:pulled" : word0 == F_FREE ? "wait:free" : "wait:splitdn", (int)target,
                 mg_sid(mg, st), tw, st);
  return GFFM_OK;
}

int32_t mg_signal(gffm_mg* mg, cudaStream_t st, const MgTargets& t, int n, uint32_t value) {
  if (n <= 0) return GFFM_OK;
  if (mg->signal_memops) {
    MemOps* mo = mem_ops();
    for (int i = 0; i < n; ++i) {
      if (!t.p[i]) continue;
      const CUresult cr = mo->write32((CUstream)st, (CUdeviceptr)(uintptr_t)t.p[i], value, CU_STREAM_WRITE_VALUE_DEFAULT);
      if (cr != CUDA_SUCCESS) GFFM_FAIL(GFFM_ERR_CUDA, "cuStreamWriteValue32 failed with CUresult %d", (int)cr);
    }
    return GFFM_OK;
  }
  mg_signal_kernel<<<1, 64, 0, st>>>(t, n, value);
  GFFM_LAUNCH_CHECK(mg->ctx);
  return GFFM_OK;
}

struct MgSet {
  GemmBPlan* plan = nullptr;
  int src = 0, src2 = -1;  // which source matrix (and optional addend) the planes are built from
  size_t off = 0;          // byte offset of this plane set inside the plane buffer
};

struct MgRound {
  int nsrc = 1;
  MatView src[2];  // root: the data; other ranks: a same-shape local matrix (receive buffer of the broadcast transport)
  int nsets = 1;
  MgSet sets[3];
  int64_t n = 0, kc = 0;
  int root = 0;
  cudaEvent_t b_ready = nullptr;
  bool fresh_data = true;  // false: a further K-chunk of the same sources (the broadcast transport does not resend them)
};

struct MgRoundOut {
  uint8_t* planes = nullptr;
  int64_t rowsPB = 0;
  int npanels = 0;
  int b = 0;
  int64_t off[MG_MAX_RANKS + 1];
  int order[MG_MAX_RANKS];
  cudaEvent_t ready[MG_MAX_RANKS];
};

// the 256-column blocks of B dealt to every rank but `root` (rank order, the first owners take the remainder); off[root + 1] == off[root]
void mg_root_free_ranges(int64_t n, int nr, int root, int64_t* off) {
  const int64_t blocks = ceil_div(n, 256), base = blocks / (nr - 1), extra = blocks % (nr - 1);
  off[0] = 0;
  for (int q = 0, o = 0; q < nr; ++q) {
    const int64_t nb = q == root ? 0 : base + (o++ < extra ? 1 : 0);
    off[q + 1] = std::min<int64_t>(n, off[q] + nb * 256);
  }
}

int32_t mg_distribute(gffm_mg* mg, MgRound& R, MgRoundOut* out) {
  gffm_ctx* ctx = mg->ctx;
  NcclApi* nc = nccl_api();
  const int nr = mg->nranks, r = mg->rank, root = R.root;
  const int64_t n = R.n, kc = R.kc;
  const int64_t per = round_up(ceil_div(n, nr), 256), rowsPB = per * nr;
  // With many ranks the root's NVLink egress is the bottleneck of the peer-memory transports when it owns a column range like
  // everybody else: it sends the other ranks' uint32 ranges (4 bytes per element) AND its own range's planes to every peer (8 x 1
  // byte per element and peer) -- 2.75 GB per 16384^2 product at 8 ranks against 1.9 GB for the others (measured: the root's split +
  // push kernel takes 4.8 ms, everybody else's 2.9 ms, and the 4 ms step waits for it; profiles/r02_notes.md).  From `root_free_min`
  // ranks on the root therefore owns NO range: the 256-column blocks of B are dealt to the other ranks, the root only scatters.
  const bool root_free_possible = root >= 0 && mg->root_free_min > 0 && nr >= mg->root_free_min && nr >= 2;
  const int64_t per_stage = root_free_possible ? round_up(ceil_div(n, nr - 1), 256) : per;  // upper bound of any owner's range
  // plane-set offsets and sizes
  size_t planes_bytes = 0;
  for (int s = 0; s < R.nsets; ++s) {
    const BPlaneSpec* sp = gffm_bplan_spec(R.sets[s].plan);
    R.sets[s].off = planes_bytes;
    planes_bytes += ((size_t)sp->nplanes * rowsPB * sp->Kp + 1023) & ~(size_t)1023;
  }
  const int64_t ld_c = round_up(kc, 32);                      // compact staging columns (peer-memory transport)
  const int64_t ld_b = R.src[0].ld;                           // NCCL scatter keeps the source's leading dimension
  // GFFM_MG_P2P_RAW stages ALL of B (compact columns) on every rank, the other transports only the own range
  const size_t stage_bytes = mg->requested == GFFM_MG_P2P_RAW ? (size_t)R.nsrc * rowsPB * ld_c * 4 : (size_t)R.nsrc * per_stage * std::max(ld_c, ld_b) * 4;
  GFFM_TRY(mg_ensure_arena(mg, stage_bytes, planes_bytes));
  const bool distributed = root < 0;  // every rank already holds its own column range of B: nothing to push / scatter
  const int transport = (distributed && mg->transport == GFFM_MG_NCCL_BCAST) ? GFFM_MG_NCCL_PLANES : mg->transport;
  const bool root_free = root_free_possible && (transport == GFFM_MG_P2P_PLANES || transport == GFFM_MG_P2P_PUSH || transport == GFFM_MG_P2P_RAW);
  if (!root_free) {
    for (int q = 0; q <= nr; ++q) out->off[q] = std::min<int64_t>(n, (int64_t)q * per);
  } else {
    mg_root_free_ranges(n, nr, root, out->off);
  }
  const uint32_t e = ++mg->epoch;
  const int b = (int)(e % (uint32_t)MG_NBUF);
  mg->rounds++;
  out->b = b;
  out->planes = (uint8_t*)mg->planes(b);
  out->rowsPB = rowsPB;
  out->npanels = nr;
  cudaEvent_t bready = R.b_ready;
  if (!bready) {  // "B is ready in the order of the context stream"
    GFFM_CUDA(cudaEventRecord(mg->ev_call, ctx->stream));
    bready = mg->ev_call;
  }
  bool push_planes = false;  // GFFM_MG_P2P_PUSH: the split stores to every rank's plane buffer
  const bool raw_stage = transport == GFFM_MG_P2P_RAW;  // staging holds all of B: source s, column c at ((s * rowsPB + c) * ld_c) words
  auto split_own = [&](int q, bool from_stage, int64_t ld_stage) -> int32_t {
    const int64_t c0 = out->off[q], cnt = out->off[q + 1] - c0;
    if (cnt <= 0) return GFFM_OK;
    cudaEvent_t ts = gffm_trace_begin(ctx, mg->s_dist);
    struct TraceGuard {
      gffm_ctx* c; cudaEvent_t a; cudaStream_t st; int q;
      ~TraceGuard() { gffm_trace_end(c, "splitB", q, 2, a, st); }
    } guard{ctx, ts, mg->s_dist, q};
    for (int s = 0; s < R.nsets; ++s) {
      const MgSet& S = R.sets[s];
      MatView v, v2;
      auto view_of_src = [&](int which) {
        if (from_stage && raw_stage) return MatView{reinterpret_cast<uint32_t*>(mg->stage(b)) + ((size_t)which * rowsPB + c0) * ld_stage, ld_stage, kc, cnt};
        if (from_stage) return MatView{reinterpret_cast<uint32_t*>(mg->stage(b)) + (size_t)which * per_stage * ld_stage, ld_stage, kc, cnt};
        return sub_view(R.src[which], 0, R.root < 0 ? 0 : c0, kc, cnt);  // distributed B: the local matrix IS the own range
      };
      v = view_of_src(S.src);
      if (S.src2 >= 0) v2 = view_of_src(S.src2);
      if (push_planes) {
        // destinations of the kernel: the local buffer, then the peers r+1, r+2, ... except the last `ce` of them (served by copy engines)
        uint8_t* bufs[MG_MAX_RANKS];
        int nb = 0;
        bufs[nb++] = (uint8_t*)(mg->base + mg->planes_off(b) + S.off);
        const int ce = (r == R.root) ? 0 : std::min(mg->push_ce_peers, nr - 1);  // the root's copy engines already carry the uint32 ranges
        for (int i = 1; i < nr - ce; ++i) bufs[nb++] = (uint8_t*)(mg->peer_base[(r + i) % nr] + mg->planes_off(b) + S.off);
        GFFM_TRY(gffm_bplan_split_push(ctx, S.plan, v, S.src2 >= 0 ? &v2 : nullptr, bufs, nb, rowsPB, c0, mg->s_dist, mg->push_sms));
      } else {
        GFFM_TRY(gffm_bplan_split(ctx, S.plan, v, S.src2 >= 0 ? &v2 : nullptr, out->planes + S.off, rowsPB, c0, mg->s_dist));
      }
    }
    return GFFM_OK;
  };

  push_planes = transport == GFFM_MG_P2P_PUSH;
  if (transport == GFFM_MG_P2P_RAW) {
    const int nc = mg->ncopy;
    const bool have_b = r == root || distributed;  // this rank reads its own range (root: every range) straight from its matrix
    auto stage_at = [&](int q, int s, int64_t c) { return mg->peer_base[q] + mg->stage_off(b) + ((size_t)s * rowsPB + c) * ld_c * 4; };
    bool used_push[gffm_mg::NCOPY] = {};
    // ---- root: every owner's uint32 range into that owner's staging buffer (copy engines, peer memory) -----------------------
    if (r == root) {
      for (int j = 0; j < nc; ++j) GFFM_CUDA(cudaStreamWaitEvent(mg->s_push[j], bready, 0));
      for (int i = 1; i < nr; ++i) {
        const int q = (root + i) % nr;
        const int64_t c0 = out->off[q], cnt = out->off[q + 1] - c0;
        const int64_t chunk = round_up(ceil_div(std::max<int64_t>(cnt, 1), nc), 8);
        for (int j = 0; j < nc; ++j) {
          const int64_t j0 = std::min<int64_t>(cnt, (int64_t)j * chunk), j1 = std::min<int64_t>(cnt, (int64_t)(j + 1) * chunk);
          if (j > 0 && j1 <= j0) continue;
          used_push[j] = true;
          GFFM_TRY(mg_wait(mg, mg->s_push[j], F_SPLIT_DONE, 1u << q, e - MG_NBUF));  // q is done with what this staging buffer held
          cudaEvent_t tp = gffm_trace_begin(ctx, mg->s_push[j]);
          for (int s = 0; s < R.nsrc && j1 > j0; ++s)
            GFFM_CUDA(cudaMemcpy2DAsync(stage_at(q, s, c0 + j0), (size_t)ld_c * 4, R.src[s].p + (c0 + j0) * R.src[s].ld, (size_t)R.src[s].ld * 4,
                                        (size_t)kc * 4, (size_t)(j1 - j0), cudaMemcpyDefault, mg->s_push[j]));
          gffm_trace_end(ctx, "push", q, 4, tp, mg->s_push[j]);
          if (j > 0) {
            GFFM_CUDA(cudaEventRecord(mg->copy_ev[1][j], mg->s_push[j]));
            GFFM_CUDA(cudaStreamWaitEvent(mg->s_push[0], mg->copy_ev[1][j], 0));
          }
        }
        MgTargets t;
        t.p[0] = mg->ctl(q) + F_STAGED;
        GFFM_TRY(mg_signal(mg, mg->s_push[0], t, 1, e));
      }
    }
    // ---- distribution stream: the own range is there (pushed by the root / part of the local matrix), the plane buffer is free -------
    if (!have_b) GFFM_TRY(mg_wait(mg, mg->s_dist, F_STAGED, 1u, e));
    else GFFM_CUDA(cudaStreamWaitEvent(mg->s_dist, bready, 0));
    GFFM_CUDA(cudaStreamWaitEvent(mg->s_dist, mg->gemm_done[b], 0));  // the local GEMMs of the epoch that used plane buffer b before are done
    // ---- owners: forward the own range to every rank that does not hold B (copy engines; posted writes over NVLink) ----------------
    {
      const int64_t c0 = out->off[r], cnt = out->off[r + 1] - c0;
      for (int i = 1; i < nr && cnt > 0; ++i) {  // an empty range has no flag: every rank knows the ranges
        const int p = (r + i) % nr;
        if (p == root) continue;  // the root splits from its own matrix
        cudaStream_t cs = mg->s_push[i % nc];
        used_push[i % nc] = true;
        {
          // the copy streams wait for the own range themselves (not through the distribution stream, which is still busy with the
          // previous product's splits): the forwards of product e+1 run while product e is being split
          if (!have_b) GFFM_TRY(mg_wait(mg, cs, F_STAGED, 1u, e));
          else GFFM_CUDA(cudaStreamWaitEvent(cs, bready, 0));
          GFFM_TRY(mg_wait(mg, cs, F_SPLIT_DONE, 1u << p, e - MG_NBUF));  // p is done with what its staging buffer b held
          cudaEvent_t tp = gffm_trace_begin(ctx, cs);
          for (int s = 0; s < R.nsrc; ++s) {
            if (have_b) {
              const uint32_t* src = R.src[s].p + (distributed ? 0 : c0) * R.src[s].ld;
              GFFM_CUDA(cudaMemcpy2DAsync(stage_at(p, s, c0), (size_t)ld_c * 4, src, (size_t)R.src[s].ld * 4, (size_t)kc * 4, (size_t)cnt, cudaMemcpyDefault, cs));
            } else {
              GFFM_CUDA(cudaMemcpyAsync(stage_at(p, s, c0), mg->stage(b) + ((size_t)s * rowsPB + c0) * ld_c * 4, (size_t)cnt * ld_c * 4, cudaMemcpyDefault, cs));
            }
          }
          gffm_trace_end(ctx, "fwd", p, 4, tp, cs);
        }
        MgTargets t;
        t.p[0] = mg->ctl(p) + F_READY + r;
        GFFM_TRY(mg_signal(mg, cs, t, 1, e));
      }
    }
    // the copies that READ this rank's matrix / staging buffer: joined on copy stream 0 (ev_push: the caller may modify B; the
    // distribution stream waits for it before it declares the staging buffer consumed)
    for (int j = 1; j < gffm_mg::NCOPY; ++j) {
      if (!used_push[j]) continue;
      GFFM_CUDA(cudaEventRecord(mg->copy_ev[1][j], mg->s_push[j]));
      GFFM_CUDA(cudaStreamWaitEvent(mg->s_push[0], mg->copy_ev[1][j], 0));
    }
    GFFM_CUDA(cudaEventRecord(mg->ev_push, mg->s_push[0]));
    // ---- every rank: split every range as it arrives, the own one first, then r-1, r-2, ...: owner q forwards to q+1, q+2, ... in
    // that order, so in "slot" i every rank receives the range of a different owner (r-i) and needs exactly that one next (a ring) ----
    for (int i = 0; i < nr; ++i) {
      const int q = (r + nr - i) % nr;
      const bool local_src = q == r ? have_b : r == root;  // read from the matrix itself instead of the staging buffer
      if (!local_src && q != r && out->off[q + 1] > out->off[q]) GFFM_TRY(mg_wait(mg, mg->s_dist, F_READY, 1u << q, e));
      GFFM_TRY(split_own(q, !local_src, ld_c));
      GFFM_CUDA(cudaEventRecord(mg->ready_ev[b][q], mg->s_dist));
      out->order[i] = q;
      out->ready[q] = mg->ready_ev[b][q];
    }
    // "staging buffer b of this rank is consumed": after the splits AND after the forwards that read it
    GFFM_CUDA(cudaStreamWaitEvent(mg->s_dist, mg->ev_push, 0));
    {
      MgTargets t;
      int k = 0;
      for (int q = 0; q < nr; ++q)
        if (q != r) t.p[k++] = mg->ctl(q) + F_SPLIT_DONE + r;
      GFFM_TRY(mg_signal(mg, mg->s_dist, t, k, e));
    }
    GFFM_CUDA(cudaEventRecord(mg->split_ev[b], mg->s_dist));
    return GFFM_OK;
  }

  if (transport == GFFM_MG_P2P_PLANES || transport == GFFM_MG_P2P_PUSH) {
    // ---- root: push every other rank's uint32 column range into its staging buffer (copy engines, peer memory) -------------
    const int nc = mg->ncopy;
    if (r == root) {
      for (int j = 0; j < nc; ++j) GFFM_CUDA(cudaStreamWaitEvent(mg->s_push[j], bready, 0));
      for (int i = 1; i < nr; ++i) {
        const int q = (root + i) % nr;
        const int64_t c0 = out->off[q], cnt = out->off[q + 1] - c0;
        // the columns of q's range are cut into `nc` chunks, one per copy stream; stream 0 joins them and raises q's flag
        const int64_t chunk = round_up(ceil_div(std::max<int64_t>(cnt, 1), nc), 8);
        for (int j = 0; j < nc; ++j) {
          const int64_t j0 = std::min<int64_t>(cnt, (int64_t)j * chunk), j1 = std::min<int64_t>(cnt, (int64_t)(j + 1) * chunk);
          if (j > 0 && j1 <= j0) continue;
          GFFM_TRY(mg_wait(mg, mg->s_push[j], F_SPLIT_DONE, 1u << q, e - MG_NBUF));  // q has consumed what this staging buffer held
          cudaEvent_t tp = gffm_trace_begin(ctx, mg->s_push[j]);
          for (int s = 0; s < R.nsrc && j1 > j0; ++s) {
            char* dst = mg->peer_base[q] + mg->stage_off(b) + ((size_t)s * per_stage + j0) * ld_c * 4;
            GFFM_CUDA(cudaMemcpy2DAsync(dst, (size_t)ld_c * 4, R.src[s].p + (c0 + j0) * R.src[s].ld, (size_t)R.src[s].ld * 4, (size_t)kc * 4,
                                        (size_t)(j1 - j0), cudaMemcpyDefault, mg->s_push[j]));
          }
          gffm_trace_end(ctx, "push", q, 4, tp, mg->s_push[j]);
          if (j > 0) {
            GFFM_CUDA(cudaEventRecord(mg->copy_ev[1][j], mg->s_push[j]));
            GFFM_CUDA(cudaStreamWaitEvent(mg->s_push[0], mg->copy_ev[1][j], 0));
          }
        }
        MgTargets t;
        t.p[0] = mg->ctl(q) + F_STAGED;
        GFFM_TRY(mg_signal(mg, mg->s_push[0], t, 1, e));
      }
      GFFM_CUDA(cudaEventRecord(mg->ev_push, mg->s_push[0]));
    }
    // ---- every rank: split the own range ------------------------------------------------------------------------------------
    if (r != root && !distributed) GFFM_TRY(mg_wait(mg, mg->s_dist, F_STAGED, 1u, e));
    else GFFM_CUDA(cudaStreamWaitEvent(mg->s_dist, bready, 0));
    GFFM_CUDA(cudaStreamWaitEvent(mg->s_dist, mg->gemm_done[b], 0));                       // the local GEMMs of epoch e-2 are done with this buffer
    const uint32_t others = (nr >= 32 ? 0xffffffffu : ((1u << nr) - 1u)) & ~(1u << r);
    if (push_planes) GFFM_TRY(mg_wait(mg, mg->s_dist, F_FREE, others, e - MG_NBUF));  // every peer's GEMMs of the epoch that used this buffer before are done with ITS copy
    else GFFM_TRY(mg_wait(mg, mg->s_dist, F_PULLED, others, e - MG_NBUF));           // ... and so are the peers' pulls from this buffer
    if (push_planes) GFFM_CUDA(cudaStreamWaitEvent(mg->s_dist, mg->ce_done[b], 0));  // copy engines of epoch e-2 no longer read this buffer
    GFFM_TRY(split_own(r, r != root && !distributed, ld_c));
    GFFM_CUDA(cudaEventRecord(mg->ready_ev[b][r], mg->s_dist));
    const int ce_peers = (push_planes && r != root) ? std::min(mg->push_ce_peers, nr - 1) : 0;
    if (ce_peers > 0) {
      // the last `ce_peers` peers get the planes from the copy engines (local plane buffer -> peer plane buffer), each followed by
      // that peer's flag, spread over the copy streams
      GFFM_CUDA(cudaEventRecord(mg->split_ev[b], mg->s_dist));
      const int64_t cnt = out->off[r + 1] - out->off[r];
      for (int i = nr - ce_peers; i < nr; ++i) {
        const int q = (r + i) % nr;
        cudaStream_t cs = mg->s_push[1 + (i % (gffm_mg::NCOPY - 1))];
        GFFM_CUDA(cudaStreamWaitEvent(cs, mg->split_ev[b], 0));
        cudaEvent_t tp = gffm_trace_begin(ctx, cs);
        for (int s2 = 0; s2 < R.nsets && cnt > 0; ++s2) {
          const BPlaneSpec* sp = gffm_bplan_spec(R.sets[s2].plan);
          const size_t pitch = (size_t)rowsPB * sp->Kp, rel = R.sets[s2].off + (size_t)out->off[r] * sp->Kp;
          for (int t = 0; t < sp->nplanes; ++t)
            GFFM_CUDA(cudaMemcpyAsync(mg->peer_base[q] + mg->planes_off(b) + rel + t * pitch, mg->planes(b) + rel + t * pitch, (size_t)cnt * sp->Kp,
                                      cudaMemcpyDefault, cs));
        }
        gffm_trace_end(ctx, "cepush", q, 4, tp, cs);
        MgTargets t1;
        t1.p[0] = mg->ctl(q) + F_READY + r;
        GFFM_TRY(mg_signal(mg, cs, t1, 1, e));
      }
      // join: the next user of this plane buffer waits for all of them
      for (int j = 1; j < gffm_mg::NCOPY; ++j) {
        GFFM_CUDA(cudaEventRecord(mg->copy_ev[1][j], mg->s_push[j]));
        GFFM_CUDA(cudaStreamWaitEvent(mg->s_pull[2], mg->copy_ev[1][j], 0));
      }
      GFFM_CUDA(cudaEventRecord(mg->ce_done[b], mg->s_pull[2]));
    }
    {
      MgTargets t;
      int k = 0;
      for (int i = 1; i < nr - ce_peers; ++i) t.p[k++] = mg->ctl((r + i) % nr) + F_READY + r;
      GFFM_TRY(mg_signal(mg, mg->s_dist, t, k, e));
      // "my staging buffer of this parity is free again": told to EVERY rank, because any of them may be the root of a later product
      k = 0;
      for (int q = 0; q < nr; ++q)
        if (q != r) t.p[k++] = mg->ctl(q) + F_SPLIT_DONE + r;
      GFFM_TRY(mg_signal(mg, mg->s_dist, t, k, e));
    }
    if (push_planes) {
      // ---- consumers: the planes arrive by themselves; the arrival stream turns the owners' flags into events for the GEMMs ----
      for (int i = 1; i < nr; ++i) {
        const int q = (r + i) % nr;
        GFFM_TRY(mg_wait(mg, mg->s_pull[0], F_READY, 1u << q, e));
        GFFM_CUDA(cudaEventRecord(mg->ready_ev[b][q], mg->s_pull[0]));
      }
      for (int i = 0; i < nr; ++i) {
        out->order[i] = (r + i) % nr;
        out->ready[i] = mg->ready_ev[b][i];
      }
      return GFFM_OK;
    }
    // ---- every rank: pull the other ranges' planes from their owners ----------------------------------------------------------
    for (int j = 0; j < nc; ++j) GFFM_CUDA(cudaStreamWaitEvent(mg->s_pull[j], mg->gemm_done[b], 0));
    for (int i = 1; i < nr; ++i) {
      const int q = (r + i) % nr;
      const int64_t cnt = out->off[q + 1] - out->off[q];
      // piece (set, plane, row chunk) of q's range goes to copy stream (running piece index) % nc; stream 0 joins them.  With fewer
      // planes than copy streams every plane is cut into row chunks so that all streams (copy engines) carry a share.
      int pl = 0, total_planes = 0;
      for (int s = 0; s < R.nsets; ++s) total_planes += gffm_bplan_spec(R.sets[s].plan)->nplanes;
      const int parts = total_planes >= nc ? 1 : nc / std::max(total_planes, 1);
      bool used[gffm_mg::NCOPY] = {};
      cudaEvent_t tpl[gffm_mg::NCOPY] = {};
      for (int s = 0; s < R.nsets && cnt > 0; ++s) {
        const BPlaneSpec* sp = gffm_bplan_spec(R.sets[s].plan);
        const size_t pitch = (size_t)rowsPB * sp->Kp;
        const int64_t rows_part = ceil_div(cnt, parts);
        for (int t = 0; t < sp->nplanes; ++t) {
          for (int64_t i0 = 0; i0 < cnt; i0 += rows_part, ++pl) {
            const int64_t ni = std::min(rows_part, cnt - i0);
            const int j = pl % nc;
            if (!used[j]) {
              GFFM_TRY(mg_wait(mg, mg->s_pull[j], F_READY, 1u << q, e));
              tpl[j] = gffm_trace_begin(ctx, mg->s_pull[j]);
              used[j] = true;
            }
            const size_t rel = R.sets[s].off + (size_t)t * pitch + ((size_t)out->off[q] + i0) * sp->Kp;
            GFFM_CUDA(cudaMemcpyAsync(mg->planes(b) + rel, mg->peer_base[q] + mg->planes_off(b) + rel, (size_t)ni * sp->Kp, cudaMemcpyDefault, mg->s_pull[j]));
          }
        }
      }
      if (!used[0]) GFFM_TRY(mg_wait(mg, mg->s_pull[0], F_READY, 1u << q, e));  // empty range: still part of the protocol
      for (int j = 0; j < nc; ++j) {
        if (!used[j]) continue;
        gffm_trace_end(ctx, "pull", q, 3, tpl[j], mg->s_pull[j]);
        if (j > 0) {
          GFFM_CUDA(cudaEventRecord(mg->copy_ev[0][j], mg->s_pull[j]));
          GFFM_CUDA(cudaStreamWaitEvent(mg->s_pull[0], mg->copy_ev[0][j], 0));
        }
      }
      GFFM_CUDA(cudaEventRecord(mg->ready_ev[b][q], mg->s_pull[0]));
      MgTargets t;
      t.p[0] = mg->ctl(q) + F_PULLED + r;
      GFFM_TRY(mg_signal(mg, mg->s_pull[0], t, 1, e));
    }
    for (int i = 0; i < nr; ++i) {
      out->order[i] = (r + i) % nr;
      out->ready[i] = mg->ready_ev[b][i];
    }
    return GFFM_OK;
  }

  if (transport == GFFM_MG_NCCL_PLANES) {
    // ---- scatter of the uint32 ranges (grouped send / recv), split of the own range, grouped in-place all-gather of the planes ----
    const int64_t cnt_r = out->off[r + 1] - out->off[r];
    GFFM_CUDA(cudaStreamWaitEvent(mg->s_comm, bready, 0));
    GFFM_CUDA(cudaStreamWaitEvent(mg->s_comm, mg->stage_free[b], 0));
    cudaEvent_t tsc = gffm_trace_begin(ctx, mg->s_comm);
    if (nr > 1 && !distributed) {
      MG_NCCL(nc->GroupStart());
      if (r == root) {
        for (int q = 0; q < nr; ++q) {
          const int64_t c0 = out->off[q], cnt = out->off[q + 1] - c0;
          if (q == root || cnt <= 0) continue;
          for (int s = 0; s < R.nsrc; ++s)
            MG_NCCL(nc->Send(R.src[s].p + c0 * R.src[s].ld, (size_t)((cnt - 1) * R.src[s].ld + kc), ncclUint32, q, mg->comm, mg->s_comm));
        }
      } else if (cnt_r > 0) {
        for (int s = 0; s < R.nsrc; ++s)
          MG_NCCL(nc->Recv(reinterpret_cast<uint32_t*>(mg->stage(b)) + (size_t)s * per_stage * ld_b, (size_t)((cnt_r - 1) * ld_b + kc), ncclUint32, root, mg->comm,
                           mg->s_comm));
      }
      MG_NCCL(nc->GroupEnd());
    }
    gffm_trace_end(ctx, "scatter", 0, 5, tsc, mg->s_comm);
    GFFM_CUDA(cudaEventRecord(mg->staged[b], mg->s_comm));
    GFFM_CUDA(cudaStreamWaitEvent(mg->s_dist, mg->staged[b], 0));
    GFFM_CUDA(cudaStreamWaitEvent(mg->s_dist, mg->gemm_done[b], 0));
    GFFM_TRY(split_own(r, r != root && !distributed, ld_b));
    GFFM_CUDA(cudaEventRecord(mg->ready_ev[b][r], mg->s_dist));
    GFFM_CUDA(cudaEventRecord(mg->stage_free[b], mg->s_dist));
    GFFM_CUDA(cudaStreamWaitEvent(mg->s_comm, mg->ready_ev[b][r], 0));
    cudaEvent_t tag = gffm_trace_begin(ctx, mg->s_comm);
    if (nr > 1) {
      MG_NCCL(nc->GroupStart());
      for (int s = 0; s < R.nsets; ++s) {
        const BPlaneSpec* sp = gffm_bplan_spec(R.sets[s].plan);
        for (int t = 0; t < sp->nplanes; ++t) {
          char* plane = mg->planes(b) + R.sets[s].off + (size_t)t * rowsPB * sp->Kp;
          MG_NCCL(nc->AllGather(plane + (size_t)r * per * sp->Kp, plane, (size_t)per * sp->Kp, ncclUint8, mg->comm, mg->s_comm));
        }
      }
      MG_NCCL(nc->GroupEnd());
    }
    gffm_trace_end(ctx, "allgather", 0, 5, tag, mg->s_comm);
    GFFM_CUDA(cudaEventRecord(mg->gathered[b], mg->s_comm));
    GFFM_CUDA(cudaEventRecord(mg->ev_comm, mg->s_comm));
    for (int i = 0; i < nr; ++i) {
      out->order[i] = (r + i) % nr;
      out->ready[i] = i == r ? mg->ready_ev[b][r] : mg->gathered[b];
    }
    return GFFM_OK;
  }

  // ---- GFFM_MG_NCCL_BCAST: broadcast of the uint32 ranges into every rank's copy of B, every rank splits everything ---------------
  GFFM_CUDA(cudaStreamWaitEvent(mg->s_comm, bready, 0));
  GFFM_CUDA(cudaStreamWaitEvent(mg->s_dist, bready, 0));
  GFFM_CUDA(cudaStreamWaitEvent(mg->s_dist, mg->gemm_done[b], 0));
  for (int q = 0; q < nr; ++q) {
    const int64_t c0 = out->off[q], cnt = out->off[q + 1] - c0;
    if (cnt > 0 && R.fresh_data && nr > 1) {
      GFFM_CUDA(cudaStreamWaitEvent(mg->s_comm, mg->bc_consumed[q], 0));  // the previous product has turned this range into planes
      cudaEvent_t tb = gffm_trace_begin(ctx, mg->s_comm);
      for (int s = 0; s < R.nsrc; ++s) {
        uint32_t* buf = R.src[s].p + c0 * R.src[s].ld;
        MG_NCCL(nc->Broadcast(buf, buf, (size_t)((cnt - 1) * R.src[s].ld + kc), ncclUint32, root, mg->comm, mg->s_comm));
      }
      gffm_trace_end(ctx, "bcast", q, 5, tb, mg->s_comm);
      GFFM_CUDA(cudaEventRecord(mg->bc_ev[q], mg->s_comm));
      GFFM_CUDA(cudaStreamWaitEvent(mg->s_dist, mg->bc_ev[q], 0));
    }
    GFFM_TRY(split_own(q, false, 0));
    GFFM_CUDA(cudaEventRecord(mg->ready_ev[b][q], mg->s_dist));
    GFFM_CUDA(cudaEventRecord(mg->bc_consumed[q], mg->s_dist));
    out->order[q] = q;
    out->ready[q] = mg->ready_ev[b][q];
  }
  GFFM_CUDA(cudaEventRecord(mg->ev_comm, mg->s_comm));
  return GFFM_OK;
}

// after the GEMMs of a round have been enqueued on the context stream
int32_t mg_round_done(gffm_mg* mg, const MgRound& R, const MgRoundOut& out) {
  gffm_ctx* ctx = mg->ctx;
  GFFM_CUDA(cudaEventRecord(mg->gemm_done[out.b], ctx->stream));
  // the caller may modify B (root) / reuse its receive buffer once the context stream has passed this point
  if (mg->transport == GFFM_MG_P2P_PUSH) {
    // tell every owner that this rank's copy of plane buffer b is free again (after the GEMMs just enqueued): side stream, so the
    // context stream carries nothing but compute
    GFFM_CUDA(cudaStreamWaitEvent(mg->s_pull[1], mg->gemm_done[out.b], 0));
    MgTargets t;
    int k = 0;
    for (int q = 0; q < mg->nranks; ++q)
      if (q != mg->rank) t.p[k++] = mg->ctl(q) + F_FREE + mg->rank;
    GFFM_TRY(mg_signal(mg, mg->s_pull[1], t, k, mg->epoch));
  }
  if (mg->transport == GFFM_MG_P2P_RAW) {
    // copies that read this rank's matrix (root: scatter + forward of its own range; distributed B: forward) and EVERY split that read
    // it or the staging buffer (the last one on the distribution stream) precede whatever the caller enqueues next
    if (mg->rank == R.root || R.root < 0) {
      GFFM_CUDA(cudaStreamWaitEvent(ctx->stream, mg->ev_push, 0));
      GFFM_CUDA(cudaStreamWaitEvent(ctx->stream, mg->split_ev[out.b], 0));
    }
  } else if (mg->transport == GFFM_MG_P2P_PLANES || mg->transport == GFFM_MG_P2P_PUSH) {
    if (mg->rank == R.root) GFFM_CUDA(cudaStreamWaitEvent(ctx->stream, mg->ev_push, 0));
  } else {
    GFFM_CUDA(cudaStreamWaitEvent(ctx->stream, mg->ev_comm, 0));
  }
  // the own range is split on the distribution stream; with a distributed B nothing else orders it before the caller's next write
  GFFM_CUDA(cudaStreamWaitEvent(ctx->stream, mg->ready_ev[out.b][mg->rank], 0));
  return GFFM_OK;
}

int32_t mg_check_common(gffm_mg* mg, int32_t root) {
  if (!mg) GFFM_FAIL(GFFM_ERR_INVALID, "null multi-GPU handle");
  if (root < GFFM_MG_DISTRIBUTED || root >= mg->nranks) GFFM_FAIL(GFFM_ERR_INVALID, "root %d out of range (%d ranks)", root, mg->nranks);
  return GFFM_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" int32_t gffm_mg_unique_id(void* id128) {
  if (!id128) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  NcclApi* nc = nccl_api();
  if (!nc) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded (set GFFM_NCCL_LIB): %s", dlerror() ? dlerror() : "symbol missing");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  MG_NCCL(nc->GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return GFFM_OK;
}

extern "C" int32_t gffm_mg_owner_ranges(int64_t n, int32_t nranks, int64_t* off) {
  if (!off || nranks < 1 || nranks > MG_MAX_RANKS || n < 0) GFFM_FAIL(GFFM_ERR_INVALID, "bad arguments");
  const int64_t per = round_up(ceil_div(n, nranks), 256);
  for (int q = 0; q <= nranks; ++q) off[q] = std::min<int64_t>(n, (int64_t)q * per);
  return GFFM_OK;
}

extern "C" int32_t gffm_mg_owner_ranges_root_free(int64_t n, int32_t nranks, int32_t root, int64_t* off) {
  if (!off || nranks < 2 || nranks > MG_MAX_RANKS || n < 0 || root < 0 || root >= nranks) GFFM_FAIL(GFFM_ERR_INVALID, "bad arguments");
  mg_root_free_ranges(n, nranks, root, off);
  return GFFM_OK;
}

extern "C" int32_t gffm_mg_create(gffm_ctx* ctx, const void* id128, int32_t nranks, int32_t rank, gffm_mg** out) {
  GFFM_ENTER_CTX(ctx);
  if (!ctx || !id128 || !out) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (nranks < 1 || nranks > MG_MAX_RANKS || rank < 0 || rank >= nranks) GFFM_FAIL(GFFM_ERR_INVALID, "rank %d of %d (at most %d ranks)", rank, nranks, MG_MAX_RANKS);
  NcclApi* nc = nccl_api();
  if (!nc) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded (set GFFM_NCCL_LIB)");
  gffm_mg* mg = new gffm_mg();
  mg->ctx = ctx;
  mg->rank = rank;
  mg->nranks = nranks;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  {
    ncclResult_t rr = nc->CommInitRank(&mg->comm, nranks, id, rank);
    if (rr != ncclSuccess) {
      gffm_set_error("ncclCommInitRank failed: %s", nc->GetErrorString(rr));
      delete mg;
      return GFFM_ERR_CUDA;
    }
  }
  GFFM_CUDA(cudaStreamCreateWithFlags(&mg->s_comm, cudaStreamNonBlocking));
  GFFM_CUDA(cudaStreamCreateWithFlags(&mg->s_dist, cudaStreamNonBlocking));
  for (int j = 0; j < gffm_mg::NCOPY; ++j) {
    GFFM_CUDA(cudaStreamCreateWithFlags(&mg->s_pull[j], cudaStreamNonBlocking));
    GFFM_CUDA(cudaStreamCreateWithFlags(&mg->s_push[j], cudaStreamNonBlocking));
  }
  if (const char* t = getenv("GFFM_MG_COPY_STREAMS")) mg->ncopy = std::max(1, std::min((int)gffm_mg::NCOPY, atoi(t)));
  if (!ctx->s_aux) GFFM_CUDA(cudaStreamCreateWithFlags(&ctx->s_aux, cudaStreamNonBlocking));
  auto mk = [](cudaEvent_t* e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming); };
  for (int b = 0; b < MG_NBUF; ++b) {
    GFFM_CUDA(mk(&mg->gemm_done[b]));
    GFFM_CUDA(mk(&mg->stage_free[b]));
    GFFM_CUDA(mk(&mg->staged[b]));
    GFFM_CUDA(mk(&mg->gathered[b]));
    for (int q = 0; q < nranks; ++q) GFFM_CUDA(mk(&mg->ready_ev[b][q]));
  }
  for (int q = 0; q < nranks; ++q) {
    GFFM_CUDA(mk(&mg->bc_ev[q]));
    GFFM_CUDA(mk(&mg->bc_consumed[q]));
  }
  for (int k = 0; k < 2; ++k)
    for (int j = 0; j < gffm_mg::NCOPY; ++j) GFFM_CUDA(mk(&mg->copy_ev[k][j]));
  {
    const char* f = getenv("GFFM_MG_FLAGS");  // "kernel": one-warp polling / signalling kernels instead of stream memory operations
    const bool want_memops = !(f && !strcmp(f, "kernel"));
    mg->wait_memops = want_memops && mem_ops() != nullptr;
    mg->signal_memops = false;  // decided per arena: the write must reach PEER memory (self-test in mg_ensure_arena)
  }
  for (int b = 0; b < MG_NBUF; ++b) {
    GFFM_CUDA(mk(&mg->split_ev[b]));
    GFFM_CUDA(mk(&mg->ce_done[b]));
  }
  if (const char* t = getenv("GFFM_MG_PUSH_SMS")) mg->push_sms = std::max(1, std::min(ctx->num_sms / 2, atoi(t)));
  if (const char* t = getenv("GFFM_MG_PUSH_CE_PEERS")) mg->push_ce_peers = std::max(0, atoi(t));
  if (const char* t = getenv("GFFM_MG_ROOT_FREE_MIN")) mg->root_free_min = std::max(0, atoi(t));
  GFFM_CUDA(mk(&mg->ev_call));
  GFFM_CUDA(mk(&mg->ev_push));
  GFFM_CUDA(mk(&mg->ev_comm));
  GFFM_CUDA(cudaMalloc(&mg->xchg_dev, (size_t)(nranks + 1) * 128 + 256));
  GFFM_CUDA(cudaMemset(mg->xchg_dev, 0, (size_t)(nranks + 1) * 128 + 256));
  if (const char* t = getenv("GFFM_MG_TRANSPORT")) mg->requested = atoi(t);
  if (const char* t = getenv("GFFM_MG_TIMEOUT_MS")) mg->timeout_ns = (unsigned long long)atoll(t) * 1000000ull;
  mg->transport = mg->requested;
  *out = mg;
  return GFFM_OK;
}

extern "C" int32_t gffm_mg_destroy(gffm_mg* mg) {
  if (!mg) return GFFM_OK;
  GFFM_ENTER_CTX(mg->ctx);
  mg_sync_streams(mg);
  if (mg->comm && mg->nranks > 1 && mg->base) {
    mg_barrier_impl(mg);  // peers may still be pulling from this arena
    mg_close_peers(mg);
    mg_barrier_impl(mg);
  } else {
    mg_close_peers(mg);
  }
  if (mg->base) cudaFree(mg->base);
  if (mg->xchg_dev) cudaFree(mg->xchg_dev);
  NcclApi* nc = nccl_api();
  if (nc && mg->comm) nc->CommDestroy(mg->comm);
  for (int b = 0; b < MG_NBUF; ++b) {
    for (cudaEvent_t e : {mg->gemm_done[b], mg->stage_free[b], mg->staged[b], mg->gathered[b]})
      if (e) cudaEventDestroy(e);
    for (int q = 0; q < MG_MAX_RANKS; ++q)
      if (mg->ready_ev[b][q]) cudaEventDestroy(mg->ready_ev[b][q]);
  }
  for (int q = 0; q < MG_MAX_RANKS; ++q) {
    if (mg->bc_ev[q]) cudaEventDestroy(mg->bc_ev[q]);
    if (mg->bc_consumed[q]) cudaEventDestroy(mg->bc_consumed[q]);
  }
  for (cudaEvent_t e : {mg->ev_call, mg->ev_push, mg->ev_comm})
    if (e) cudaEventDestroy(e);
  for (int b = 0; b < MG_NBUF; ++b)
    for (cudaEvent_t e : {mg->split_ev[b], mg->ce_done[b]})
      if (e) cudaEventDestroy(e);
  if (mg->saved_gemm_ctas >= 0) mg->ctx->gemm_ctas = mg->saved_gemm_ctas;
  for (cudaStream_t s : {mg->s_comm, mg->s_dist})
    if (s) cudaStreamDestroy(s);
  for (int j = 0; j < gffm_mg::NCOPY; ++j) {
    if (mg->s_pull[j]) cudaStreamDestroy(mg->s_pull[j]);
    if (mg->s_push[j]) cudaStreamDestroy(mg->s_push[j]);
    for (int k = 0; k < 2; ++k)
      if (mg->copy_ev[k][j]) cudaEventDestroy(mg->copy_ev[k][j]);
  }
  cudaGetLastError();
  delete mg;
  return GFFM_OK;
}

extern "C" int32_t gffm_mg_info(gffm_mg* mg, int32_t* rank, int32_t* nranks, int32_t* transport, int32_t* peer_memory) {
  if (!mg) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  if (rank) *rank = mg->rank;
  if (nranks) *nranks = mg->nranks;
  if (transport) *transport = mg->transport;
  if (peer_memory) *peer_memory = (mg->p2p_ok ? 1 : 0) | (mg->wait_memops ? 2 : 0) | (mg->signal_memops ? 4 : 0);
  return GFFM_OK;
}

extern "C" int32_t gffm_mg_set_transport(gffm_mg* mg, int32_t transport) {
  if (!mg) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  GFFM_ENTER_CTX(mg->ctx);
  if (transport < GFFM_MG_AUTO || transport > GFFM_MG_P2P_RAW) GFFM_FAIL(GFFM_ERR_INVALID, "unknown transport %d", transport);
  if (transport == mg->requested) return GFFM_OK;  // every rank passes the same value, so every rank returns here or nobody does
  // collective: drain everything, then let the next product rebuild the arena (and the peer mappings) under the new setting
  GFFM_TRY(mg_sync_streams(mg));
  if (mg->nranks > 1) GFFM_TRY(mg_barrier_impl(mg));
  mg_close_peers(mg);
  if (mg->nranks > 1) GFFM_TRY(mg_barrier_impl(mg));
  if (mg->base) {
    GFFM_CUDA(cudaFree(mg->base));
    mg->base = nullptr;
  }
  mg->stage_bytes = mg->planes_bytes = 0;
  mg->requested = transport;
  mg->transport = transport;
  return GFFM_OK;
}

extern "C" int32_t gffm_mg_barrier(gffm_mg* mg) {
  if (!mg) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  GFFM_ENTER_CTX(mg->ctx);
  GFFM_TRY(mg_sync_streams(mg));
  gffm_trace_dump(mg->ctx, mg->rank);
  if (mg->nranks > 1) GFFM_TRY(mg_barrier_impl(mg));
  if (mg->base) {  // a polling kernel that gave up leaves its mark in the control words
    uint32_t err = 0;
    GFFM_CUDA(cudaMemcpy(&err, reinterpret_cast<uint32_t*>(mg->base) + F_ERROR, 4, cudaMemcpyDeviceToHost));
    if (err) {
      uint32_t zero = 0;
      cudaMemcpy(reinterpret_cast<uint32_t*>(mg->base) + F_ERROR, &zero, 4, cudaMemcpyHostToDevice);
      GFFM_FAIL(GFFM_ERR_CUDA, "multi-GPU layer: rank %d timed out waiting for a peer flag (lane %u); results of the affected products are invalid", mg->rank,
                err - 1);
    }
  }
  return GFFM_OK;
}

// C_shard = A_shard * B mod P  (mul!(C,A,B), CuModMatrix.jl:767-787, on row blocks; B is distributed from `root`)
extern "C" int32_t gffm_mg_gemm(gffm_mg* mg, gffm_mat* C, gffm_mat* A, gffm_mat* B, int32_t root, void* b_ready_event, uint64_t R, uint64_t P) {
  GFFM_TRY(mg_check_common(mg, root));
  GFFM_ENTER_CTX(mg->ctx);
  if (!C || !A || !B) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  GFFM_NARROW_ONLY(C, A, B);
  if (C == A || C == B) GFFM_FAIL(GFFM_ERR_INVALID, "mg_gemm: C must not alias an operand");
  if (C->ctx != mg->ctx || A->ctx != mg->ctx || B->ctx != mg->ctx) GFFM_FAIL(GFFM_ERR_INVALID, "mg_gemm: matrices belong to another context");
  if (!P && (A->N != B->N || A->N != C->N)) GFFM_FAIL(GFFM_ERR_MODULUS_MISMATCH, "gemm operands have different moduli");
  int64_t own_cols = B->cols;
  if (root == GFFM_MG_DISTRIBUTED) {  // B holds this rank's own column range of the k x cols(C) matrix
    const int64_t per = round_up(ceil_div(C->cols, mg->nranks), 256);
    own_cols = std::min<int64_t>(C->cols, (int64_t)(mg->rank + 1) * per) - std::min<int64_t>(C->cols, (int64_t)mg->rank * per);
  }
  if (A->cols != B->rows || C->rows != A->rows || (root != GFFM_MG_DISTRIBUTED && C->cols != B->cols) || B->cols != own_cols)
    GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "mg_gemm: C %lldx%lld = A %lldx%lld * B %lldx%lld%s", (long long)C->rows, (long long)C->cols, (long long)A->rows,
              (long long)A->cols, (long long)B->rows, (long long)B->cols, root == GFFM_MG_DISTRIBUTED ? " (B = this rank's owner range)" : "");
  if (!P) P = C->N;
  if (!R) R = A->N > B->N ? A->N : B->N;
  if (P >= (1ull << 32) || R >= (1ull << 32)) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "mg_gemm needs R, P < 2^32");
  if (!gffm_tc_available(mg->ctx)) GFFM_FAIL(GFFM_ERR_CUDA, "tensor-map encoder unavailable");
  gffm_ctx* ctx = mg->ctx;
  const int64_t m = A->rows, k = A->cols, n = C->cols;
  gffm_touch(C);
  if (mg->rank != root && root >= 0) gffm_touch(B);  // receive buffer of the broadcast transport
  if (n == 0) return GFFM_OK;
  if (k == 0) return m > 0 ? gffm_fill_view(ctx, view_of(C), 0) : GFFM_OK;
  if (ctx->profile) gffm_profile_reset_tiles(ctx);
  const bool rns = R > 65536;
  const int64_t kmax = gffm_gemm_kchunk(R, rns);
  for (int64_t k0 = 0; k0 < k; k0 += kmax) {
    const int64_t kc = std::min(kmax, k - k0);
    GemmBPlan* plan = nullptr;
    GFFM_TRY(gffm_bplan_create(kc, R, P, rns && (R % P) == 0, k0 == 0 ? GFFM_GEMM_STORE : GFFM_GEMM_ADD, 0, &plan));
    MgRound round;
    round.nsrc = 1;
    round.src[0] = sub_view(view_of(B), k0, 0, kc, B->cols);
    round.nsets = 1;
    round.sets[0].plan = plan;
    round.n = n;
    round.kc = kc;
    round.root = root;
    round.b_ready = (cudaEvent_t)b_ready_event;
    round.fresh_data = k0 == 0;
    MgRoundOut out;
    int32_t st = mg_distribute(mg, round, &out);
    if (st == GFFM_OK) {
      ExtBPlanes ext;
      ext.planes = out.planes + round.sets[0].off;
      ext.rowsPB = out.rowsPB;
      ext.npanels = out.npanels;
      ext.off = out.off;
      ext.order = out.order;
      ext.ready = out.ready;
      if (m > 0) st = gffm_bplan_gemm(ctx, plan, view_of(C), sub_view(cached_view_of(A), 0, k0, m, kc), nullptr, n, ext, nullptr, 0);
      if (st == GFFM_OK) st = mg_round_done(mg, round, out);
    }
    gffm_bplan_destroy(plan);
    GFFM_TRY(st);
  }
  return GFFM_OK;
}

// KMatMul!(C,A,B) on row blocks (KaratsubaMatrix.jl:133-204): (C1 + N1*C2) = (A1 + N1*A2) * (B1 + N1*B2) mod N1*N2.  The planes of
// B1 (for P1), B1 + B2 (for P2, the limb add fused into the owner's split) and B2 (for P3) travel together in one round.
extern "C" int32_t gffm_mg_kmat_mul(gffm_mg* mg, gffm_mat* C1, gffm_mat* C2, gffm_mat* A1, gffm_mat* A2, gffm_mat* B1, gffm_mat* B2, uint64_t N1,
                                    uint64_t N2, int32_t root, void* b_ready_event) {
  GFFM_TRY(mg_check_common(mg, root));
  GFFM_ENTER_CTX(mg->ctx);
  if (!C1 || !C2 || !A1 || !A2 || !B1 || !B2) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  GFFM_NARROW_ONLY(C1, C2, A1, A2, B1, B2);
  if (N1 == 0 || N2 == 0 || N1 % N2 != 0) GFFM_FAIL(GFFM_ERR_INVALID, "Karatsuba product requires N2 | N1");
  if (N1 > (1ull << 26) || N1 * N2 > (1ull << 52)) GFFM_FAIL(GFFM_ERR_MODULUS_TOO_LARGE, "N1 <= 2^26 and N1*N2 <= 2^52 required");
  const int64_t m = A1->rows, k = A1->cols, n = B1->cols;
  if (A2->rows != m || A2->cols != k || B1->rows != k || B2->rows != k || B2->cols != n || C1->rows != m || C1->cols != n || C2->rows != m ||
      C2->cols != n)
    GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "mg KMatMul!: inconsistent sizes");
  if (B1->ld != B2->ld) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "limb pairs must share a leading dimension");
  if (root < 0) GFFM_FAIL(GFFM_ERR_UNSUPPORTED, "mg KMatMul!: B must live on a root rank");
  for (gffm_mat* c : {C1, C2})
    for (gffm_mat* x : {A1, A2, B1, B2})
      if (c == x) GFFM_FAIL(GFFM_ERR_INVALID, "mg KMatMul!: C must not alias an operand");
  gffm_ctx* ctx = mg->ctx;
  gffm_touch(C1);
  gffm_touch(C2);
  if (mg->rank != root) {
    gffm_touch(B1);
    gffm_touch(B2);
  }
  if (n == 0) return GFFM_OK;
  if (k == 0) {
    if (m > 0) {
      GFFM_TRY(gffm_fill_view(ctx, view_of(C1), 0));
      GFFM_TRY(gffm_fill_view(ctx, view_of(C2), 0));
    }
    return GFFM_OK;
  }
  if (ctx->profile) gffm_profile_reset_tiles(ctx);
  GemmBPlan* plans[3] = {nullptr, nullptr, nullptr};
  int32_t st = gffm_bplan_create(k, N1, N1 * N2, false, GFFM_GEMM_STORE, N1, &plans[0]);
  if (st == GFFM_OK) st = gffm_bplan_create(k, N1 + N2, N2, false, GFFM_GEMM_STORE, 0, &plans[1]);
  if (st == GFFM_OK) st = gffm_bplan_create(k, N2, N2, true, GFFM_GEMM_STORE, 0, &plans[2]);
  MgRound round;
  MgRoundOut out;
  if (st == GFFM_OK) {
    round.nsrc = 2;
    round.src[0] = view_of(B1);
    round.src[1] = view_of(B2);
    round.nsets = 3;
    round.sets[0].plan = plans[0]; round.sets[0].src = 0; round.sets[0].src2 = -1;
    round.sets[1].plan = plans[1]; round.sets[1].src = 0; round.sets[1].src2 = 1;
    round.sets[2].plan = plans[2]; round.sets[2].src = 1; round.sets[2].src2 = -1;
    round.n = n;
    round.kc = k;
    round.root = root;
    round.b_ready = (cudaEvent_t)b_ready_event;
    st = mg_distribute(mg, round, &out);
  }
  if (st == GFFM_OK && m > 0) {
    const int64_t ldt = round_up(m, 32);
    st = gffm_ws_reserve(ctx, &ctx->ws_misc, (size_t)3 * ldt * n * 4);
    if (st == GFFM_OK) {
      uint32_t* carry = (uint32_t*)ctx->ws_misc.ptr;
      uint32_t* P2 = carry + ldt * n;
      uint32_t* P3 = P2 + ldt * n;
      MatView vA1 = view_of(A1), vA2 = view_of(A2);
      MatView vP2{P2, ldt, m, n}, vP3{P3, ldt, m, n};
      auto ext_of = [&](int s) {
        ExtBPlanes x;
        x.planes = out.planes + round.sets[s].off;
        x.rowsPB = out.rowsPB;
        x.npanels = out.npanels;
        x.off = out.off;
        x.order = out.order;
        x.ready = out.ready;
        return x;
      };
      st = gffm_bplan_gemm(ctx, plans[0], view_of(C1), vA1, nullptr, n, ext_of(0), carry, ldt);
      if (st == GFFM_OK) st = gffm_bplan_gemm(ctx, plans[1], vP2, vA1, &vA2, n, ext_of(1), nullptr, 0);
      if (st == GFFM_OK) st = gffm_bplan_gemm(ctx, plans[2], vP3, vA2, nullptr, n, ext_of(2), nullptr, 0);
      if (st == GFFM_OK) st = gffm_kara_recombine(ctx, view_of(C2), view_of(C1), P2, P3, carry, ldt, N2);
    }
  }
  if (st == GFFM_OK) st = mg_round_done(mg, round, out);
  for (GemmBPlan* p : plans)
    if (p) gffm_bplan_destroy(p);
  return st;
}

// z_shard = A_shard * x mod P (mul!(z,A,x), CuModMatrix.jl:816-836, on row blocks): x is tiny and broadcast whole
extern "C" int32_t gffm_mg_gemv(gffm_mg* mg, gffm_mat* z, gffm_mat* A, gffm_mat* x, int32_t root, uint64_t R, uint64_t P) {
  GFFM_TRY(mg_check_common(mg, root));
  GFFM_ENTER_CTX(mg->ctx);
  if (!z || !A || !x) GFFM_FAIL(GFFM_ERR_INVALID, "null");
  GFFM_NARROW_ONLY(z, A, x);
  if (x->cols != 1 || z->cols != 1 || A->cols != x->rows || A->rows != z->rows) GFFM_FAIL(GFFM_ERR_SIZE_MISMATCH, "mg_gemv: inconsistent sizes");
  NcclApi* nc = nccl_api();
  if (mg->nranks > 1 && x->rows > 0) {
    if (mg->rank != root) gffm_touch(x);
    MG_NCCL(nc->Broadcast(x->data, x->data, (size_t)x->rows, ncclUint32, root, mg->comm, mg->ctx->stream));
  }
  return gffm_gemv(z, A, x, R, P);
}
