// Internal header of libshrimp_b200.so: context, error plumbing, device buffers, bit helpers.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/shrimp_b200.h"

namespace shrimp {

void set_error(const char *fmt, ...);

// Every wait of a host thread for its stream goes through here: SHRIMP_BLOCKING_SYNC=1 makes it block on an event
// instead of spinning (for hosts that need the core; see ctx.cu for what was measured).
cudaError_t sync_stream(cudaStream_t s);
#ifndef SHRIMP_NO_SYNC_WRAP
#define cudaStreamSynchronize(s) shrimp::sync_stream(s)
#endif

#define SH_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (call);                                                            \
    if (_e != cudaSuccess) {                                                            \
      shrimp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return SHRIMP_E_CUDA;                                                             \
    }                                                                                   \
  } while (0)

#define SH_TRY(call)            \
  do {                          \
    int _r = (call);            \
    if (_r != SHRIMP_OK) return _r; \
  } while (0)

// Growable device allocation; never shrinks (sized for 180 GB of HBM3e, buffers are reused
// across chunks so steady-state mapping does no cudaMalloc).
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return SHRIMP_OK;
    if (p) cudaFree(p);
    p = nullptr;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cap = 0;
      p = nullptr;
      set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
      return SHRIMP_E_NOMEM;
    }
    cap = want;
    return SHRIMP_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T> T *as() const { return (T *)p; }
};

// Pinned host staging buffer (H2D/D2H run at full PCIe/C2C rate only from pinned memory).
struct HostBuf {
  void *p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return SHRIMP_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) {
      cap = 0;
      set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
      return SHRIMP_E_NOMEM;
    }
    cap = want;
    return SHRIMP_OK;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T> T *as() const { return (T *)p; }
};

enum Stage {
  ST_INDEX = 0,
  ST_SCAN,
  ST_VECTOR,
  ST_PASS1,
  ST_FULL,
  ST_POST,
  ST_OTHER,
  ST_COUNT
};

struct StageTimer {
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float ms = 0.f;
  uint64_t launches = 0;
};

// positive-sign score block used by the kernels
struct SwScores {
  int match, mismatch;      // mismatch < 0 (full SW)
  int vec_mismatch;         // what the vector filter uses: colour space = match + crossover (gmapper.c:2935)
  int a_open, a_ext;        // >= 0 (negated CLI values, sw-vector.c:422-425)
  int b_open, b_ext;
  int xover;                // < 0
  int use_colours;
  int anchor_width;
  int indel_taboo_len;
  int max_read_len, max_window_len;
  int shift;                // code shift of the packed-int16 equality trick (sw_vector.cu)
  bool valid = false;
};

}  // namespace shrimp

// Kernels with dynamic shared memory are always opted in to the device maximum (227 KB per CTA on sm_100): the
// attribute is per function and process-wide, so setting it to the size of ONE launch would race with the launches of
// the other host threads' contexts (a smaller value set in between makes a launch fail with "invalid argument").
#define SHRIMP_MAX_DYN_SMEM (227 * 1024)
// ... minus the kernel's static shared memory; once per kernel and device
#define SH_OPT_IN_SMEM(K, DEVICE)                                                                              \
  do {                                                                                                         \
    static std::atomic<unsigned long long> _done{0ull};                                                        \
    const unsigned long long _bit = 1ull << ((DEVICE) & 63);                                                   \
    if (!(_done.load() & _bit)) {                                                                              \
      cudaFuncAttributes _fa;                                                                                  \
      SH_CUDA(cudaFuncGetAttributes(&_fa, K));                                                                 \
      SH_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize,                             \
                                   SHRIMP_MAX_DYN_SMEM - (int)_fa.sharedSizeBytes));                           \
      _done.fetch_or(_bit);                                                                                    \
    }                                                                                                          \
  } while (0)

#define SHRIMP_AUX_STREAMS 4

struct shrimp_gpu_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  // side streams + fork/join events: independent launches of one stage (the ring-width classes of the
  // full SW) run concurrently and join back into `stream`
  cudaStream_t aux[SHRIMP_AUX_STREAMS] = {nullptr};
  cudaEvent_t fork_ev = nullptr, join_ev[SHRIMP_AUX_STREAMS] = {nullptr};
  uint64_t launches = 0;
  shrimp::SwScores sw;
  shrimp::StageTimer timers[shrimp::ST_COUNT];
  // sw_vector batch scratch
  shrimp::DevBuf d_genome, d_genome_ls, d_reads, d_task, d_scores, d_boundary;
  // the caller's timing bracket and the L2 flush buffer (bench entries): per context, freed with it
  cudaEvent_t user_ev[2] = {nullptr, nullptr};
  shrimp::DevBuf d_flush;
  // opaque owners of the resident genome/index and the chunk pipeline (index.cu / pipeline.cu)
  void *genome = nullptr;
  bool genome_borrowed = false;   // shared from another context (shrimp_gpu_share_genome): not freed here
  void *pipeline = nullptr;
};

namespace shrimp {

static inline int ceil_div_i(long long a, long long b) { return (int)((a + b - 1) / b); }

// Stage timing: events recorded on ctx->stream around each stage; read back lazily.
struct ScopedStage {
  shrimp_gpu_ctx *c;
  int st;
  ScopedStage(shrimp_gpu_ctx *ctx, int stage) : c(ctx), st(stage) {
    cudaEventRecord(c->timers[st].ev0, c->stream);
  }
  ~ScopedStage() {
    cudaEventRecord(c->timers[st].ev1, c->stream);
    cudaEventSynchronize(c->timers[st].ev1);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->timers[st].ev0, c->timers[st].ev1) == cudaSuccess) c->timers[st].ms += ms;
  }
};

#define SH_LAUNCHED(ctx, stage) \
  do {                          \
    (ctx)->launches++;          \
    (ctx)->timers[stage].launches++; \
  } while (0)

__host__ __device__ __forceinline__ uint32_t extract4(const uint32_t *a, uint64_t i) {
  return (a[i >> 3] >> (4u * (uint32_t)(i & 7))) & 0xfu;
}

// device entry used by both the raw batch API and the chunk pipeline (sw_vector.cu)
struct VecTaskArrays {
  const uint32_t *goff;   // nibble offset into genome (and genome_ls)
  const int32_t *glen;
  const int32_t *ridx;    // read index into reads (stride words each)
  const int32_t *rlen;
  const int8_t *initbp;   // colour space only
  const uint32_t *out;    // optional: scores[out[t]] receives the score of task t (dense task lists); else scores[t]
};
int launch_sw_vector(shrimp_gpu_ctx *ctx, const uint32_t *d_genome, const uint32_t *d_genome_ls,
                     const uint32_t *d_reads, int read_stride_words, int n_tasks, int max_rlen, int max_glen,
                     const VecTaskArrays &t, int32_t *d_scores, int stage);

}  // namespace shrimp
