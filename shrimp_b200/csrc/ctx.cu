// Context, error plumbing and scoring set-up of libshrimp_b200.so.
#include <stdarg.h>
#include <stdlib.h>
#define SHRIMP_NO_SYNC_WRAP   // this file defines the wrapper
#include "common.cuh"

namespace shrimp {
static int blocking_sync_mode() {
  // measured on an 8-GPU box with four host cores per GPU: blocking waits leave the end-to-end rate where it is
  // (95 M reads/s) and cost the device-resident path 10 % (a wake-up per wait): spinning stays the default
  static const int mode = [] {
    const char *e = getenv("SHRIMP_BLOCKING_SYNC");
    return (e && atoi(e) != 0) ? 1 : 0;
  }();
  return mode;
}
cudaError_t sync_stream(cudaStream_t s) {
  if (!blocking_sync_mode()) return cudaStreamSynchronize(s);
  static thread_local cudaEvent_t ev[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaStreamSynchronize(s);
  if (!ev[dev]) {
    e = cudaEventCreateWithFlags(&ev[dev], cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  e = cudaEventRecord(ev[dev], s);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ev[dev]);
}

static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void free_genome(shrimp_gpu_ctx *ctx);
void free_pipeline(shrimp_gpu_ctx *ctx);
}  // namespace shrimp

using namespace shrimp;

extern "C" const char *shrimp_gpu_last_error(void) { return g_err; }

extern "C" int shrimp_gpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int shrimp_gpu_create(int device, shrimp_gpu_ctx **out) {
  if (!out) {
    set_error("shrimp_gpu_create: out == NULL");
    return SHRIMP_E_ARG;
  }
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s); libshrimp_b200 has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    cudaGetLastError();
    return SHRIMP_E_CUDA;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range [0,%d)", device, n);
    return SHRIMP_E_ARG;
  }
  SH_CUDA(cudaSetDevice(device));
  // The seed scan gathers short index lists (a handful of 4-byte positions each): ask for 32-byte L2 fetches so that a
  // list costs the sectors it touches, not whole 64/128-byte lines (a hint; SHRIMP_L2_FETCH=64|128 overrides)
  {
    size_t gran = 32;
    if (const char *e = getenv("SHRIMP_L2_FETCH")) gran = (size_t)atoi(e);
    if (gran == 32 || gran == 64 || gran == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    cudaGetLastError();
  }
  shrimp_gpu_ctx *c = new shrimp_gpu_ctx();
  c->device = device;
  cudaDeviceProp prop;
  SH_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  SH_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  SH_CUDA(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
  for (int i = 0; i < SHRIMP_AUX_STREAMS; i++) {
    SH_CUDA(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
    SH_CUDA(cudaEventCreateWithFlags(&c->join_ev[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < ST_COUNT; i++) {
    SH_CUDA(cudaEventCreate(&c->timers[i].ev0));
    SH_CUDA(cudaEventCreate(&c->timers[i].ev1));
  }
  *out = c;
  return SHRIMP_OK;
}

extern "C" void shrimp_gpu_destroy(shrimp_gpu_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  free_pipeline(c);
  free_genome(c);
  c->d_genome.release();
  c->d_genome_ls.release();
  c->d_reads.release();
  c->d_task.release();
  c->d_scores.release();
  c->d_boundary.release();
  c->d_flush.release();
  for (int i = 0; i < 2; i++)
    if (c->user_ev[i]) cudaEventDestroy(c->user_ev[i]);
  for (int i = 0; i < ST_COUNT; i++) {
    if (c->timers[i].ev0) cudaEventDestroy(c->timers[i].ev0);
    if (c->timers[i].ev1) cudaEventDestroy(c->timers[i].ev1);
  }
  for (int i = 0; i < SHRIMP_AUX_STREAMS; i++) {
    if (c->aux[i]) cudaStreamDestroy(c->aux[i]);
    if (c->join_ev[i]) cudaEventDestroy(c->join_ev[i]);
  }
  if (c->fork_ev) cudaEventDestroy(c->fork_ev);
  cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" uint64_t shrimp_gpu_launch_count(const shrimp_gpu_ctx *c) { return c ? c->launches : 0; }

static const char *k_stage_names[ST_COUNT] = {"index_build", "seed_scan", "sw_vector", "pass1_select", "sw_full", "post_sw", "other"};

extern "C" int shrimp_gpu_stage_times(shrimp_gpu_ctx *c, const char **names, float *ms, uint64_t *launches, int max_n) {
  if (!c) return 0;
  int n = ST_COUNT < max_n ? ST_COUNT : max_n;
  for (int i = 0; i < n; i++) {
    if (names) names[i] = k_stage_names[i];
    if (ms) ms[i] = c->timers[i].ms;
    if (launches) launches[i] = c->timers[i].launches;
  }
  return n;
}

extern "C" void shrimp_gpu_stage_times_reset(shrimp_gpu_ctx *c) {
  if (!c) return;
  for (int i = 0; i < ST_COUNT; i++) {
    c->timers[i].ms = 0.f;
    c->timers[i].launches = 0;
  }
}

// sw_vector_setup (sw-vector.c:388-439) + sw_full_{ls,cs}_setup argument conventions: scores come
// in with the CLI sign and are negated here; match*qrlen must stay below 2^15 for the int16 lanes.
extern "C" int shrimp_gpu_sw_setup(shrimp_gpu_ctx *c, const shrimp_sw_params *p) {
  if (!c || !p) {
    set_error("shrimp_gpu_sw_setup: NULL argument");
    return SHRIMP_E_ARG;
  }
  if (p->match <= 0 || p->mismatch >= 0 || p->a_gap_open > 0 || p->a_gap_ext > 0 || p->b_gap_open > 0 ||
      p->b_gap_ext > 0 || p->max_read_len <= 0 || p->max_window_len <= 0) {
    set_error("shrimp_gpu_sw_setup: scores must be match>0, mismatch<0, gap scores<=0, lengths>0");
    return SHRIMP_E_ARG;
  }
  if ((long long)p->match * p->max_read_len >= 32768) {
    // same guard as sw-vector.c:393
    set_error("Match Value is too high/reads are too long: match x longest_read_length must be < 32768");
    return SHRIMP_E_RANGE;
  }
  SwScores s;
  s.match = p->match;
  s.mismatch = p->mismatch;
  s.vec_mismatch = p->use_colours ? p->match + p->crossover : p->mismatch;
  s.a_open = -p->a_gap_open;
  s.a_ext = -p->a_gap_ext;
  s.b_open = -p->b_gap_open;
  s.b_ext = -p->b_gap_ext;
  s.xover = p->crossover;
  s.use_colours = p->use_colours ? 1 : 0;
  s.anchor_width = p->anchor_width;
  s.indel_taboo_len = p->indel_taboo_len;
  s.max_read_len = p->max_read_len;
  s.max_window_len = p->max_window_len;
  // smallest shift with 2^shift > match - mismatch; 5-bit codes << shift must stay below 2^15
  int sh = 1;
  if (s.vec_mismatch >= 0) {
    set_error("shrimp_gpu_sw_setup: colour space needs match + crossover < 0");
    return SHRIMP_E_ARG;
  }
  while ((1 << sh) <= p->match - s.vec_mismatch) sh++;
  if (sh > 9 || s.a_open + s.a_ext >= 16384 || s.b_open + s.b_ext >= 16384) {
    set_error("shrimp_gpu_sw_setup: |match - mismatch| or gap scores too large for the packed int16 kernel");
    return SHRIMP_E_RANGE;
  }
  s.shift = sh;
  s.valid = true;
  c->sw = s;
  return SHRIMP_OK;
}

// Two user events on the library's stream so a caller can bracket a timed region on the device
// (bench.py: CUDA-event timing of exactly K steps on the launching stream).
extern "C" int shrimp_gpu_event_record(shrimp_gpu_ctx *c, int which) {
  if (!c || which < 0 || which > 1) {
    set_error("shrimp_gpu_event_record: invalid argument");
    return SHRIMP_E_ARG;
  }
  SH_CUDA(cudaSetDevice(c->device));
  if (!c->user_ev[which]) SH_CUDA(cudaEventCreate(&c->user_ev[which]));
  SH_CUDA(cudaEventRecord(c->user_ev[which], c->stream));
  return SHRIMP_OK;
}
extern "C" int shrimp_gpu_event_elapsed_ms(shrimp_gpu_ctx *c, float *ms) {
  if (!c || !ms || !c->user_ev[0] || !c->user_ev[1]) {
    set_error("shrimp_gpu_event_elapsed_ms: events not recorded");
    return SHRIMP_E_STATE;
  }
  SH_CUDA(cudaEventSynchronize(c->user_ev[1]));
  SH_CUDA(cudaEventElapsedTime(ms, c->user_ev[0], c->user_ev[1]));
  return SHRIMP_OK;
}
// Writes a buffer larger than L2 (126 MB) on the library's stream: L2 flush between timed steps.
extern "C" int shrimp_gpu_flush_l2(shrimp_gpu_ctx *c) {
  if (!c) return SHRIMP_E_ARG;
  SH_CUDA(cudaSetDevice(c->device));
  const size_t bytes = (size_t)256 << 20;
  SH_TRY(c->d_flush.ensure(bytes));
  SH_CUDA(cudaMemsetAsync(c->d_flush.p, 1, bytes, c->stream));
  SH_CUDA(cudaStreamSynchronize(c->stream));
  return SHRIMP_OK;
}
