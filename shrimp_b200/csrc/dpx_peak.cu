// Integer-pipe (DPX) peak micro-benchmark: register-resident chains of VIADDMNMX.S16x2, the
// instruction the sw_vector recurrence is built from.  MEASURED_PEAKS.json has no integer-pipe
// figure, so bench.py measures the denominator of the sw_vector roofline live with this kernel
// (SURVEY.md section 8(d)).
#include "common.cuh"

namespace shrimp {

template <int CHAINS>
__global__ void __launch_bounds__(256) dpx_peak_kernel(uint32_t *out, int iters, uint32_t a0, uint32_t b0) {
  uint32_t v[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) v[c] = a0 + threadIdx.x * (c + 1);
  const uint32_t k1 = b0, k2 = b0 ^ 0x00070003u;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int c = 0; c < CHAINS; c++) v[c] = __viaddmax_s16x2(v[c], k1, k2);
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) acc ^= v[c];
  if (acc == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;  // keep the chains alive
}

}  // namespace shrimp

using namespace shrimp;

// Returns giga warp-lane DPX instructions per second (thread-instructions / 1e9 / s).
extern "C" int shrimp_gpu_dpx_peak(shrimp_gpu_ctx *ctx, double *ginstr_per_s) {
  if (!ctx || !ginstr_per_s) {
    set_error("shrimp_gpu_dpx_peak: NULL argument");
    return SHRIMP_E_ARG;
  }
  SH_CUDA(cudaSetDevice(ctx->device));
  const int CH = 8, iters = 4096, threads = 256;
  const int blocks = ctx->sm_count * 8;
  SH_TRY(ctx->d_scores.ensure((size_t)blocks * threads * 4));
  cudaEvent_t e0, e1;
  SH_CUDA(cudaEventCreate(&e0));
  SH_CUDA(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    SH_CUDA(cudaEventRecord(e0, ctx->stream));
    dpx_peak_kernel<CH><<<blocks, threads, 0, ctx->stream>>>(ctx->d_scores.as<uint32_t>(), iters, 0x00010002u, 0xfffdfffeu);
    SH_CUDA(cudaGetLastError());
    SH_CUDA(cudaEventRecord(e1, ctx->stream));
    SH_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    SH_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double instr = (double)blocks * threads * (double)iters * 8 * CH;
    double g = instr / (ms * 1e-3) / 1e9;
    if (rep > 0 && g > best) best = g;
    ctx->launches++;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ginstr_per_s = best;
  return SHRIMP_OK;
}

// FP64 issue peak: register-resident chains of DFMA, the instruction post_sw's libm transcription is mostly made of.
// MEASURED_PEAKS.json has no FP64 figure either; bench.py takes the denominator of the post_sw roofline from here.
namespace shrimp {
template <int CHAINS>
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a0, double b0) {
  double v[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) v[c] = a0 + (double)(threadIdx.x * (c + 1));
  const double k1 = b0, k2 = 1.0 - b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int c = 0; c < CHAINS; c++) v[c] = fma(v[c], k1, k2);
    }
  }
  double acc = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) acc += v[c];
  if (acc == 0.123456789) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;  // keep the chains alive
}
}  // namespace shrimp

// Returns giga thread-level FP64 instructions (DFMA) per second.
extern "C" int shrimp_gpu_fp64_peak(shrimp_gpu_ctx *ctx, double *ginstr_per_s) {
  if (!ctx || !ginstr_per_s) {
    set_error("shrimp_gpu_fp64_peak: NULL argument");
    return SHRIMP_E_ARG;
  }
  SH_CUDA(cudaSetDevice(ctx->device));
  const int CH = 8, iters = 1024, threads = 256;
  const int blocks = ctx->sm_count * 8;
  SH_TRY(ctx->d_scores.ensure((size_t)blocks * threads * 8));
  cudaEvent_t e0, e1;
  SH_CUDA(cudaEventCreate(&e0));
  SH_CUDA(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    SH_CUDA(cudaEventRecord(e0, ctx->stream));
    fp64_peak_kernel<CH><<<blocks, threads, 0, ctx->stream>>>(ctx->d_scores.as<double>(), iters, 0.5, 0.999999);
    SH_CUDA(cudaGetLastError());
    SH_CUDA(cudaEventRecord(e1, ctx->stream));
    SH_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    SH_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double instr = (double)blocks * threads * (double)iters * 8 * CH;
    const double g = instr / (ms * 1e-3) / 1e9;
    if (rep > 0 && g > best) best = g;
    ctx->launches++;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ginstr_per_s = best;
  return SHRIMP_OK;
}
