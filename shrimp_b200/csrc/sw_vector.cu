// Inter-task vector Smith-Waterman filter on Blackwell DPX (packed int16x2 add/max).
//
// Replaces common/sw-vector.c (sw_vector :453-515, vect_sw_same_gap/diff_gap :68-377): score-only
// affine-gap local alignment of a read (rows) against a genome window (columns),
//     E(i,j) = max(E(i,j-1) - a_ext, H(i,j-1) - a_open - a_ext)
//     F(i,j) = max(F(i-1,j) - b_ext, H(i-1,j) - b_open - b_ext)
//     H(i,j) = max(0, H(i-1,j-1) + s(i,j), E(i,j), F(i,j)),   s = match if codes equal else mismatch
// and, in colour space, row 0 scored against lstocs(letter genome, initbp) (sw-vector.c:116-146).
//
// Design (B200-first, not the SSE anti-diagonal layout):
//  * one thread owns TWO independent tasks, one per 16-bit lane of every register, so each DPX
//    instruction (VIADDMNMX.S16x2 / VIMNMX.S16x2) advances two cells;
//  * a strip of T read rows lives in registers (H of the previous column and E per row); the thread
//    walks the window column by column, so F and the diagonal are carried in registers too.  Reads
//    longer than T take several strips, with the strip's last row (H, F per column) parked in a
//    column-major global scratch that stays in L2;
//  * the query (read codes, pre-shifted, both lanes) is staged in shared memory as q[row][thread]
//    (conflict-free) -- one LDS per cell pair instead of T more registers;
//  * the substitution score needs no table: codes are stored shifted left by `sh` bits with
//    2^sh > match - mismatch; with the genome code complemented, x = ~(d<<sh) ^ (q<<sh) is -1 on a
//    match and <= -1-2^sh otherwise, so s = max(x + match + 1, mismatch) is ONE viaddmax;
//  * E and F are kept offset by their open+extend cost (E~ = E + ao + ae, F~ = F + bo + be) so every
//    add of the recurrence is fused into a max: 8 integer-pipe instructions per cell pair.
// Padded rows/columns use codes that match nothing; a cell that is only reachable through mismatches
// and gaps can never exceed the cell it came from, so padding cannot change the maximum.
#include "common.cuh"

namespace shrimp {

#define SWV_BLOCK 128

struct SwvParams {
  const uint32_t *genome;
  const uint32_t *genome_ls;
  const uint32_t *reads;
  int read_stride;
  int n_tasks;
  VecTaskArrays t;
  int32_t *scores;
  uint32_t *boundary;       // [2][max_glen][n_threads] (H then F~), multi-strip only
  uint32_t n_threads;
  int n_strips;
  int max_glen;
  uint32_t ma1, mm, naoe, nae, nboe, nbe;  // packed int16x2 constants
  int sh;
};

__device__ __forceinline__ uint32_t dup16(int v) { return ((uint32_t)v & 0xffffu) * 0x10001u; }

__device__ __forceinline__ uint32_t vmax2(uint32_t a, uint32_t b) { return __vimax3_s16x2(a, b, b); }

// colour of (letter, initbp) as lstocs() (util.h:182-207) gives it: XOR of 2-bit letters, N otherwise
__device__ __forceinline__ uint32_t colour_of(uint32_t letter, int initbp) {
  return (letter > 3u || (uint32_t)initbp > 3u) ? 15u : (letter ^ (uint32_t)initbp);
}

// Nibble stream over a packed sequence starting at an arbitrary base offset.
struct NibbleStream {
  const uint32_t *p;
  uint32_t word;
  int rem;
  __device__ __forceinline__ void init(const uint32_t *base, uint64_t pos) {
    p = base + (pos >> 3);
    int ph = (int)(pos & 7);
    word = __ldg(p) >> (4 * ph);
    rem = 8 - ph;
  }
  __device__ __forceinline__ uint32_t next() {
    uint32_t c = word & 15u;
    word >>= 4;
    if (--rem == 0) {
      word = __ldg(++p);
      rem = 8;
    }
    return c;
  }
};

template <int T, bool CS>
__global__ void __launch_bounds__(SWV_BLOCK) sw_vector_kernel(const SwvParams P) {
  extern __shared__ uint32_t q_s[];  // [T][SWV_BLOCK]
  const uint32_t tid = threadIdx.x;
  const uint32_t gtid = blockIdx.x * SWV_BLOCK + tid;
  const int tA = 2 * (int)gtid;
  if (tA >= P.n_tasks) return;
  const bool hasB = (tA + 1) < P.n_tasks;
  const int tB = hasB ? tA + 1 : tA;

  // glen <= 0 marks a placeholder slot (pipeline hit slots that are gaps or ineligible): its other
  // fields may be stale, so they are not read and the lane runs on padding only.
  const int glenA = P.t.glen[tA], glenB = P.t.glen[tB];
  const bool okA = glenA > 0, okB = glenB > 0;
  if (!okA && !okB) return;
  const int rlenA = okA ? P.t.rlen[tA] : 0, rlenB = okB ? P.t.rlen[tB] : 0;
  const uint64_t goffA = okA ? P.t.goff[tA] : 0, goffB = okB ? P.t.goff[tB] : 0;
  const uint32_t *readA = P.reads + (okA ? (size_t)P.t.ridx[tA] * P.read_stride : 0);
  const uint32_t *readB = P.reads + (okB ? (size_t)P.t.ridx[tB] * P.read_stride : 0);
  const int ncols = max(glenA, glenB);
  const int sh = P.sh;
  const uint32_t DPAD = 16u << sh, QPAD = 17u << sh;
  int ibA = 0, ibB = 0;
  if (CS) {
    ibA = okA ? P.t.initbp[tA] : 0;
    ibB = okB ? P.t.initbp[tB] : 0;
  }

  const uint32_t MA1 = P.ma1, MM = P.mm, NAOE = P.naoe, NAE = P.nae, NBOE = P.nboe, NBE = P.nbe;
  uint32_t best = 0;

  for (int strip = 0; strip < P.n_strips; strip++) {
    const int row0 = strip * T;
    // stage this strip's query rows: q[r][tid] = (codeA << sh) | (codeB << sh) << 16
#pragma unroll 4
    for (int r = 0; r < T; r++) {
      int i = row0 + r;
      uint32_t qa = i < rlenA ? (extract4(readA, i) << sh) : QPAD;
      uint32_t qb = i < rlenB ? (extract4(readB, i) << sh) : QPAD;
      q_s[r * SWV_BLOCK + tid] = qa | (qb << 16);
    }
    // q_s column `tid` is private to this thread: no barrier needed.

    uint32_t h[T], e[T];
#pragma unroll
    for (int r = 0; r < T; r++) {
      h[r] = 0;
      e[r] = 0;
    }
    NibbleStream gA, gB, lA, lB;
    gA.init(P.genome, goffA);
    gB.init(P.genome, goffB);
    if (CS && strip == 0) {
      lA.init(P.genome_ls, goffA);
      lB.init(P.genome_ls, goffB);
    }
    uint32_t diag = 0;
    uint32_t *bH = P.boundary + gtid;
    uint32_t *bF = bH + (size_t)P.max_glen * P.n_threads;

    for (int j = 0; j < ncols; j++) {
      uint32_t cA = gA.next(), cB = gB.next();
      uint32_t dA = j < glenA ? (cA << sh) : DPAD;
      uint32_t dB = j < glenB ? (cB << sh) : DPAD;
      const uint32_t dn = ~(dA | (dB << 16));
      uint32_t dn0 = dn;
      if (CS && strip == 0) {
        uint32_t a0 = colour_of(lA.next(), ibA), b0 = colour_of(lB.next(), ibB);
        uint32_t d0A = j < glenA ? (a0 << sh) : DPAD;
        uint32_t d0B = j < glenB ? (b0 << sh) : DPAD;
        dn0 = ~(d0A | (d0B << 16));
      }
      uint32_t hd = diag, f = 0, hd_pair = 0;
      if (strip > 0) {
        size_t o = (size_t)j * P.n_threads;
        diag = bH[o];  // H(row0-1, j): diagonal input of the next column
        f = bF[o];
      }
#pragma unroll
      for (int r = 0; r < T; r++) {
        const uint32_t x = ((CS && r == 0) ? dn0 : dn) ^ q_s[r * SWV_BLOCK + tid];
        const uint32_t s = __viaddmax_s16x2(x, MA1, MM);
        const uint32_t u = __viaddmax_s16x2_relu(hd, s, 0u);
        const uint32_t v = __viaddmax_s16x2(e[r], NAOE, u);
        const uint32_t H = __viaddmax_s16x2(f, NBOE, v);
        hd = h[r];
        h[r] = H;
        e[r] = __viaddmax_s16x2(e[r], NAE, H);
        f = __viaddmax_s16x2(f, NBE, H);
        if (r & 1) best = __vimax3_s16x2(best, hd_pair, H);   // one three-way maximum per two rows (T is even)
        else hd_pair = H;
      }
      if (P.n_strips > 1) {
        size_t o = (size_t)j * P.n_threads;
        bH[o] = h[T - 1];
        bF[o] = f;
      }
    }
  }
  // tasks with glen <= 0 are placeholders (pipeline slots that pass 1 will never read)
  if (glenA > 0) P.scores[P.t.out ? P.t.out[tA] : (uint32_t)tA] = (int)(int16_t)(best & 0xffffu);
  if (hasB && glenB > 0) P.scores[P.t.out ? P.t.out[tB] : (uint32_t)tB] = (int)(int16_t)(best >> 16);
}

template <int T>
static int launch_T(shrimp_gpu_ctx *ctx, const SwvParams &P, bool cs, int n_blocks) {
  size_t smem = (size_t)T * SWV_BLOCK * sizeof(uint32_t);
  if (cs) {
    sw_vector_kernel<T, true><<<n_blocks, SWV_BLOCK, smem, ctx->stream>>>(P);
  } else {
    sw_vector_kernel<T, false><<<n_blocks, SWV_BLOCK, smem, ctx->stream>>>(P);
  }
  SH_CUDA(cudaGetLastError());
  return SHRIMP_OK;
}

// rows per register strip: minimise padded rows + per-strip overhead, prefer fewer strips
static int choose_T(int max_rlen) {
  static const int cand[] = {8, 16, 24, 32, 40, 48, 56, 64};
  int bestT = 64;
  long bestCost = -1;
  for (int T : cand) {
    long strips = (max_rlen + T - 1) / T;
    long cost = strips * (T + 3);
    if (bestCost < 0 || cost < bestCost || (cost == bestCost && T > bestT)) {
      bestCost = cost;
      bestT = T;
    }
  }
  return bestT;
}

int launch_sw_vector(shrimp_gpu_ctx *ctx, const uint32_t *d_genome, const uint32_t *d_genome_ls,
                     const uint32_t *d_reads, int read_stride_words, int n_tasks, int max_rlen, int max_glen,
                     const VecTaskArrays &t, int32_t *d_scores, int stage) {
  if (n_tasks <= 0) return SHRIMP_OK;
  const SwScores &s = ctx->sw;
  if (!s.valid) {
    set_error("sw_vector: shrimp_gpu_sw_setup() has not been called");
    return SHRIMP_E_STATE;
  }
  if ((long long)s.match * max_rlen >= 32768) {
    set_error("sw_vector: match x read length >= 32768");
    return SHRIMP_E_RANGE;
  }
  if (s.use_colours && d_genome_ls == nullptr) {
    set_error("sw_vector: colour space needs the letter genome for row 0");
    return SHRIMP_E_ARG;
  }
  const int T = choose_T(max_rlen);
  const int n_strips = (max_rlen + T - 1) / T;
  // reads longer than one register strip park a boundary row pair per thread and window column in global memory:
  // such launches are cut into batches whose boundary rows stay under 1 GB (the sensitive configuration scores
  // hundreds of millions of windows per chunk)
  int batch = n_tasks;
  if (n_strips > 1) {
    const size_t per_thread = (size_t)2 * (size_t)std::max(1, max_glen) * sizeof(uint32_t);
    const size_t max_threads = std::max<size_t>((size_t)SWV_BLOCK, ((size_t)1 << 30) / per_thread);
    batch = (int)std::min<size_t>((size_t)n_tasks, 2 * (max_threads / SWV_BLOCK) * SWV_BLOCK);
  }
  int rc = SHRIMP_OK;
  for (int t0 = 0; t0 < n_tasks && rc == SHRIMP_OK; t0 += batch) {
    const int nb = std::min(batch, n_tasks - t0);
    SwvParams P;
    P.genome = d_genome;
    P.genome_ls = d_genome_ls;
    P.reads = d_reads;
    P.read_stride = read_stride_words;
    P.n_tasks = nb;
    P.t = t;
    P.t.goff += t0;
    P.t.glen += t0;
    P.t.ridx += t0;
    P.t.rlen += t0;
    if (P.t.initbp) P.t.initbp += t0;
    if (P.t.out) P.t.out += t0;
    P.scores = t.out ? d_scores : d_scores + t0;
    P.n_strips = n_strips;
    P.max_glen = max_glen;
    const int n_pairs = (nb + 1) / 2;
    const int n_blocks = (n_pairs + SWV_BLOCK - 1) / SWV_BLOCK;
    P.n_threads = (uint32_t)n_blocks * SWV_BLOCK;
    P.boundary = nullptr;
    if (P.n_strips > 1) {
      SH_TRY(ctx->d_boundary.ensure((size_t)2 * max_glen * P.n_threads * sizeof(uint32_t)));
      P.boundary = ctx->d_boundary.as<uint32_t>();
    }
    auto pk = [](int v) { return ((uint32_t)v & 0xffffu) * 0x10001u; };
    P.ma1 = pk(s.match + 1);
    P.mm = pk(s.vec_mismatch);
    P.naoe = pk(-(s.a_open + s.a_ext));
    P.nae = pk(-s.a_ext);
    P.nboe = pk(-(s.b_open + s.b_ext));
    P.nbe = pk(-s.b_ext);
    P.sh = s.shift;
    const bool cs = s.use_colours != 0;
    switch (T) {
      case 8: rc = launch_T<8>(ctx, P, cs, n_blocks); break;
      case 16: rc = launch_T<16>(ctx, P, cs, n_blocks); break;
      case 24: rc = launch_T<24>(ctx, P, cs, n_blocks); break;
      case 32: rc = launch_T<32>(ctx, P, cs, n_blocks); break;
      case 40: rc = launch_T<40>(ctx, P, cs, n_blocks); break;
      case 48: rc = launch_T<48>(ctx, P, cs, n_blocks); break;
      case 56: rc = launch_T<56>(ctx, P, cs, n_blocks); break;
      default: rc = launch_T<64>(ctx, P, cs, n_blocks); break;
    }
    if (rc == SHRIMP_OK) SH_LAUNCHED(ctx, stage);
  }
  return rc;
}


}  // namespace shrimp

using namespace shrimp;

// sw_vector (sw-vector.c:453) for a batch of independent tasks, host buffers in, scores out.
extern "C" int shrimp_gpu_sw_vector_batch(shrimp_gpu_ctx *ctx, const uint32_t *genome, size_t genome_words,
                                          const uint32_t *genome_ls, const uint32_t *reads, int read_stride_words,
                                          int n_reads, int n_tasks, const uint32_t *goff, const int32_t *glen,
                                          const int32_t *read_idx, const int32_t *rlen, const int8_t *initbp,
                                          int32_t *scores_out) {
  if (!ctx || !genome || !reads || !goff || !glen || !read_idx || !rlen || !scores_out || n_tasks < 0 ||
      n_reads <= 0 || read_stride_words <= 0) {
    set_error("shrimp_gpu_sw_vector_batch: invalid argument");
    return SHRIMP_E_ARG;
  }
  if (!ctx->sw.valid) {
    set_error("shrimp_gpu_sw_vector_batch: shrimp_gpu_sw_setup() has not been called");
    return SHRIMP_E_STATE;
  }
  const bool cs = ctx->sw.use_colours != 0;
  if (cs && (!genome_ls || !initbp)) {
    set_error("shrimp_gpu_sw_vector_batch: colour space needs genome_ls and initbp");
    return SHRIMP_E_ARG;
  }
  if (n_tasks == 0) return SHRIMP_OK;
  int max_rlen = 0, max_glen = 0;
  for (int i = 0; i < n_tasks; i++) {
    if (glen[i] <= 0 || rlen[i] <= 0 || read_idx[i] < 0 || read_idx[i] >= n_reads ||
        (uint64_t)goff[i] + (uint64_t)glen[i] > (uint64_t)genome_words * 8 || rlen[i] > read_stride_words * 8) {
      set_error("shrimp_gpu_sw_vector_batch: task %d out of range", i);
      return SHRIMP_E_ARG;
    }
    if (glen[i] > ctx->sw.max_window_len || rlen[i] > ctx->sw.max_read_len) {
      set_error("shrimp_gpu_sw_vector_batch: task %d exceeds the dblen/qrlen given at setup", i);
      return SHRIMP_E_ARG;
    }
    if (rlen[i] > max_rlen) max_rlen = rlen[i];
    if (glen[i] > max_glen) max_glen = glen[i];
  }
  SH_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t gbytes = genome_words * sizeof(uint32_t);
  SH_TRY(ctx->d_genome.ensure(gbytes + 16));
  SH_CUDA(cudaMemsetAsync((char *)ctx->d_genome.p + gbytes, 0, 16, st));
  SH_CUDA(cudaMemcpyAsync(ctx->d_genome.p, genome, gbytes, cudaMemcpyHostToDevice, st));
  if (cs) {
    SH_TRY(ctx->d_genome_ls.ensure(gbytes + 16));
    SH_CUDA(cudaMemsetAsync((char *)ctx->d_genome_ls.p + gbytes, 0, 16, st));
    SH_CUDA(cudaMemcpyAsync(ctx->d_genome_ls.p, genome_ls, gbytes, cudaMemcpyHostToDevice, st));
  }
  const size_t rbytes = (size_t)n_reads * read_stride_words * sizeof(uint32_t);
  SH_TRY(ctx->d_reads.ensure(rbytes));
  SH_CUDA(cudaMemcpyAsync(ctx->d_reads.p, reads, rbytes, cudaMemcpyHostToDevice, st));
  // task arrays packed in one buffer: goff | glen | ridx | rlen | initbp
  const size_t n = (size_t)n_tasks;
  const size_t tbytes = n * 4 * 4 + ((n + 3) & ~(size_t)3);
  SH_TRY(ctx->d_task.ensure(tbytes));
  char *tb = (char *)ctx->d_task.p;
  SH_CUDA(cudaMemcpyAsync(tb, goff, n * 4, cudaMemcpyHostToDevice, st));
  SH_CUDA(cudaMemcpyAsync(tb + n * 4, glen, n * 4, cudaMemcpyHostToDevice, st));
  SH_CUDA(cudaMemcpyAsync(tb + n * 8, read_idx, n * 4, cudaMemcpyHostToDevice, st));
  SH_CUDA(cudaMemcpyAsync(tb + n * 12, rlen, n * 4, cudaMemcpyHostToDevice, st));
  if (cs) SH_CUDA(cudaMemcpyAsync(tb + n * 16, initbp, n, cudaMemcpyHostToDevice, st));
  SH_TRY(ctx->d_scores.ensure(n * 4));
  VecTaskArrays t;
  t.goff = (const uint32_t *)tb;
  t.glen = (const int32_t *)(tb + n * 4);
  t.ridx = (const int32_t *)(tb + n * 8);
  t.rlen = (const int32_t *)(tb + n * 12);
  t.initbp = cs ? (const int8_t *)(tb + n * 16) : nullptr;
  t.out = nullptr;
  {
    ScopedStage ss(ctx, ST_VECTOR);
    SH_TRY(launch_sw_vector(ctx, ctx->d_genome.as<uint32_t>(), cs ? ctx->d_genome_ls.as<uint32_t>() : nullptr,
                            ctx->d_reads.as<uint32_t>(), read_stride_words, n_tasks, max_rlen, max_glen, t,
                            ctx->d_scores.as<int32_t>(), ST_VECTOR));
  }
  SH_CUDA(cudaMemcpyAsync(scores_out, ctx->d_scores.p, n * 4, cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaStreamSynchronize(st));
  return SHRIMP_OK;
}
