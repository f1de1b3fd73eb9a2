// Pass 1 around the vector filter: task construction for sw_vector, the f1 window-cache hash,
// the sequential replay of read_pass1_per_strand and the bounded top-k heap.
//
// The reference scores windows one by one and lets earlier results decide whether later windows
// are scored at all (overlap skip, mapping.c:1287-1293) or looked up in a lossy 2^20-slot cache
// (f1-wrapper.h:97-134).  On the device every eligible window is scored first (sw_vector.cu, a
// superset of what the reference scores); `pass1_select_kernel` then replays the reference's
// sequential rules per read strand in list order -- skipped windows get score 0, windows whose cache
// slot was written earlier in the same (read, strand) pass get the writer's score -- and finally
// emulates extheap_unpaired_pass1 (heap.h:226-307, mapping.c:1376-1411) so that the set AND order of
// hits handed to pass 2 are the reference's.
#include "stages.cuh"

namespace shrimp {


// hash_accumulate / hash_finalize (common/hash.h:69-92), hash_genome_window (util.h:220-241)
// The window is read a word (8 codes) at a time: a group of 16 codes is two words at the window's fixed
// misalignment, squeezed to 2 bits per code and reversed pairwise (the reference shifts the codes in first-to-last,
// so the first code ends up most significant; a short last group is right-aligned).
__device__ __forceinline__ uint32_t squeeze8_pairs(uint32_t w) {   // 8 codes -> 16 bits, code j at bits 2j
  w &= 0x33333333u;
  w = (w | (w >> 2)) & 0x0f0f0f0fu;
  w = (w | (w >> 4)) & 0x00ff00ffu;
  w = (w | (w >> 8)) & 0x0000ffffu;
  return w;
}
__device__ __forceinline__ uint32_t hash_genome_window_dev(const uint32_t *genome, uint64_t goff, uint32_t glen) {
  uint32_t key = 0;
  const uint32_t *wp = genome + (goff >> 3);
  const uint32_t sh = (uint32_t)(goff & 7u), shb = 4u * sh;
  uint32_t w0 = glen ? wp[0] : 0u;
  for (uint32_t i = 0; i < (glen + 15) / 16; i++) {
    const uint32_t cnt = min(16u, glen - i * 16u), need = sh + cnt;   // codes of this group, codes from wp[2i] on
    const uint32_t w1 = need > 8u ? wp[2 * i + 1] : 0u, w2 = need > 16u ? wp[2 * i + 2] : 0u;
    const uint32_t lo = __funnelshift_r(w0, w1, shb), hi = __funnelshift_r(w1, w2, shb);
    uint32_t v = squeeze8_pairs(lo) | (squeeze8_pairs(hi) << 16);    // code j of the group at bits 2j
    if (cnt < 16u) v &= (1u << (2u * cnt)) - 1u;
    v = __brev(v);
    v = ((v & 0x55555555u) << 1) | ((v >> 1) & 0x55555555u);         // code j at bits 2 (15 - j)
    const uint32_t buffer = v >> (2u * (16u - cnt));
    key += (buffer >> 16);
    uint32_t tmp = ((buffer & 0xFFFFu) << 11) ^ key;
    key = (key << 16) ^ tmp;
    key += key >> 11;
    w0 = w2;
    if (sh == 0u && i * 16u + 16u < glen) w0 = wp[2 * i + 2];
  }
  key ^= key << 3;
  key += key >> 5;
  key ^= key << 4;
  key += key >> 17;
  key ^= key << 25;
  key += key >> 6;
  return key;
}

// one thread per read strand: DENSE sw_vector task lists (one per genome orientation) of its eligible hits.
// A strand's tasks stay together and in list order, so the two tasks a thread of sw_vector_kernel packs into
// its 16-bit lanes come from the same read (same length) whenever possible; out[t] names the hit slot the
// score goes to.
__global__ void build_vec_tasks_kernel(const TaskBuildParams P) {
  const uint32_t rs = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = rs < 2u * (uint32_t)P.n_reads;
  const int lane = threadIdx.x & 31;
  const uint2 rg = live ? P.rs_range[rs] : make_uint2(0u, 0u);
  const int r = (int)(rs >> 1), st = (int)(rs & 1u);
  const int rl = live ? P.read_len[r] : 0;
  const bool cs = P.M.colour_space != 0;
  // orientation used by pass 1: letter space always scores read strand st on the forward genome;
  // colour space scores the forward read and flips strand-1 windows onto the rc genome.
  const int in_st = live ? P.M.rev_mate[r & 1] : 0;   // re->input_strand (mapping.c:1303)
  const int ori = (cs && st != in_st) ? 1 : 0;
  uint32_t n_elig = 0;
  unsigned long long elig_cells = 0;
  for (uint32_t k = 0; k < rg.y; k++) {
    const DevHit h = P.hits[rg.x + k];
    if (h.matches >= P.M.min_matches) {
      n_elig++;
      elig_cells += (unsigned long long)h.w_len * (unsigned long long)rl;
    }
  }
  // dense slots: one atomic per warp and orientation (millions of per-strand atomics on one counter serialise),
  // the strands of a warp in lane order behind it
  uint32_t t = 0;
#pragma unroll
  for (int o = 0; o < 2; o++) {
    const uint32_t mine = ori == o ? n_elig : 0u;
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t base = 0;
    if (lane == 31 && total) base = atomicAdd(&P.task_stats[4 + o], total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (ori == o) t = base + incl - mine;
    if (o == 0) {   // statistics of the whole warp
      uint32_t ne = n_elig;
      unsigned long long ec = elig_cells;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        ne += __shfl_xor_sync(0xffffffffu, ne, d);
        ec += __shfl_xor_sync(0xffffffffu, ec, d);
      }
      if (lane == 0 && ne) {
        atomicAdd(&P.task_stats[0], ne);
        atomicAdd((unsigned long long *)&P.task_stats[2], ec);
      }
    }
  }
  for (uint32_t k = 0; k < rg.y; k++) {
    const uint32_t hi = rg.x + k;
    const DevHit h = P.hits[hi];
    if (h.matches < P.M.min_matches) continue;
    const uint32_t coff = P.G.contig_off[h.cn];
    const uint32_t g = ori ? coff + (P.G.contig_len[h.cn] - h.g_off - (uint32_t)h.w_len) : coff + h.g_off;
    P.goff[ori][t] = g;
    P.glen[ori][t] = h.w_len;
    P.ridx[ori][t] = cs ? (int32_t)(2 * r + in_st) : (int32_t)rs;
    P.rlen[ori][t] = rl;
    if (cs) P.initbp_out[ori][t] = P.initbp[r];
    P.out[ori][t] = hi;
    t++;
  }
}

// f1 cache slot (f1-wrapper.h:97-134: hash_genome_window % 2^20) of every dense task's window, a thread per task;
// the slots of the hits without a task stay 0xffffffff (the caller's memset)
__global__ void window_slots_kernel(const uint32_t *genome, const uint32_t *goff, const int32_t *glen, const uint32_t *out,
                                    uint32_t n_tasks, uint32_t *slot) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tasks) return;
  slot[out[t]] = hash_genome_window_dev(genome, goff[t], (uint32_t)glen[t]) % 1048576u;
}

__global__ void window_hash_kernel(const uint32_t *genome, const uint32_t *goff, const int32_t *glen, uint32_t n, uint32_t *hash) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) hash[t] = hash_genome_window_dev(genome, goff[t], (uint32_t)glen[t]);
}


// one thread per read: read_pass1 (both strands), mapping.c:1261-1366.  Serves the first pass of unpaired
// reads, the only_paired pass of read pairs (pair_min != nullptr) and the half-paired second pass (saved
// flags set, hits that already carry a positive score keep it, :1296).
__global__ void pass1_replay_kernel(const Pass1Params P) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P.n_reads) return;
  const MapParamsDev &M = P.M;
  const int rl = P.read_len[r];
  const int window_len = (int)(unsigned short)abs_or_pct_d(M.window_len, M.window_len_frac, (double)rl);
  const unsigned int ovl = (unsigned int)abs_or_pct_d(M.overlap_thr, M.overlap_frac, (double)window_len);
  const bool cs = M.colour_space != 0;
  uint32_t calls = 0, bypassed = 0;
  unsigned long long cells = 0;
  for (int st = 0; st < 2; st++) {
    const uint2 rg = P.rs_range[2 * r + st];
    const int ori = (cs && st != M.rev_mate[r & 1]) ? 1 : 0;
    int last_good_cn = -1;
    unsigned int last_good_g_off = 0;
    for (uint32_t k = 0; k < rg.y; k++) {
      const uint32_t hi = rg.x + k;
      DevHit h = P.hits[hi];
      P.writer[hi] = 0;
      if (P.pair_min && P.pair_min[hi] < 0) continue;  // only_paired (:1272-1274)
      if (h.matches < M.min_matches) continue;
      if (P.saved && P.saved[hi]) {  // :1281-1285
        last_good_cn = h.cn;
        last_good_g_off = h.g_off;
        continue;
      }
      // window overlap with the last good window (:1287-1293): llint + unsigned  <=  unsigned + int
      if (last_good_cn >= 0 && h.cn == last_good_cn &&
          (long long)h.g_off + (long long)ovl <= (long long)(unsigned int)(last_good_g_off + (unsigned int)window_len)) {
        P.hits[hi].score_vector = 0;
        P.hits[hi].pct_vector = 0;
        continue;
      }
      if (h.score_vector > 0) continue;  // :1296, only possible in a second pass
      int score = P.vtrue[ori][hi];
      bool hit_in_cache = false;
      if (M.hash_filter_calls) {
        const uint32_t sl = P.slot[hi];
        for (uint32_t q = 0; q < k; q++) {
          const uint32_t hq = rg.x + q;
          if (P.writer[hq] && P.slot[hq] == sl) {
            score = P.vtrue[ori][hq];
            hit_in_cache = true;
            break;
          }
        }
        if (!hit_in_cache) P.writer[hi] = 1;
      }
      if (hit_in_cache) {
        bypassed++;
      } else {
        calls++;
        if (!M.gapless) cells += (unsigned long long)h.w_len * (unsigned long long)rl;
      }
      const int pct = (1000 * 100 * score) / h.score_max;
      P.hits[hi].score_vector = score;
      P.hits[hi].pct_vector = pct;
      if (score >= (int)abs_or_pct_d(M.vect_thr, M.vect_frac, (double)h.score_max)) {
        last_good_cn = h.cn;
        last_good_g_off = h.g_off;
      }
    }
  }
  if (calls) atomicAdd(&P.stats[4], calls);
  if (bypassed) atomicAdd(&P.stats[5], bypassed);
  if (cells) atomicAdd((unsigned long long *)&P.stats[6], cells);
}

// one thread per read: read_get_vector_hits (:1376-1411): min-heap of capacity num_tmp_outputs keyed on pass1_key
__global__ void select_unpaired_kernel(const Pass1Params P) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P.n_reads) return;
  const MapParamsDev &M = P.M;
  const bool absolute = M.vect_thr < 0;
  int32_t *a = P.sel + (size_t)r * M.num_tmp_outputs;
  int load = 0;
#define P1KEY(slot_) (absolute ? P.hits[slot_].score_vector : P.hits[slot_].pct_vector)
  for (int st = 0; st < 2; st++) {
    const uint2 rg = P.rs_range[2 * r + st];
    for (uint32_t k = 0; k < rg.y; k++) {
      const int32_t hi = (int32_t)(rg.x + k);
      if (P.saved && P.saved[hi]) continue;
      const DevHit h = P.hits[hi];
      if (h.score_vector >= (int)abs_or_pct_d(M.vect_thr, M.vect_frac, (double)h.score_max) &&
          (load < M.num_tmp_outputs || (absolute ? h.score_vector : h.pct_vector) > P1KEY(a[0]))) {
        const int key = absolute ? h.score_vector : h.pct_vector;
        if (load < M.num_tmp_outputs) {  // extheap insert + percolate_up
          a[load] = hi;
          load++;
          int node = load, parent = node / 2;
          while (node > 1 && key < P1KEY(a[parent - 1])) {
            int32_t tmp = a[parent - 1];
            a[parent - 1] = a[node - 1];
            a[node - 1] = tmp;
            node = parent;
            parent = node / 2;
          }
        } else {  // replace_min + percolate_down
          a[0] = hi;
          int node = 1;
          for (;;) {
            int left = node * 2, right = left + 1, mn = node;
            if (left <= load && P1KEY(a[left - 1]) < P1KEY(a[node - 1])) mn = left;
            if (right <= load && P1KEY(a[right - 1]) < P1KEY(a[mn - 1])) mn = right;
            if (mn == node) break;
            int32_t tmp = a[mn - 1];
            a[mn - 1] = a[node - 1];
            a[node - 1] = tmp;
            node = mn;
          }
        }
      }
    }
  }
#undef P1KEY
  P.n_sel[r] = load;
}

// one thread per dense task: sw_gapless (common/sw-gapless.c:57-117) -- best ungapped segment on the diagonal
// through the hit's anchor, walked over the WHOLE contig/read overlap as f1_run does (f1-wrapper.h:121-124: genome =
// contig, glen = contig length, g_idx = g_off + anchor.x, r_idx = anchor.y).  Colour space (mapping.c:1297-1318):
// strand-1 hits are reversed onto the reverse-complement arrays (reverse_hit :254-263), the read is the input-strand
// one, and the first colour of the read is forced through the LETTER genome and the initial base (:83-93).
__global__ void sw_gapless_kernel(const GaplessParams P) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.n_tasks) return;
  const uint32_t hi = P.out[t];
  const DevHit h = P.hits[hi];
  const uint32_t *read = P.reads + (size_t)P.ridx[t] * P.stride;
  const int rlen = P.rlen[t];
  const uint64_t coff = P.G.contig_off[h.cn];
  const int glen = (int)P.G.contig_len[h.cn];
  int g_off = (int)h.g_off, ax = h.ax, ay = h.ay;
  if (P.ori) {
    g_off = glen - (int)h.g_off - h.w_len;
    ax = -h.ax + (h.w_len - 1) - (h.alen - 1) - (h.awidth - 1);
    ay = -h.ay + (rlen - 1) - (h.alen - 1) + (h.awidth - 1);
  }
  const uint32_t *genome = !P.cs ? P.G.ls : P.ori ? P.G.cs_rc : P.G.cs;
  const int g_idx = g_off + ax, r_idx = ay;
  int g = g_idx < r_idx ? 0 : g_idx - r_idx;
  int r = g_idx < r_idx ? r_idx - g_idx : 0;
  int score = 0;
  if (P.cs && r == 0) {  // forcefully match the first colour of the read
    const uint32_t *genome_ls = P.ori ? P.G.ls_rc : P.G.ls;
    const uint32_t letter = extract4(genome_ls, coff + (uint64_t)g);
    const int ib = P.initbp[t];
    const uint32_t real_colour = (letter > 3u || (uint32_t)ib > 3u) ? 15u : (letter ^ (uint32_t)ib);  // lstocs
    if (real_colour == extract4(read, 0)) score = P.match;
    r++;
    g++;
  }
  int max_score = score;
  while (g < glen && r < rlen) {
    score += (extract4(genome, coff + (uint64_t)g) == extract4(read, (uint64_t)r)) ? P.match : P.mismatch;
    if (score > max_score) max_score = score;
    g++;
    r++;
    if (score < 0) score = 0;
  }
  P.scores[hi] = max_score;
}

int launch_sw_gapless(shrimp_gpu_ctx *ctx, const GaplessParams &P) {
  if (P.n_tasks == 0) return SHRIMP_OK;
  sw_gapless_kernel<<<(P.n_tasks + 127) / 128, 128, 0, ctx->stream>>>(P);
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_VECTOR);
  return SHRIMP_OK;
}

int launch_build_vec_tasks(shrimp_gpu_ctx *ctx, const TaskBuildParams &P) {
  const unsigned n = 2u * (unsigned)P.n_reads;
  build_vec_tasks_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(P);
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_PASS1);
  return SHRIMP_OK;
}

int launch_window_slots(shrimp_gpu_ctx *ctx, const uint32_t *genome, const uint32_t *goff, const int32_t *glen,
                        const uint32_t *out, uint32_t n_tasks, uint32_t *slot) {
  if (n_tasks == 0) return SHRIMP_OK;
  window_slots_kernel<<<(n_tasks + 255) / 256, 256, 0, ctx->stream>>>(genome, goff, glen, out, n_tasks, slot);
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_PASS1);
  return SHRIMP_OK;
}

int launch_pass1_replay(shrimp_gpu_ctx *ctx, const Pass1Params &P) {
  pass1_replay_kernel<<<(P.n_reads + 127) / 128, 128, 0, ctx->stream>>>(P);
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_PASS1);
  return SHRIMP_OK;
}

int launch_select_unpaired(shrimp_gpu_ctx *ctx, const Pass1Params &P) {
  select_unpaired_kernel<<<(P.n_reads + 127) / 128, 128, 0, ctx->stream>>>(P);
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_PASS1);
  return SHRIMP_OK;
}

// sw_gapless (common/sw-gapless.c:57-117), one thread per independent task of the batch entry below: the best
// ungapped segment on the diagonal through (g_idx, r_idx) of a `glen`-long genome piece that starts at nibble goff.
__global__ void sw_gapless_batch_kernel(const uint32_t *genome, const uint32_t *genome_ls, const uint32_t *reads,
                                        int stride, const uint32_t *goff, const int32_t *glen, const int32_t *ridx,
                                        const int32_t *rlen, const int32_t *g_idx, const int32_t *r_idx,
                                        const int8_t *initbp, uint32_t n_tasks, int match, int mismatch,
                                        int32_t *scores) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tasks) return;
  const uint32_t *read = reads + (size_t)ridx[t] * stride;
  const uint64_t g0 = goff[t];
  const int gl = glen[t], rl = rlen[t];
  int g = g_idx[t] < r_idx[t] ? 0 : g_idx[t] - r_idx[t];
  int r = g_idx[t] < r_idx[t] ? r_idx[t] - g_idx[t] : 0;
  int score = 0;
  if (genome_ls != nullptr && r == 0) {  // forcefully match the first colour of the read (:83-93)
    const uint32_t letter = extract4(genome_ls, g0 + (uint64_t)g);
    const int ib = initbp[t];
    const uint32_t real_colour = (letter > 3u || (uint32_t)ib > 3u) ? 15u : (letter ^ (uint32_t)ib);  // lstocs
    if (real_colour == extract4(read, 0)) score = match;
    r++;
    g++;
  }
  int max_score = score;
  while (g < gl && r < rl) {
    score += (extract4(genome, g0 + (uint64_t)g) == extract4(read, (uint64_t)r)) ? match : mismatch;
    if (score > max_score) max_score = score;
    g++;
    r++;
    if (score < 0) score = 0;
  }
  scores[t] = max_score;
}

}  // namespace shrimp

// sw_gapless (sw-gapless.c:57) for a batch of independent tasks, host buffers in, scores out.
extern "C" int shrimp_gpu_sw_gapless_batch(shrimp_gpu_ctx *ctx, const uint32_t *genome, size_t genome_words,
                                           const uint32_t *genome_ls, const uint32_t *reads, int read_stride_words,
                                           int n_reads, int n_tasks, const uint32_t *goff, const int32_t *glen,
                                           const int32_t *read_idx, const int32_t *rlen, const int32_t *g_idx,
                                           const int32_t *r_idx, const int8_t *initbp, int32_t *scores_out) {
  using namespace shrimp;
  if (!ctx || !genome || !reads || !goff || !glen || !read_idx || !rlen || !g_idx || !r_idx || !scores_out ||
      n_tasks < 0 || n_reads <= 0 || read_stride_words <= 0 || (genome_ls && !initbp)) {
    set_error("shrimp_gpu_sw_gapless_batch: invalid argument");
    return SHRIMP_E_ARG;
  }
  if (!ctx->sw.valid) {
    set_error("shrimp_gpu_sw_gapless_batch: shrimp_gpu_sw_setup() has not been called");
    return SHRIMP_E_STATE;
  }
  if (n_tasks == 0) return SHRIMP_OK;
  for (int i = 0; i < n_tasks; i++)
    if (glen[i] < 0 || rlen[i] < 0 || read_idx[i] < 0 || read_idx[i] >= n_reads || g_idx[i] < 0 || r_idx[i] < 0 ||
        (uint64_t)goff[i] + (uint64_t)glen[i] > (uint64_t)genome_words * 8 || rlen[i] > read_stride_words * 8) {
      set_error("shrimp_gpu_sw_gapless_batch: task %d out of range", i);
      return SHRIMP_E_ARG;
    }
  SH_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t gbytes = genome_words * sizeof(uint32_t);
  SH_TRY(ctx->d_genome.ensure(gbytes + 16));
  SH_CUDA(cudaMemsetAsync((char *)ctx->d_genome.p + gbytes, 0, 16, st));
  SH_CUDA(cudaMemcpyAsync(ctx->d_genome.p, genome, gbytes, cudaMemcpyHostToDevice, st));
  if (genome_ls) {
    SH_TRY(ctx->d_genome_ls.ensure(gbytes + 16));
    SH_CUDA(cudaMemsetAsync((char *)ctx->d_genome_ls.p + gbytes, 0, 16, st));
    SH_CUDA(cudaMemcpyAsync(ctx->d_genome_ls.p, genome_ls, gbytes, cudaMemcpyHostToDevice, st));
  }
  const size_t rbytes = (size_t)n_reads * read_stride_words * sizeof(uint32_t);
  SH_TRY(ctx->d_reads.ensure(rbytes));
  SH_CUDA(cudaMemcpyAsync(ctx->d_reads.p, reads, rbytes, cudaMemcpyHostToDevice, st));
  // task arrays packed in one buffer: goff | glen | ridx | rlen | g_idx | r_idx | initbp
  const size_t n = (size_t)n_tasks;
  SH_TRY(ctx->d_task.ensure(n * 4 * 6 + ((n + 3) & ~(size_t)3)));
  char *tb = (char *)ctx->d_task.p;
  const void *src[6] = {goff, glen, read_idx, rlen, g_idx, r_idx};
  for (int k = 0; k < 6; k++) SH_CUDA(cudaMemcpyAsync(tb + n * 4 * k, src[k], n * 4, cudaMemcpyHostToDevice, st));
  if (genome_ls) SH_CUDA(cudaMemcpyAsync(tb + n * 24, initbp, n, cudaMemcpyHostToDevice, st));
  SH_TRY(ctx->d_scores.ensure(n * 4));
  {
    ScopedStage ss(ctx, ST_VECTOR);
    sw_gapless_batch_kernel<<<(n_tasks + 127) / 128, 128, 0, st>>>(
        ctx->d_genome.as<uint32_t>(), genome_ls ? ctx->d_genome_ls.as<uint32_t>() : nullptr, ctx->d_reads.as<uint32_t>(),
        read_stride_words, (const uint32_t *)tb, (const int32_t *)(tb + n * 4), (const int32_t *)(tb + n * 8),
        (const int32_t *)(tb + n * 12), (const int32_t *)(tb + n * 16), (const int32_t *)(tb + n * 20),
        (const int8_t *)(tb + n * 24), (uint32_t)n_tasks, ctx->sw.match, ctx->sw.vec_mismatch, ctx->d_scores.as<int32_t>());
    SH_CUDA(cudaGetLastError());
    SH_LAUNCHED(ctx, ST_VECTOR);
  }
  SH_CUDA(cudaMemcpyAsync(scores_out, ctx->d_scores.p, n * 4, cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaStreamSynchronize(st));
  return SHRIMP_OK;
}

// Diagnostic: hash_genome_window (util.h:220-241) of n windows of a packed genome (8 codes per word) as the device
// computes it for the f1 window cache, for the test that compares it with the reference's.
extern "C" int shrimp_gpu_hash_windows(shrimp_gpu_ctx *ctx, const uint32_t *genome_words, uint64_t n_words,
                                       const uint32_t *goff, const int32_t *glen, int n, uint32_t *hash_out) {
  using namespace shrimp;
  if (!ctx || !genome_words || !goff || !glen || !hash_out || n < 0) {
    set_error("shrimp_gpu_hash_windows: invalid argument");
    return SHRIMP_E_ARG;
  }
  if (n == 0) return SHRIMP_OK;
  for (int t = 0; t < n; t++)
    if (glen[t] < 0 || (uint64_t)goff[t] + (uint64_t)glen[t] > n_words * 8) {
      set_error("shrimp_gpu_hash_windows: window %d outside the genome", t);
      return SHRIMP_E_RANGE;
    }
  DevBuf dg, do_, dl, dh;
  int rc = SHRIMP_OK;
  if ((rc = dg.ensure(n_words * 4)) == SHRIMP_OK && (rc = do_.ensure((size_t)n * 4)) == SHRIMP_OK &&
      (rc = dl.ensure((size_t)n * 4)) == SHRIMP_OK && (rc = dh.ensure((size_t)n * 4)) == SHRIMP_OK) {
    cudaMemcpyAsync(dg.p, genome_words, n_words * 4, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(do_.p, goff, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dl.p, glen, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream);
    window_hash_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(dg.as<uint32_t>(), do_.as<uint32_t>(), dl.as<int32_t>(),
                                                                  (uint32_t)n, dh.as<uint32_t>());
    ctx->launches++;
    cudaMemcpyAsync(hash_out, dh.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      set_error("shrimp_gpu_hash_windows: %s", cudaGetErrorString(e));
      rc = SHRIMP_E_CUDA;
    }
  }
  dg.release(); do_.release(); dl.release(); dh.release();
  return rc;
}
