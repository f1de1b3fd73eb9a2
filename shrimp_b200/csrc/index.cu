// Genome residency and the spaced-seed projection ("genome map") built and held in HBM.
//
// Replaces the data side of load_genome (gmapper/genome.c:1012-1182): the reverse-complement and
// colour-space arrays (util.c:541-598, fasta.c:586-612) are derived on the device from the packed
// forward contigs, and the projection loop (genome.c:1138-1166, one realloc per position) becomes
//   project (one thread per genome position and seed)  ->  stable radix sort by bucket id
//   ->  CSR offsets,
// which reproduces the content and order (ascending position) of every genomemap[sn][mapidx] list.
// The radix sort is cub::DeviceRadixSort (library code, one-time index build, not the hot path).
#include <cub/cub.cuh>
#include "genome.cuh"

namespace shrimp {

__constant__ uint8_t c_cmpl[16] = {3, 2, 1, 0, 0, 10, 9, 7, 8, 6, 5, 14, 13, 12, 11, 15};  // util.h:128-145

// last contig whose offset is <= p (get_contig_num, gmapper.h:373-405)
__device__ __forceinline__ int contig_of(const uint32_t *off, int n, uint32_t p) {
  int lo = 0, hi = n;  // off[lo] <= p < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= p) lo = mid; else hi = mid;
  }
  return lo;
}

// dst nibbles [dst_off, dst_off+len) <- src nibbles [0, len); one thread per destination word.
__global__ void append_contig_kernel(uint32_t *dst, uint64_t dst_off, const uint32_t *src, uint32_t len) {
  uint64_t w0 = dst_off >> 3;
  uint64_t w = w0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t wlast = (dst_off + len - 1) >> 3;
  if (w > wlast) return;
  uint32_t keep = dst[w], val = 0, m = 0;
  for (int n = 0; n < 8; n++) {
    uint64_t p = w * 8 + n;
    if (p >= dst_off && p < dst_off + len) {
      val |= extract4(src, p - dst_off) << (4 * n);
      m |= 0xfu << (4 * n);
    }
  }
  dst[w] = (keep & ~m) | val;
}

// rc[off+i] = cmpl(ls[off+len-1-i]); colour arrays: c[i] = lstocs(prev letter or T, letter i)
__global__ void derive_genome_kernel(const uint32_t *ls, uint32_t *rc, uint32_t *cs, uint32_t *cs_rc,
                                     const uint32_t *off, const uint32_t *glen, int num_contigs, uint64_t total,
                                     int pass) {
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w * 8 >= total) return;
  uint32_t out = 0, out2 = 0;
  for (int n = 0; n < 8; n++) {
    uint64_t p = w * 8 + n;
    if (p >= total) break;
    int cn = contig_of(off, num_contigs, (uint32_t)p);
    uint32_t i = (uint32_t)p - off[cn];
    if (pass == 0) {
      out |= (uint32_t)c_cmpl[extract4(ls, (uint64_t)off[cn] + glen[cn] - 1 - i)] << (4 * n);
    } else {
      uint32_t a = extract4(ls, p), b = extract4(rc, p);
      uint32_t pa = i == 0 ? 3u : extract4(ls, p - 1), pb = i == 0 ? 3u : extract4(rc, p - 1);
      out |= ((a > 3u || pa > 3u) ? 15u : (a ^ pa)) << (4 * n);
      out2 |= ((b > 3u || pb > 3u) ? 15u : (b ^ pb)) << (4 * n);
    }
  }
  if (pass == 0) {
    rc[w] = out;
  } else {
    cs[w] = out;
    cs_rc[w] = out2;
  }
}

// key[p] = bucket id of the k-mer starting at global position p, or `invalid` when the k-mer
// crosses the contig end or contains N/X (genome.c:1147-1154).
__global__ void project_kernel(const uint32_t *seq, const uint32_t *off, const uint32_t *glen, int num_contigs,
                               uint64_t total, SeedTable S, int sn, uint32_t invalid, uint32_t *key, uint32_t *val) {
  uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int span = S.span[sn];
  int cn = contig_of(off, num_contigs, (uint32_t)p);
  uint32_t k = invalid;
  if ((uint64_t)p + span <= (uint64_t)off[cn] + glen[cn]) {
    bool ok = true;
    for (int i = 0; i < span; i++) ok &= (extract4(seq, p + i) != 15u);
    if (ok) k = kmer_to_mapidx(S, sn, seq, p);
  }
  key[p] = k;
  val[p] = (uint32_t)p;
}

// offs[m] = first sorted index whose key is >= m, for m in [0, nbuckets]; invalid k-mers carry the
// key `nbuckets`, so offs[nbuckets] is the number of valid positions.
__global__ void csr_offsets_kernel(const uint32_t *key, uint64_t n, uint32_t nbuckets, uint32_t *offs) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  long long cur = i < n ? (long long)min(key[i], nbuckets) : (long long)nbuckets;
  long long prev = i == 0 ? -1ll : (long long)min(key[i - 1], nbuckets);
  for (long long m = prev + 1; m <= cur; m++) offs[m] = (uint32_t)i;
}

__global__ void bucket_heads_kernel(const uint32_t *__restrict__ offs, const uint32_t *__restrict__ pos, uint32_t nbuckets,
                                    uint32_t *__restrict__ heads) {
  // a 16-lane group per bucket: one coalesced 64-byte store
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t m = t >> 4;
  const uint32_t j = (uint32_t)t & 15u;
  if (m >= nbuckets) return;
  const uint32_t b = offs[m], len = offs[m + 1] - b;
  uint32_t w = 0;
  if (j == 0) w = len;
  else if (len <= SHRIMP_HEAD_WORDS - 1) w = j <= len ? pos[b + j - 1] : 0u;
  else if (j == 1) w = b;
  heads[m * SHRIMP_HEAD_WORDS + j] = w;
}

int build_bucket_heads(shrimp_gpu_ctx *ctx, DeviceGenome *g) {
  const char *env = getenv("SHRIMP_BUCKET_HEADS");   // 0 = never, 1 = always (tests), default: sparse projections only
  for (int sn = 0; sn < g->seeds.n_seeds; sn++) {
    g->head_off[sn] = 0;
    const uint64_t nb = g->nbuckets[sn], total = g->total[sn];
    const uint64_t off = (total + 1 + 63) & ~(uint64_t)63;
    bool want = total <= 8 * nb;   // a mean of at most eight entries per bucket: most lists fit a head
    if (env) want = atoi(env) != 0;
    if (!want || off + nb * SHRIMP_HEAD_WORDS >= 0xffffffffull) continue;
    DevBuf both;
    SH_TRY(both.ensure((off + nb * SHRIMP_HEAD_WORDS) * 4));
    SH_CUDA(cudaMemcpyAsync(both.p, g->d_pos[sn].p, (size_t)total * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    const uint64_t threads = nb * SHRIMP_HEAD_WORDS;
    bucket_heads_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(
        g->d_offs[sn].as<uint32_t>(), both.as<uint32_t>(), (uint32_t)nb, both.as<uint32_t>() + off);
    SH_CUDA(cudaGetLastError());
    SH_LAUNCHED(ctx, ST_INDEX);
    SH_CUDA(cudaStreamSynchronize(ctx->stream));
    g->d_pos[sn].release();
    g->d_pos[sn] = both;
    both.p = nullptr;
    both.cap = 0;
    g->head_off[sn] = (uint32_t)off;
  }
  return SHRIMP_OK;
}

void free_genome(shrimp_gpu_ctx *ctx) {
  DeviceGenome *g = genome_of(ctx);
  if (!g) return;
  if (ctx->genome_borrowed) {
    ctx->genome = nullptr;
    ctx->genome_borrowed = false;
    return;
  }
  g->d_off.release();
  g->d_len.release();
  g->d_ls.release();
  g->d_ls_rc.release();
  g->d_cs.release();
  g->d_cs_rc.release();
  for (int i = 0; i < SHRIMP_MAX_SEEDS; i++) {
    g->d_offs[i].release();
    g->d_pos[i].release();
  }
  delete g;
  ctx->genome = nullptr;
}

}  // namespace shrimp

using namespace shrimp;

// Replaces the genome arrays load_genome builds (genome.c:1092-1124): takes the reference's own
// globals -- genome_contigs[] (packed letters), genome_len[], num_contigs -- and derives the
// reverse complement and (colour space) both colour arrays in HBM.
extern "C" int shrimp_gpu_genome_load(shrimp_gpu_ctx *ctx, int num_contigs, const uint32_t *const *genome_contigs,
                                      const uint32_t *genome_len, int colour_space) {
  if (!ctx || num_contigs <= 0 || !genome_contigs || !genome_len) {
    set_error("shrimp_gpu_genome_load: invalid argument");
    return SHRIMP_E_ARG;
  }
  SH_CUDA(cudaSetDevice(ctx->device));
  free_genome(ctx);
  DeviceGenome *g = new DeviceGenome();
  ctx->genome = g;
  g->num_contigs = num_contigs;
  g->colour_space = colour_space ? 1 : 0;
  g->h_off.resize(num_contigs + 1);
  g->h_len.assign(genome_len, genome_len + num_contigs);
  uint64_t tot = 0;
  uint32_t max_len = 0;
  for (int c = 0; c < num_contigs; c++) {
    if (genome_len[c] == 0 || !genome_contigs[c]) {
      set_error("shrimp_gpu_genome_load: contig %d is empty", c);
      return SHRIMP_E_ARG;
    }
    g->h_off[c] = (uint32_t)tot;
    tot += genome_len[c];
    if (genome_len[c] > max_len) max_len = genome_len[c];
  }
  if (tot >= 0xffffffffull) {  // global coordinates are uint32_t in the reference too (gmapper.h:264)
    set_error("shrimp_gpu_genome_load: total genome length %llu does not fit 32-bit coordinates",
              (unsigned long long)tot);
    return SHRIMP_E_RANGE;
  }
  g->h_off[num_contigs] = (uint32_t)tot;
  g->total_len = tot;
  g->words = (size_t)((tot + 7) / 8) + 4;
  cudaStream_t st = ctx->stream;
  const size_t bytes = g->words * 4;
  SH_TRY(g->d_off.ensure((size_t)(num_contigs + 1) * 4));
  SH_TRY(g->d_len.ensure((size_t)num_contigs * 4));
  SH_CUDA(cudaMemcpyAsync(g->d_off.p, g->h_off.data(), (size_t)(num_contigs + 1) * 4, cudaMemcpyHostToDevice, st));
  SH_CUDA(cudaMemcpyAsync(g->d_len.p, g->h_len.data(), (size_t)num_contigs * 4, cudaMemcpyHostToDevice, st));
  SH_TRY(g->d_ls.ensure(bytes));
  SH_TRY(g->d_ls_rc.ensure(bytes));
  SH_CUDA(cudaMemsetAsync(g->d_ls.p, 0, bytes, st));
  SH_CUDA(cudaMemsetAsync(g->d_ls_rc.p, 0, bytes, st));
  DevBuf stage;
  SH_TRY(stage.ensure(((size_t)max_len + 7) / 8 * 4 + 16));
  ScopedStage ss(ctx, ST_INDEX);
  for (int c = 0; c < num_contigs; c++) {
    size_t cw = ((size_t)genome_len[c] + 7) / 8;
    SH_CUDA(cudaMemcpyAsync(stage.p, genome_contigs[c], cw * 4, cudaMemcpyHostToDevice, st));
    uint64_t nwords = ((g->h_off[c] + (uint64_t)genome_len[c] - 1) >> 3) - ((uint64_t)g->h_off[c] >> 3) + 1;
    append_contig_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(g->d_ls.as<uint32_t>(), g->h_off[c],
                                                                              stage.as<uint32_t>(), genome_len[c]);
    SH_CUDA(cudaGetLastError());
    SH_LAUNCHED(ctx, ST_INDEX);
    SH_CUDA(cudaStreamSynchronize(st));  // staging buffer is reused
  }
  stage.release();
  const unsigned nb = (unsigned)(((tot + 7) / 8 + 255) / 256);
  derive_genome_kernel<<<nb, 256, 0, st>>>(g->d_ls.as<uint32_t>(), g->d_ls_rc.as<uint32_t>(), nullptr, nullptr,
                                           g->d_off.as<uint32_t>(), g->d_len.as<uint32_t>(), num_contigs, tot, 0);
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_INDEX);
  if (g->colour_space) {
    SH_TRY(g->d_cs.ensure(bytes));
    SH_TRY(g->d_cs_rc.ensure(bytes));
    SH_CUDA(cudaMemsetAsync(g->d_cs.p, 0, bytes, st));
    SH_CUDA(cudaMemsetAsync(g->d_cs_rc.p, 0, bytes, st));
    derive_genome_kernel<<<nb, 256, 0, st>>>(g->d_ls.as<uint32_t>(), g->d_ls_rc.as<uint32_t>(),
                                             g->d_cs.as<uint32_t>(), g->d_cs_rc.as<uint32_t>(),
                                             g->d_off.as<uint32_t>(), g->d_len.as<uint32_t>(), num_contigs, tot, 1);
    SH_CUDA(cudaGetLastError());
    SH_LAUNCHED(ctx, ST_INDEX);
  }
  SH_CUDA(cudaStreamSynchronize(st));
  return SHRIMP_OK;
}

// Replaces the projection loop of load_genome (genome.c:1138-1166) for the seeds of
// add_spaced_seed (seeds.c:9-42): masks[sn] bit 0 = rightmost seed character.
namespace shrimp {
// The seed table the kernels take by value, from the reference's seed_type fields (gmapper-definitions.h:59-63).
int seed_table_init(SeedTable &S, int n_seeds, const uint64_t *masks, const int32_t *spans, const int32_t *weights,
                    int hflag, const char *who) {
  if (n_seeds <= 0 || n_seeds > SHRIMP_MAX_SEEDS || !masks || !spans || !weights) {
    set_error("%s: invalid seed table", who);
    return SHRIMP_E_ARG;
  }
  S = SeedTable{};
  S.n_seeds = n_seeds;
  S.hflag = hflag ? 1 : 0;
  S.min_span = 64;
  for (int sn = 0; sn < n_seeds; sn++) {
    if (spans[sn] < 1 || spans[sn] > 64 || weights[sn] < 1 || (!hflag && weights[sn] > 14)) {
      // MAX_SEED_SPAN / MAX_SEED_WEIGHT, gmapper-definitions.h:52-56
      set_error("%s: seed %d has span %d weight %d (max 64 / 14 without -H)", who, sn, spans[sn], weights[sn]);
      return SHRIMP_E_ARG;
    }
    S.mask[sn] = masks[sn];
    S.span[sn] = spans[sn];
    S.weight[sn] = weights[sn];
    if (spans[sn] > S.max_span) S.max_span = spans[sn];
    if (spans[sn] < S.min_span) S.min_span = spans[sn];
    // runs of care positions in k-mer order: k-mer base j pairs with mask bit span-1-j (seeds.c:9-42), and
    // the first care base lands in the lowest two bits of the bucket id (kmer_to_mapidx_orig, gmapper.h:349-368)
    S.n_runs[sn] = 0;
    if (!hflag && spans[sn] <= 32) {
      int nr = 0, rank = 0;
      bool ok = true;
      for (int j = 0; j < spans[sn];) {
        if (!((masks[sn] >> (spans[sn] - 1 - j)) & 1ull)) {
          j++;
          continue;
        }
        int e = j;
        while (e < spans[sn] && ((masks[sn] >> (spans[sn] - 1 - e)) & 1ull)) e++;
        if (nr == SHRIMP_MAX_RUNS) {
          ok = false;
          break;
        }
        S.run_src[sn][nr] = (unsigned char)j;
        S.run_len[sn][nr] = (unsigned char)(e - j);
        S.run_dst[sn][nr] = (unsigned char)rank;
        rank += e - j;
        nr++;
        j = e;
      }
      if (ok) S.n_runs[sn] = (unsigned char)nr;
    }
  }
  return SHRIMP_OK;
}
}  // namespace shrimp

extern "C" int shrimp_gpu_index_build(shrimp_gpu_ctx *ctx, int n_seeds, const uint64_t *masks, const int32_t *spans,
                                      const int32_t *weights, int hflag) {
  DeviceGenome *g = ctx ? genome_of(ctx) : nullptr;
  if (!g) {
    set_error("shrimp_gpu_index_build: load the genome first");
    return SHRIMP_E_STATE;
  }
  SeedTable S{};
  SH_TRY(seed_table_init(S, n_seeds, masks, spans, weights, hflag, "shrimp_gpu_index_build"));
  SH_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint64_t L = g->total_len;
  const uint32_t *seq = g->colour_space ? g->d_cs.as<uint32_t>() : g->d_ls.as<uint32_t>();
  DevBuf key_a, key_b, val_a, val_b, tmp;
  SH_TRY(key_a.ensure(L * 4));
  SH_TRY(key_b.ensure(L * 4));
  SH_TRY(val_a.ensure(L * 4));
  SH_TRY(val_b.ensure(L * 4));
  ScopedStage ss(ctx, ST_INDEX);
  int rc = SHRIMP_OK;
  for (int sn = 0; sn < n_seeds && rc == SHRIMP_OK; sn++) {
    const int bits = 2 * (hflag ? 12 : weights[sn]);
    const uint32_t nbuckets = 1u << bits;
    g->nbuckets[sn] = nbuckets;
    project_kernel<<<(unsigned)((L + 255) / 256), 256, 0, st>>>(seq, g->d_off.as<uint32_t>(), g->d_len.as<uint32_t>(),
                                                                g->num_contigs, L, S, sn, nbuckets,
                                                                key_a.as<uint32_t>(), val_a.as<uint32_t>());
    SH_CUDA(cudaGetLastError());
    SH_LAUNCHED(ctx, ST_INDEX);
    size_t tmp_bytes = 0;
    cub::DoubleBuffer<uint32_t> dk(key_a.as<uint32_t>(), key_b.as<uint32_t>());
    cub::DoubleBuffer<uint32_t> dv(val_a.as<uint32_t>(), val_b.as<uint32_t>());
    SH_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int64_t)L, 0, bits + 1, st));
    SH_TRY(tmp.ensure(tmp_bytes));
    SH_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dk, dv, (int64_t)L, 0, bits + 1, st));
    ctx->launches += 4;
    SH_TRY(g->d_offs[sn].ensure(((size_t)nbuckets + 2) * 4));
    csr_offsets_kernel<<<(unsigned)((L + 1 + 255) / 256), 256, 0, st>>>(dk.Current(), L, nbuckets,
                                                                          g->d_offs[sn].as<uint32_t>());
    SH_CUDA(cudaGetLastError());
    SH_LAUNCHED(ctx, ST_INDEX);
    uint32_t n_valid = 0;
    SH_CUDA(cudaMemcpyAsync(&n_valid, g->d_offs[sn].as<uint32_t>() + nbuckets, 4, cudaMemcpyDeviceToHost, st));
    SH_CUDA(cudaStreamSynchronize(st));
    g->total[sn] = n_valid;
    SH_TRY(g->d_pos[sn].ensure(((size_t)n_valid + 1) * 4));
    SH_CUDA(cudaMemcpyAsync(g->d_pos[sn].p, dv.Current(), (size_t)n_valid * 4, cudaMemcpyDeviceToDevice, st));
    SH_CUDA(cudaStreamSynchronize(st));
  }
  key_a.release();
  key_b.release();
  val_a.release();
  val_b.release();
  tmp.release();
  g->seeds = S;
  g->have_index = true;
  if (rc == SHRIMP_OK) rc = build_bucket_heads(ctx, g);
  return rc;
}

// Several host threads can drive one GPU, each with its own context (stream, chunk buffers, scoring set-up),
// the way gmapper's -N threads share the genome and the projection (gmapper.h:262-275 are process globals):
// dst borrows the genome and index resident in src.  src must outlive dst and must not reload meanwhile.
extern "C" int shrimp_gpu_share_genome(shrimp_gpu_ctx *dst, shrimp_gpu_ctx *src) {
  if (!dst || !src || dst == src || !genome_of(src)) {
    set_error("shrimp_gpu_share_genome: invalid argument / nothing resident in src");
    return SHRIMP_E_ARG;
  }
  if (dst->device != src->device) {
    set_error("shrimp_gpu_share_genome: contexts are on different devices");
    return SHRIMP_E_ARG;
  }
  free_genome(dst);
  dst->genome = src->genome;
  dst->genome_borrowed = true;
  return SHRIMP_OK;
}

// Copies the projection of seed sn back in the layout of the reference's -S files
// (genome.c:37-63): lens[4^W] (genomemap_len) and the concatenated position lists.
extern "C" int shrimp_gpu_index_export(shrimp_gpu_ctx *ctx, int sn, uint32_t *lens_out, uint32_t *pos_out,
                                       uint64_t *total_out) {
  DeviceGenome *g = ctx ? genome_of(ctx) : nullptr;
  if (!g || !g->have_index || sn < 0 || sn >= g->seeds.n_seeds) {
    set_error("shrimp_gpu_index_export: no index / bad seed number");
    return SHRIMP_E_STATE;
  }
  SH_CUDA(cudaSetDevice(ctx->device));
  if (total_out) *total_out = g->total[sn];
  if (lens_out) {
    std::vector<uint32_t> offs((size_t)g->nbuckets[sn] + 1);
    SH_CUDA(cudaMemcpy(offs.data(), g->d_offs[sn].p, offs.size() * 4, cudaMemcpyDeviceToHost));
    for (uint32_t m = 0; m < g->nbuckets[sn]; m++) lens_out[m] = offs[m + 1] - offs[m];
  }
  if (pos_out && g->total[sn]) SH_CUDA(cudaMemcpy(pos_out, g->d_pos[sn].p, g->total[sn] * 4, cudaMemcpyDeviceToHost));
  return SHRIMP_OK;
}

extern "C" int shrimp_gpu_index_nbuckets(shrimp_gpu_ctx *ctx, int sn, uint32_t *nbuckets, uint64_t *total) {
  DeviceGenome *g = ctx ? genome_of(ctx) : nullptr;
  if (!g || !g->have_index || sn < 0 || sn >= g->seeds.n_seeds) {
    set_error("shrimp_gpu_index_nbuckets: no index / bad seed number");
    return SHRIMP_E_STATE;
  }
  if (nbuckets) *nbuckets = g->nbuckets[sn];
  if (total) *total = g->total[sn];
  return SHRIMP_OK;
}

// Copies a genome orientation back (0 letters fwd, 1 letters rc, 2 colours fwd, 3 colours rc) in
// global packed coordinates -- used by the parity tests against the reference's arrays.
extern "C" int shrimp_gpu_genome_export(shrimp_gpu_ctx *ctx, int which, uint32_t *out_words, size_t n_words) {
  DeviceGenome *g = ctx ? genome_of(ctx) : nullptr;
  if (!g) {
    set_error("shrimp_gpu_genome_export: no genome");
    return SHRIMP_E_STATE;
  }
  DevBuf *b = which == 0 ? &g->d_ls : which == 1 ? &g->d_ls_rc : which == 2 ? &g->d_cs : &g->d_cs_rc;
  if (!b->p || n_words > g->words) {
    set_error("shrimp_gpu_genome_export: array %d not resident or size too large", which);
    return SHRIMP_E_ARG;
  }
  SH_CUDA(cudaSetDevice(ctx->device));
  SH_CUDA(cudaMemcpy(out_words, b->p, n_words * 4, cudaMemcpyDeviceToHost));
  return SHRIMP_OK;
}
