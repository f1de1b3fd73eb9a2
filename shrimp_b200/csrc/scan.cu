// Seed scan: spaced-seed projection of the reads, index lookup, region filter, anchor merge and
// candidate-window (hit) generation, everything between the HBM index and the hit list staged in shared
// memory.
//
// Replaces, per read strand (gmapper/mapping.c):
//   read_get_mapidxs_per_strand :37-70, read_get_region_counts :459-542,
//   advance_index_in_genomemap :646-805 (unpaired branch), read_get_anchor_list_per_strand :861-1006,
//   read_get_hit_list_per_strand :1025-1229.
//
// The reference marks 2 kb regions in a 4 MB table while walking every index list, then walks the lists
// again through a k-way heap merge.  Here (one warp per read strand; one CTA for the rare strands with
// very many index hits):
//   1. the read is recoded to 2 bits per base in shared memory; every k-mer is projected ONCE with a few
//      shift/mask operations per run of care positions of the seed (SeedTable::run_*), its bucket bounds are
//      fetched and kept in shared memory with a running prefix sum of the list lengths;
//   2. pass A streams all list entries, flattened over lanes by a search in the prefix sums so that long
//      lists coalesce and short ones balance, and marks a hashed region bitmap pair in shared memory
//      ("touched" / "touched twice") -- the reference's region_map (gmapper.h:284-294) without its 4 MB;
//   3. pass B streams the same entries again (L1/L2 hits) and keeps only those whose region (or region-1
//      within the overlap) is marked twice: a superset of the reference's survivors (hash collisions add
//      false positives, never lose one), a few dozen entries instead of hundreds;
//   4. the candidates are sorted by position (warp / CTA bitonic sort) -- the order the heap merge produces --
//      and "region has >= 2 hits" (RG_HAS_2) is decided EXACTLY as a neighbour test on the sorted array: a
//      region's catchment [r*2^11, (r+1)*2^11 + overlap) is contiguous, every witness of a true survivor is
//      itself a candidate, so the previous/next candidate tells;
//   5. equal positions on different read offsets would be popped in binary-heap order by the reference
//      (SURVEY hard part 3a): such strands (about 0.05 %) replay heap_uu (common/heap.h:43-113) exactly;
//   6. colinear collapse (:957-971) and the hit list run on the shared-memory anchors.
#include "stages.cuh"

namespace shrimp {

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

__device__ __forceinline__ int contig_of_dev(const uint32_t *off, int n, uint32_t p) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= p) lo = mid; else hi = mid;
  }
  return lo;
}

// anchor_join of two unit-width anchors (anchors.c:9-54); a[] = {x, y, length, width}
__device__ __forceinline__ void anchor_join2(long long x0, long long y0, int l0, int w0, long long x1, long long y1,
                                             int l1, int w1, int &ox, int &oy, int &olen, int &owidth) {
  long long nw0 = x0 + y0, sw0 = x0 - y0, ne0 = sw0 + 2 * (w0 - 1), se0 = nw0 + 2 * (l0 - 1);
  long long nw1 = x1 + y1, sw1 = x1 - y1, ne1 = sw1 + 2 * (w1 - 1), se1 = nw1 + 2 * (l1 - 1);
  long long nw_min = nw0 < nw1 ? nw0 : nw1, sw_min = sw0 < sw1 ? sw0 : sw1;
  long long ne_max = ne0 > ne1 ? ne0 : ne1, se_max = se0 > se1 ? se0 : se1;
  if ((nw_min + sw_min) % 2 != 0) nw_min--;
  long long dx = (nw_min + sw_min) / 2;
  ox = (int)dx;
  oy = (int)(nw_min - dx);
  if ((ne_max - sw_min) % 2 != 0) ne_max++;
  owidth = (int)((ne_max - sw_min) / 2 + 1);
  if ((se_max - nw_min) % 2 != 0) se_max++;
  olen = (int)((se_max - nw_min) / 2 + 1);
}



// ---- projection and lookup helpers ----------------------------------------------------------------
// 8 bases (4 bits each) -> 16 bits (2 bits each, low two bits of every code as KMER_TO_MAPIDX uses them)
__device__ __forceinline__ uint32_t squeeze8(uint32_t w) {
  w &= 0x33333333u;
  w = (w | (w >> 2)) & 0x0f0f0f0fu;
  w = (w | (w >> 4)) & 0x00ff00ffu;
  w = (w | (w >> 8)) & 0x0000ffffu;
  return w;
}

// bucket id of the k-mer of seed sn starting at base `start` of the 2-bit read r2 (padded with 2 zero words)
__device__ __forceinline__ uint32_t mapidx_fast(const SeedTable &S, int sn, const uint32_t *r2, int start) {
  const int o = 2 * start, w = o >> 5, sh = o & 31;
  unsigned long long x = ((unsigned long long)r2[w + 1] << 32) | r2[w];
  x >>= sh;
  if (sh) x |= (unsigned long long)r2[w + 2] << (64 - sh);
  uint32_t m = 0;
  const int nr = S.n_runs[sn];
  for (int q = 0; q < nr; q++) {
    const unsigned long long part = (x >> (2 * S.run_src[sn][q])) & ((1ull << (2 * S.run_len[sn][q])) - 1ull);
    m |= (uint32_t)(part << (2 * S.run_dst[sn][q]));
  }
  return m;
}

struct KmerTables {   // shared memory, per read strand
  uint32_t *kst;      // [K] list start in pos[sn]
  uint32_t *kpre;     // [K + 1] exclusive prefix sums of the (cut-off) list lengths
};

// entry t of the flattened lists -> (address of its position, k-mer slot sn * max_n_kmers + i)
__device__ __forceinline__ const uint32_t *flat_addr(const ScanParams &P, const KmerTables &T, int K, const int *kbase,
                                                     int max_n_kmers, uint32_t t, uint32_t &slot) {
  int lo = 0, hi = K;  // kpre[lo] <= t < kpre[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (T.kpre[mid] <= t) lo = mid; else hi = mid;
  }
  int sn = 0;
  while (sn + 1 < P.S.n_seeds && kbase[sn + 1] <= lo) sn++;
  slot = (uint32_t)(sn * max_n_kmers + (lo - kbase[sn]));
  return P.I.pos[sn] + T.kst[lo] + (t - T.kpre[lo]);
}
__device__ __forceinline__ void flat_entry(const ScanParams &P, const KmerTables &T, int K, const int *kbase,
                                           int max_n_kmers, uint32_t t, uint32_t &x, uint32_t &slot) {
  x = __ldg(flat_addr(P, T, K, kbase, max_n_kmers, t, slot));
}

// Walks entries [t, tend) of the flattened lists as one contiguous stream per thread: one search for the first
// k-mer, then list after list with four loads in flight; f(position, k-mer slot) per entry.  A thread's stream
// reads whole sectors (8 positions each), so the lists still cost one HBM/L2 sector fetch per 32 bytes.
template <typename F>
__device__ __forceinline__ void walk_entries(const ScanParams &P, const KmerTables &T, int K, const int *kbase,
                                             int max_n_kmers, uint32_t t, uint32_t tend, F f) {
  if (t >= tend) return;
  int lo = 0, hi = K;  // kpre[lo] <= t < kpre[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (T.kpre[mid] <= t) lo = mid; else hi = mid;
  }
  int sn = 0;
  while (sn + 1 < P.S.n_seeds && kbase[sn + 1] <= lo) sn++;
  int k = lo;
  while (t < tend) {
    const uint32_t kp = T.kpre[k], lend = min(tend, T.kpre[k + 1]);
    if (lend > t) {
      const uint32_t *p = P.I.pos[sn] + T.kst[k] + (t - kp);
      const uint32_t slot = (uint32_t)(sn * max_n_kmers + (k - kbase[sn]));
      uint32_t n = lend - t;
      while (n >= 4) {
        const uint32_t x0 = __ldg(p), x1 = __ldg(p + 1), x2 = __ldg(p + 2), x3 = __ldg(p + 3);
        f(x0, slot); f(x1, slot); f(x2, slot); f(x3, slot);
        p += 4;
        n -= 4;
      }
      while (n > 0) {
        f(__ldg(p), slot);
        p++;
        n--;
      }
      t = lend;
    }
    k++;
    while (sn + 1 < P.S.n_seeds && k >= kbase[sn + 1]) sn++;
  }
}

__device__ __forceinline__ uint32_t region_hash(uint32_t region, int bm_log2) {
  return (region * 2654435761u) >> (32 - bm_log2);
}
__device__ __forceinline__ void region_mark(uint32_t *bm1, uint32_t *bm2, uint32_t region, int bm_log2) {
  const uint32_t h = region_hash(region, bm_log2), bit = 1u << (h & 31);
  const uint32_t old = atomicOr(&bm1[h >> 5], bit);
  if (old & bit) atomicOr(&bm2[h >> 5], bit);
}
__device__ __forceinline__ bool region_twice(const uint32_t *bm2, uint32_t region, int bm_log2) {
  const uint32_t h = region_hash(region, bm_log2);
  return (bm2[h >> 5] >> (h & 31)) & 1u;
}


// ---- hit list helpers (read_get_hit_list_per_strand, mapping.c:1025-1229) ------------------------------
// Chain search of anchor i (:1058-1158): best predecessor and window-generation score; true = the window passes.
__device__ __forceinline__ bool hit_chain(const ScanParams &P, const AnchorRec *rec, int i, int rl, int window_len,
                                          int &max_idx, int &max_score) {
  const MapParamsDev &M = P.M;
  const AnchorRec ai = rec[i];
  const long long coff = (long long)P.G.contig_off[ai.cn];
  const long long glen = (long long)P.G.contig_len[ai.cn];
  int w_len = window_len;
  if ((long long)w_len > glen) w_len = (int)glen;
  long long gend = ((long long)ai.x - coff) + rl - 1 - ai.y, gstart;
  if (gend > glen - 1) gend = glen - 1;
  gstart = gend >= window_len ? gend - window_len : 0;
  max_idx = i;
  max_score = ai.len * M.match;
  // hit-list mode 3 (paired -n 3): a window whose mate has two k-mer hits in range passes on one match (:1082-1094);
  // the flag rides in bit 30 of the head's weight
  const bool heavy_mp = M.match_mode == 3 && ((ai.weight >> 30) & 1);
  if (!M.gapless) {
    if ((M.match_mode == 2 || (M.match_mode == 3 && !heavy_mp)) && (ai.weight & 0x3fffffff) == 1) max_score = -1;
    for (int j = i - 1; j >= 0; j--) {
      const AnchorRec aj = rec[j];
      if ((long long)aj.x < coff + gstart) break;
      if (aj.y >= ai.y) continue;
      int short_len, long_len;
      if ((long long)ai.x - coff - ai.y > (long long)aj.x - coff - aj.y) {
        short_len = (int)(ai.y - aj.y) + ai.len;
        long_len = (int)((long long)ai.x - (long long)aj.x) + ai.len;
      } else {
        short_len = (int)((long long)ai.x - (long long)aj.x) + ai.len;
        long_len = (int)(ai.y - aj.y) + ai.len;
      }
      int tmp_score = short_len * M.match;
      if (long_len > short_len) tmp_score += M.b_gap_open + (long_len - short_len) * M.b_gap_ext;  // :1134
      if (tmp_score > max_score) {
        max_idx = j;
        max_score = tmp_score;
      }
    }
  }
  const int base_len = rl < w_len ? rl : w_len;
  const int score_max = base_len * M.match;
  return M.gapless || M.match_mode == 1 || heavy_mp ||
         max_score >= (int)abs_or_pct_d(M.wgen_thr, M.wgen_frac, (double)score_max);
}

// The window of anchor i chained to max_idx (:1161-1204)
__device__ __forceinline__ DevHit hit_make(const ScanParams &P, const AnchorRec *rec, int i, int max_idx, int max_score,
                                           int rl, int window_len) {
  const MapParamsDev &M = P.M;
  const AnchorRec ai = rec[i], am = rec[max_idx];
  const int cn = ai.cn;
  const long long coff = (long long)P.G.contig_off[cn];
  const long long glen = (long long)P.G.contig_len[cn];
  int w_len = window_len;
  if ((long long)w_len > glen) w_len = (int)glen;
  const int base_len = rl < w_len ? rl : w_len;
  DevHit h;
  const int x_len = (int)((long long)ai.x - (long long)am.x) + ai.len;
  long long goff;
  if ((long long)((window_len - x_len) / 2) < (long long)am.x - coff)
    goff = ((long long)am.x - coff) - (window_len - x_len) / 2;
  else
    goff = 0;
  if (goff + w_len > glen) goff = glen - w_len;
  const long long rel = coff + goff;
  if (max_idx < i) {
    anchor_join2((long long)ai.x - rel, ai.y, ai.len, 1, (long long)am.x - rel, am.y, am.len, 1, h.ax, h.ay, h.alen,
                 h.awidth);
  } else {
    h.ax = (int)((long long)ai.x - rel);
    h.ay = ai.y;
    h.alen = ai.len;
    h.awidth = 1;
  }
  h.g_off = (uint32_t)goff;
  h.cn = cn;
  h.w_len = w_len;
  h.wg = max_score;
  h.matches = (M.gapless || max_idx == i) ? (ai.weight & 0x3fffffff) : (ai.weight & 0x3fffffff) + (am.weight & 0x3fffffff);
  h.score_max = base_len * M.match;
  h.score_vector = -1;
  h.pct_vector = 0;
  return h;
}
#define HITPACK_EMIT (1ull << 63)
__device__ __forceinline__ unsigned long long hit_pack(bool emit, int max_idx, int max_score) {
  return (emit ? HITPACK_EMIT : 0ull) | ((unsigned long long)(uint32_t)max_idx << 32) | (uint32_t)max_score;
}

// stable insertion sort by g_off inside a contig (:1210-1223); the list is almost sorted
__device__ __forceinline__ void hit_sort_serial(DevHit *H, int nh) {
  for (int i = 1; i < nh; i++) {
    const DevHit cur = H[i];
    int j = i;
    while (j >= 1 && H[j - 1].cn == cur.cn && H[j - 1].g_off > cur.g_off) j--;
    if (j < i) {
      for (int k = i - 1; k >= j; k--) H[k + 1] = H[k];
      H[j] = cur;
    }
  }
}

// Colinear collapse (read_get_anchor_list_per_strand :941-971 with anchor_uw_join, anchors.c:98-119) of the
// candidates rec[0, m_surv) -- in pop order, ascending x, so only the "extend to the right" branch of the join can
// fire -- by one warp in lockstep, 32 candidates per step: a candidate merges into the anchor of the previous
// candidate with the same cache slot (diagonal mod read length) iff that one lies on the same diagonal of the same
// contig; predecessors inside the step come from match_any, earlier ones from the cache; merged anchors take their
// length and weight by atomics.  Anchors are written back in place (anchor id <= candidate index).  Returns n_anch.
__device__ __forceinline__ int collapse_lockstep(AnchorRec *rec, int32_t *cache, int m_surv, int rl, int lane) {
  const uint32_t lt = (1u << lane) - 1u;
  int n_anch = 0;
  for (int t0 = 0; t0 < m_surv; t0 += 32) {
    const int t = t0 + lane;
    const bool valid = t < m_surv;
    AnchorRec a;
    a.x = 0; a.cn = 0; a.y = 0; a.len = 0; a.weight = 0;
    if (valid) a = rec[t];
    const long long diag = (long long)a.x - a.y;
    // the reference's cache slot (x + len - y) % len of the diagonal, in 32-bit arithmetic (y < len)
    const int slot = valid ? (int)((a.x % (uint32_t)rl + (uint32_t)rl - (uint32_t)a.y) % (uint32_t)rl) : -1 - lane;
    const uint32_t grp = __match_any_sync(0xffffffffu, slot);
    const uint32_t below = grp & lt;
    const int pl = below ? 31 - __clz(below) : lane;
    const long long pdiag = __shfl_sync(0xffffffffu, diag, pl);
    const int pcn = __shfl_sync(0xffffffffu, a.cn, pl);
    bool head = false;
    int cached = -1;
    if (valid) {
      if (below) {
        head = !(pdiag == diag && pcn == a.cn);
      } else {
        cached = cache[slot];
        head = !(cached >= 0 && rec[cached].cn == a.cn && (long long)rec[cached].x - rec[cached].y == diag);
      }
    }
    const uint32_t hm = __ballot_sync(0xffffffffu, head);
    // the anchor this candidate belongs to: the nearest head of its slot at or below it, else the cached one
    int aid = -1;
    if (valid) {
      const uint32_t hb = grp & hm & (lt | (1u << lane));
      if (hb) {
        const int hl = 31 - __clz(hb);
        aid = n_anch + __popc(hm & ((1u << hl) - 1u));
      } else {
        // no head below in this step: the chain starts at the lowest lane of the group, which read the cache
        const int first = __ffs(grp) - 1;
        aid = -2 - first;   // resolved below
      }
    }
    const int cached_first = __shfl_sync(0xffffffffu, cached, (aid <= -2) ? (-2 - aid) : lane);
    if (aid <= -2) aid = cached_first;
    __syncwarp();
    if (valid && head) rec[aid] = a;   // aid <= t: every slot at or below t0+31 is already in registers
    __syncwarp();
    if (valid && !head) {
      const AnchorRec d = rec[aid];
      const int newlen = (int)((long long)a.x - (long long)d.x) + a.len;
      atomicMax((unsigned int *)&rec[aid].y, ((unsigned int)newlen << 16) | (unsigned int)(uint16_t)d.y);
      atomicAdd(&rec[aid].weight, 1);
    }
    if (valid && (grp >> lane) <= 1u) cache[slot] = aid;   // highest lane of the group
    n_anch += __popc(hm);
    __syncwarp();
  }
  return n_anch;
}

// ---- steps 4-7 on the sorted candidates ent[0, total): one warp ---------------------------------------
__device__ void scan_tail(const ScanParams &P, uint32_t rs, int r, int rl, int max_n_kmers, int total, int gathered,
                          unsigned long long *ent, AnchorRec *rec, int32_t *cache, uint32_t *keep, int32_t *first_of,
                          unsigned long long *heap64, uint16_t *order16, int lane) {
  const SeedTable &S = P.S;
  const MapParamsDev &M = P.M;
  const int mkp = M.colour_space ? 1 : 0;
  const uint32_t rmask = (1u << M.region_bits) - 1u;
  // ---- 4. region filter (RG_HAS_2 as a neighbour test) + ordered compaction -----------------
  int m_surv = total;
  if (M.use_region_counts) {
    for (int t0 = 0; t0 < total; t0 += 32) {
      const int t = t0 + lane;
      bool kp = false;
      if (t < total) {
        const unsigned long long x = ent[t] >> 32;
        const unsigned long long xl = t > 0 ? (ent[t - 1] >> 32) : 0ull;
        const unsigned long long xr = t + 1 < total ? (ent[t + 1] >> 32) : ~0ull;
        const unsigned long long region = x >> M.region_bits;
        const unsigned long long lo = region << M.region_bits;
        const unsigned long long hi = ((region + 1) << M.region_bits) + (unsigned long long)M.region_overlap;
        kp = (t > 0 && xl >= lo) || (t + 1 < total && xr < hi);
        if (!kp && region > 0 && ((uint32_t)x & rmask) < (uint32_t)M.region_overlap) {
          const unsigned long long lo2 = (region - 1) << M.region_bits;
          const unsigned long long hi2 = lo + (unsigned long long)M.region_overlap;
          kp = (t > 0 && xl >= lo2) || (t + 1 < total && xr < hi2);
        }
      }
      const uint32_t b = __ballot_sync(0xffffffffu, kp);
      if (lane == 0) keep[t0 >> 5] = b;
    }
    __syncwarp();
    int outn = 0;
    for (int t0 = 0; t0 < total; t0 += 32) {
      const uint32_t b = keep[t0 >> 5];
      const int t = t0 + lane;
      const unsigned long long v = t < total ? ent[t] : 0ull;
      __syncwarp();
      if ((b >> lane) & 1u) ent[outn + __popc(b & ((1u << lane) - 1u))] = v;
      outn += __popc(b);
      __syncwarp();
    }
    m_surv = outn;
  }
  if (lane == 0) {
    atomicAdd(&P.stats64[0], (unsigned long long)gathered);
    atomicAdd(&P.stats64[1], (unsigned long long)m_surv);
  }
  if (m_surv == 0) return;

  // ---- 5. equal position on different read offsets -> replay the reference's heap -----------
  bool tie = false;
  for (int t = lane; t + 1 < m_surv; t += 32) {
    const unsigned long long a = ent[t], b = ent[t + 1];
    if ((a >> 32) == (b >> 32) && ((uint32_t)a % (uint32_t)max_n_kmers) != ((uint32_t)b % (uint32_t)max_n_kmers))
      tie = true;
  }
  tie = __any_sync(0xffffffffu, tie);
  const uint16_t *order = nullptr;
  if (tie) {
    // everything the emulation touches is in shared memory: first_of[] reuses the k-mer table (dead by now),
    // the heap holds (position << 32 | survivor index), the per-slot chains live in the spare upper half of
    // each entry's low word, the pop order goes to order16[]
    const int K = S.n_seeds * max_n_kmers;
    for (int k = lane; k < K; k += 32) first_of[k] = -1;
    __syncwarp();
    if (lane == 0) {
      atomicAdd(&P.stats[0], 1u);
      for (int t = m_surv - 1; t >= 0; t--) {
        const unsigned long long e = ent[t];
        const int off = (int)((uint32_t)e & 0xffffu);
        const uint32_t nx = first_of[off] < 0 ? 0xffffu : (uint32_t)first_of[off];
        ent[t] = (e & 0xffffffff0000ffffull) | ((unsigned long long)nx << 16);
        first_of[off] = t;
      }
      // heap_uu on key = position; elements are survivor indices.  Load order: sn-major, i.e.
      // ascending slot (mapping.c:913-935).
      int load = 0;
      for (int off = 0; off < K; off++) {
        const int t = first_of[off];
        if (t < 0) continue;
        heap64[load++] = (ent[t] & 0xffffffff00000000ull) | (unsigned)t;
        int node = load, parent = node / 2;  // percolate_up, heap.h:43-60
        while (node > 1 && (heap64[node - 1] >> 32) < (heap64[parent - 1] >> 32)) {
          const unsigned long long tmp = heap64[parent - 1];
          heap64[parent - 1] = heap64[node - 1];
          heap64[node - 1] = tmp;
          node = parent;
          parent = node / 2;
        }
      }
      int outn = 0;
      while (load > 0) {
        const int t = (int)(uint32_t)heap64[0];
        order16[outn++] = (uint16_t)t;
        const uint32_t nx = ((uint32_t)ent[t] >> 16) & 0xffffu;
        if (nx != 0xffffu) {
          heap64[0] = (ent[nx] & 0xffffffff00000000ull) | nx;  // heap_uu_replace_min
        } else {
          load--;  // heap_uu_extract_min
          if (load > 0) heap64[0] = heap64[load];
        }
        if (load > 0) {  // percolate_down, heap.h:62-89
          int node = 1;
          const unsigned long long cur = heap64[0];
          for (;;) {
            const int left = node * 2, right = left + 1;
            int mn = node;
            unsigned long long mk = cur;
            if (left <= load) {
              const unsigned long long lk = heap64[left - 1];
              if ((lk >> 32) < (mk >> 32)) { mn = left; mk = lk; }
            }
            if (right <= load) {
              const unsigned long long rk = heap64[right - 1];
              if ((rk >> 32) < (mk >> 32)) { mn = right; mk = rk; }
            }
            if (mn == node) break;
            heap64[node - 1] = mk;
            heap64[mn - 1] = cur;
            node = mn;
          }
        }
      }
    }
    __syncwarp();
    order = order16;
  }

  // ---- 6. anchors: contig lookup in parallel, colinear collapse in pop order ----------------
  for (int t = lane; t < m_surv; t += 32) {
    const unsigned long long e = ent[order ? order[t] : t];
    const uint32_t slot = (uint32_t)e & 0xffffu;
    const int sn = (int)(slot / (uint32_t)max_n_kmers), i = (int)(slot % (uint32_t)max_n_kmers);
    AnchorRec a;
    a.x = (uint32_t)(e >> 32);
    a.y = (int16_t)(mkp + i);
    a.len = (int16_t)S.span[sn];
    a.weight = 1;
    a.cn = contig_of_dev(P.G.contig_off, P.G.num_contigs, a.x);
    rec[t] = a;
  }
  for (int t = lane; t < rl; t += 32) cache[t] = -1;
  __syncwarp();
  const int n_anch = collapse_lockstep(rec, cache, m_surv, rl, lane);
  __syncwarp();

  // ---- 7. hit list (read_get_hit_list_per_strand) ---------------------------------------------
  // chain search per anchor first (results parked in ent[], dead by now), so that only the windows that
  // pass reserve hit slots
  const int window_len = (int)(unsigned short)abs_or_pct_d(M.window_len, M.window_len_frac, (double)rl);
  int nh = 0;
  for (int i0 = 0; i0 < n_anch; i0 += 32) {
    const int i = i0 + lane;
    bool emit = false;
    if (i < n_anch) {
      int max_idx, max_score;
      emit = hit_chain(P, rec, i, rl, window_len, max_idx, max_score);
      ent[i] = hit_pack(emit, max_idx, max_score);
    }
    nh += __popc(__ballot_sync(0xffffffffu, emit));
  }
  __syncwarp();
  if (lane == 0) atomicAdd(&P.stats64[2], (unsigned long long)n_anch);
  if (nh == 0) return;
  uint32_t out0 = 0;
  if (lane == 0) out0 = atomicAdd(P.hits_used, (uint32_t)nh);
  out0 = __shfl_sync(0xffffffffu, out0, 0);
  if ((unsigned long long)out0 + (unsigned long long)nh > (unsigned long long)P.hits_cap) {
    if (lane == 0) atomicOr(P.status, 1u);
    return;
  }
  int at = 0;
  for (int i0 = 0; i0 < n_anch; i0 += 32) {
    const int i = i0 + lane;
    const unsigned long long pk = i < n_anch ? ent[i] : 0ull;
    const bool emit = (pk & HITPACK_EMIT) != 0;
    const uint32_t b = __ballot_sync(0xffffffffu, emit);
    if (emit)
      P.hits[out0 + at + __popc(b & ((1u << lane) - 1u))] =
          hit_make(P, rec, i, (int)((pk >> 32) & 0x7fffffffu), (int)(uint32_t)pk, rl, window_len);
    at += __popc(b);
  }
  __syncwarp();
  if (lane == 0) {
    hit_sort_serial(P.hits + out0, nh);
    P.rs_range[rs] = make_uint2(out0, (uint32_t)nh);
  }
  __syncwarp();
}

// shared memory of one warp of the small kernel / of the CTA of the big kernel
struct ScanSmem {
  size_t r2, kst, kpre, bm1, bm2, ent, rec, cache, keep, heap, order, stash, sslot, total;
};
__host__ __device__ inline ScanSmem scan_layout(int cap, int max_rl, int k_cap, int bm_log2, int stash) {
  ScanSmem L;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
  L.r2 = take(((size_t)max_rl / 16 + 4) * 4);
  L.kst = take((size_t)k_cap * 4);
  L.kpre = take(((size_t)k_cap + 1) * 4);
  L.cache = take((size_t)max_rl * 4);
  L.keep = take(((size_t)cap / 32 + 2) * 4);   // + the candidate counter of the warp kernel
  // Lifetimes: stash / sslot (the strand's list entries, staged once: warp kernel, short lists) and the bitmaps die
  // with pass B; the replay heap dies with the tie replay, which leaves `order`; the anchors (16 B per candidate)
  // are born after that from ent + order.  So `order` lives in the stash, and the anchors overlay sslot, the
  // bitmaps and the heap.  The sort pads ent to a power of two, possibly past `cap`: into sslot, dead by then.
  L.stash = take((size_t)stash * 4 > (size_t)cap * 2 ? (size_t)stash * 4 : (size_t)cap * 2);
  L.order = L.stash;                    // pop order of the tie replay
  L.ent = take((size_t)cap * 8);
  L.sslot = take((size_t)stash * 2);
  L.bm1 = take(((size_t)1 << bm_log2) / 8);
  L.bm2 = take(((size_t)1 << bm_log2) / 8);
  L.heap = take((size_t)k_cap * 8);
  L.rec = L.sslot;
  if (o - L.rec < (size_t)cap * 16) o = L.rec + (size_t)cap * 16;
  L.total = (o + 15) & ~(size_t)15;
  return L;
}

// number of k-mers of seed sn on a read strand (gmapper.c:473-480, mapping.c:47-66)
__device__ __forceinline__ int n_kmers_of(const SeedTable &S, int sn, int rl, int mkp) {
  const int nk = rl - S.span[sn] + 1 - mkp;
  return nk > 0 ? nk : 0;
}

// Steps 1-4 for one warp; returns the number of candidates in ent (sorted), or -1 when the strand needs more
// than `cap` candidate slots / k_cap k-mers (-> overflow list).
__global__ void __launch_bounds__(SCAN_WARPS * 32) scan_kernel(const ScanParams P, int warps_per_cta) {
  extern __shared__ unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  if (wib >= warps_per_cta) return;
  const int cap = P.cap;
  const ScanSmem L = scan_layout(cap, P.max_rl, P.k_cap, P.bm_log2, P.stash);
  unsigned char *base = smem_raw + L.total * wib;
  uint32_t *r2 = (uint32_t *)(base + L.r2);
  KmerTables T;
  T.kst = (uint32_t *)(base + L.kst);
  T.kpre = (uint32_t *)(base + L.kpre);
  uint32_t *bm1 = (uint32_t *)(base + L.bm1), *bm2 = (uint32_t *)(base + L.bm2);
  unsigned long long *ent = (unsigned long long *)(base + L.ent);
  AnchorRec *rec = (AnchorRec *)(base + L.rec);
  int32_t *cache = (int32_t *)(base + L.cache);
  uint32_t *keep = (uint32_t *)(base + L.keep);
  uint32_t *s_cnt = keep + (cap / 32 + 1);   // candidate counter of pass B
  const int bm_words = 1 << (P.bm_log2 - 5);

  const uint32_t gwarp = blockIdx.x * warps_per_cta + wib;
  const uint32_t n_warps = gridDim.x * warps_per_cta;
  const uint32_t n_items = 2u * (uint32_t)P.n_reads;
  const SeedTable &S = P.S;
  const MapParamsDev &M = P.M;
  const int mkp = M.colour_space ? 1 : 0;  // min_kmer_pos (gmapper.c:478-480)
  const uint32_t rmask = (1u << M.region_bits) - 1u;

  for (uint32_t item = gwarp; item < n_items; item += n_warps) {
    const uint32_t rs = item;
    const int r = (int)(rs >> 1);
    const int rl = P.read_len[r];
    const uint32_t *seq = P.reads + (size_t)rs * P.stride;
    int max_n_kmers = M.colour_space ? rl - S.min_span : rl - S.min_span + 1;  // gmapper.c:473-479
    if (max_n_kmers < 0) max_n_kmers = 0;
    if (lane == 0) P.rs_range[rs] = make_uint2(0u, 0u);
    if (rl <= 0 || max_n_kmers == 0) continue;
    __syncwarp();

    // ---- 1. recode the read, project every k-mer once, fetch bucket bounds --------------------------
    const int nw2 = (rl + 15) / 16;
    for (int w = lane; w < nw2 + 3; w += 32) {
      uint32_t v = 0;
      if (w < nw2) {
        v = squeeze8(seq[2 * w]);
        if (2 * w + 1 < P.stride) v |= squeeze8(seq[2 * w + 1]) << 16;
      }
      r2[w] = v;
    }
    for (int w = lane; w < bm_words; w += 32) {
      bm1[w] = 0u;
      bm2[w] = 0u;
    }
    __syncwarp();
    int kbase[SHRIMP_MAX_SEEDS + 1];
    int K = 0;
    for (int sn = 0; sn < S.n_seeds; sn++) {
      kbase[sn] = K;
      K += n_kmers_of(S, sn, rl, mkp);
    }
    kbase[S.n_seeds] = K;
    if (K > P.k_cap || S.n_seeds * max_n_kmers > P.k_cap) {  // longer than this launch's shared-memory tables: CTA kernel
      if (lane == 0) P.overflow[atomicAdd(P.n_overflow, 1u)] = rs;
      continue;
    }
    uint32_t total = 0;
    for (int sn = 0; sn < S.n_seeds; sn++) {
      const int nk = kbase[sn + 1] - kbase[sn];
      for (int i0 = 0; i0 < nk; i0 += 32) {
        const int i = i0 + lane;
        uint32_t start = 0, len = 0;
        if (i < nk) {
          const uint32_t m = S.n_runs[sn] ? mapidx_fast(S, sn, r2, mkp + i) : kmer_to_mapidx(S, sn, seq, (uint64_t)(mkp + i));
          if (P.I.head_off[sn]) {   // bucket head: length and list in one 64-byte record (genome.cuh)
            const uint32_t hw = P.I.head_off[sn] + m * SHRIMP_HEAD_WORDS;
            const uint2 hv = __ldg((const uint2 *)(P.I.pos[sn] + hw));
            len = hv.x;
            start = len <= SHRIMP_HEAD_WORDS - 1 ? hw + 1 : hv.y;
          } else {
            const uint2 be = make_uint2(__ldg(P.I.offs[sn] + m), __ldg(P.I.offs[sn] + m + 1));
            start = be.x;
            len = be.y - be.x;
          }
          if (len > M.list_cutoff) len = 0;  // mapping.c:497,:889
        }
        const uint32_t incl = (uint32_t)warp_incl_scan((int)len, lane);
        if (i < nk) {
          T.kst[kbase[sn] + i] = start;
          T.kpre[kbase[sn] + i] = total + incl - len;
        }
        total += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    if (lane == 0) T.kpre[K] = total;
    __syncwarp();
    if (total == 0) continue;

    int ns = (int)total;
    if (M.use_region_counts && total <= (uint32_t)P.stash) {
      // short lists: every entry is fetched ONCE, by an asynchronous 4-byte copy into shared memory (all the
      // strand's loads in flight together), then both passes run on the staged copy
      uint32_t *stash = (uint32_t *)(base + L.stash);
      uint16_t *sslot = (uint16_t *)(base + L.sslot);
      // a lane per list (the lists are a handful of entries each); lists of more than 32 entries are copied by
      // the whole warp afterwards
      for (int k0 = 0; k0 < K; k0 += 32) {
        const int kk = k0 + lane;
        uint32_t f0 = 0, n = 0, slot = 0;
        const uint32_t *p = nullptr;
        if (kk < K) {
          f0 = T.kpre[kk];
          n = T.kpre[kk + 1] - f0;
          int sn = 0;
          while (sn + 1 < S.n_seeds && kbase[sn + 1] <= kk) sn++;
          p = P.I.pos[sn] + T.kst[kk];
          slot = (uint32_t)(sn * max_n_kmers + (kk - kbase[sn]));
        }
        if (n <= 32u)
          for (uint32_t j = 0; j < n; j++) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stash + f0 + j);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(p + j) : "memory");
            sslot[f0 + j] = (uint16_t)slot;
          }
        uint32_t big = __ballot_sync(0xffffffffu, n > 32u);
        while (big) {
          const int src = __ffs(big) - 1;
          big &= big - 1;
          const uint32_t bf0 = __shfl_sync(0xffffffffu, f0, src), bn = __shfl_sync(0xffffffffu, n, src);
          const uint32_t bslot = __shfl_sync(0xffffffffu, slot, src);
          const unsigned long long bp = __shfl_sync(0xffffffffu, (unsigned long long)p, src);
          for (uint32_t j = lane; j < bn; j += 32) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stash + bf0 + j);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"((const uint32_t *)bp + j) : "memory");
            sslot[bf0 + j] = (uint16_t)bslot;
          }
        }
      }
      asm volatile("cp.async.wait_all;\n" ::: "memory");
      __syncwarp();
      for (uint32_t t = lane; t < total; t += 32) {   // pass A: mark regions
        const uint32_t x = stash[t];
        const uint32_t region = x >> M.region_bits;
        region_mark(bm1, bm2, region, P.bm_log2);
        if ((x & rmask) < (uint32_t)M.region_overlap && region > 0) region_mark(bm1, bm2, region - 1, P.bm_log2);
      }
      __syncwarp();
      ns = 0;
      for (uint32_t t0 = 0; t0 < total; t0 += 32) {   // pass B: keep the entries of regions marked twice
        const uint32_t t = t0 + lane;
        bool kp = false;
        uint32_t x = 0;
        if (t < total) {
          x = stash[t];
          const uint32_t region = x >> M.region_bits;
          kp = region_twice(bm2, region, P.bm_log2) ||
               ((x & rmask) < (uint32_t)M.region_overlap && region > 0 && region_twice(bm2, region - 1, P.bm_log2));
        }
        const uint32_t b = __ballot_sync(0xffffffffu, kp);
        if (kp) {
          const int at = ns + __popc(b & ((1u << lane) - 1u));
          if (at < cap) ent[at] = ((unsigned long long)x << 32) | sslot[t];
        }
        ns += __popc(b);
      }
    } else if (M.use_region_counts) {
      if (P.stream) {  // long lists: one contiguous stream per lane
        // ---- 2. pass A: mark regions ------------------------------------------------------------------
        const uint32_t chunk = (total + 31u) / 32u;
        const uint32_t tb = min(total, lane * chunk), te = min(total, tb + chunk);
        walk_entries(P, T, K, kbase, max_n_kmers, tb, te, [&](uint32_t x, uint32_t) {
          const uint32_t region = x >> M.region_bits;
          region_mark(bm1, bm2, region, P.bm_log2);
          if ((x & rmask) < (uint32_t)M.region_overlap && region > 0) region_mark(bm1, bm2, region - 1, P.bm_log2);
        });
        if (lane == 0) *s_cnt = 0u;
        __syncwarp();
        // ---- 3. pass B: keep the entries of regions marked twice ----------------------------------------
        walk_entries(P, T, K, kbase, max_n_kmers, tb, te, [&](uint32_t x, uint32_t slot) {
          const uint32_t region = x >> M.region_bits;
          const bool kp = region_twice(bm2, region, P.bm_log2) ||
                          ((x & rmask) < (uint32_t)M.region_overlap && region > 0 &&
                           region_twice(bm2, region - 1, P.bm_log2));
          if (kp) {
            const uint32_t at = atomicAdd(s_cnt, 1u);
            if (at < (uint32_t)cap) ent[at] = ((unsigned long long)x << 32) | slot;
          }
        });
        __syncwarp();
        ns = (int)*s_cnt;
      } else {         // short lists: entries interleaved over the lanes, one search per entry
        // ---- 2. pass A: mark regions ----------------------------------------------------------------
        for (uint32_t t = lane; t < total; t += 32) {
          uint32_t x, slot;
          flat_entry(P, T, K, kbase, max_n_kmers, t, x, slot);
          const uint32_t region = x >> M.region_bits;
          region_mark(bm1, bm2, region, P.bm_log2);
          if ((x & rmask) < (uint32_t)M.region_overlap && region > 0) region_mark(bm1, bm2, region - 1, P.bm_log2);
        }
        __syncwarp();
        // ---- 3. pass B: keep the entries of regions marked twice --------------------------------------
        ns = 0;
        for (uint32_t t0 = 0; t0 < total; t0 += 32) {
          const uint32_t t = t0 + lane;
          bool kp = false;
          uint32_t x = 0, slot = 0;
          if (t < total) {
            flat_entry(P, T, K, kbase, max_n_kmers, t, x, slot);
            const uint32_t region = x >> M.region_bits;
            kp = region_twice(bm2, region, P.bm_log2) ||
                 ((x & rmask) < (uint32_t)M.region_overlap && region > 0 && region_twice(bm2, region - 1, P.bm_log2));
          }
          const uint32_t b = __ballot_sync(0xffffffffu, kp);
          if (kp) {
            const int at = ns + __popc(b & ((1u << lane) - 1u));
            if (at < cap) ent[at] = ((unsigned long long)x << 32) | slot;
          }
          ns += __popc(b);
        }
      }
    } else {
      if (ns <= cap) {
        for (uint32_t t = lane; t < total; t += 32) {
          uint32_t x, slot;
          flat_entry(P, T, K, kbase, max_n_kmers, t, x, slot);
          ent[t] = ((unsigned long long)x << 32) | slot;
        }
      }
    }
    if (ns > cap) {
      if (lane == 0) P.overflow[atomicAdd(P.n_overflow, 1u)] = rs;
      continue;
    }
    if (ns == 0) {
      if (lane == 0) atomicAdd(&P.stats64[0], (unsigned long long)total);
      continue;
    }
    int Pn = 32;
    while (Pn < ns) Pn <<= 1;
    for (int t = ns + lane; t < Pn; t += 32) ent[t] = ~0ull;
    __syncwarp();
    // ---- 4a. bitonic sort by position ---------------------------------------------------------------
    for (int k = 2; k <= Pn; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < (Pn >> 1); t += 32) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          const int l = i | j;
          const unsigned long long a = ent[i], b = ent[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            ent[i] = b;
            ent[l] = a;
          }
        }
        __syncwarp();
      }
    }
    scan_tail(P, rs, r, rl, max_n_kmers, ns, (int)total, ent, rec, cache, keep, (int32_t *)T.kst,
              (unsigned long long *)(base + L.heap), (uint16_t *)(base + L.order), lane);
    __syncwarp();
  }
}

// ---- CTA per read strand: the dense regime (long index lists: large genomes) -----------------------------
// Same steps as the warp kernel, organised for thousands of list entries per strand:
//   * every thread projects k-mers; a WARP streams one index list at a time (lane-strided, coalesced 128-byte
//     requests, four loads in flight per lane), lists handed out through a shared-memory counter;
//   * the region bitmaps are EXACT (one bit per 2 kb region, no hashing): the genome is cut into partitions of
//     2^bm_log2 regions, each list is cut once per strand by a binary search (lists ascend), and the partitions are
//     filtered one after the other with the same two bitmaps.  Entries of the first / last region of a partition
//     depend on marks across the cut and are kept unconditionally; the exact neighbour test of step 4 decides;
//   * CTA-wide bitonic sort, neighbour test and compaction; the colinear collapse runs in warp lockstep
//     (32 candidates per step: predecessors on the same diagonal slot by match_any, merged anchors by atomics);
//   * the chain search of the hit list runs on all threads; only windows that pass reserve hit slots.
// Strands with more candidates than the shared-memory slab go to a second launch whose candidate arrays live in
// a global slab per CTA (P.g_ent != nullptr).
struct CtaSmem {
  size_t r2, kst, klen, kA, kpre, cuts, cache, keep, order, ent, bm1, bm2, rec, heap, buf, bslot, total;
};
__host__ __device__ inline CtaSmem cta_layout(int cap, int max_rl, int k_cap, int bm_log2, int n_part, int win,
                                              bool global_arrays) {
  CtaSmem L;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
  const size_t c = global_arrays ? 0 : (size_t)cap;
  L.r2 = take(((size_t)max_rl / 16 + 4) * 4);
  L.kst = take((size_t)k_cap * 4);
  L.klen = take((size_t)k_cap * 4);
  L.kA = take((size_t)k_cap * 4);
  L.kpre = take(((size_t)k_cap + 1) * 4);
  L.buf = take((size_t)win * 4);
  L.bslot = take((size_t)win * 2);
  L.cuts = take((size_t)k_cap * (size_t)((n_part > 1 && win > 16) ? n_part - 1 : 0) * 4);   // win 16 = cursor walk: no cuts
  L.cache = take((size_t)max_rl * 4);
  L.heap = take((size_t)k_cap * 8);
  L.keep = take((c / 32 + 2) * 4);
  L.order = take(c * 2 + 16);
  L.ent = take(c * 8);
  L.bm1 = take(((size_t)1 << bm_log2) / 8);
  L.bm2 = take(((size_t)1 << bm_log2) / 8);
  L.rec = L.bm1;   // anchors: born after the bitmaps have died
  if (o - L.bm1 < c * 16) o = L.bm1 + c * 16;
  L.total = (o + 15) & ~(size_t)15;
  return L;
}

#define SCAN_CTA_MAX_THREADS 768
// phase timing (P.prof != nullptr): thread 0 adds the cycles since the previous mark to prof[phase]
#define PROF_MARK(ph)                                                   \
  do {                                                                  \
    if (P.prof && tid == 0) {                                           \
      const long long _now = clock64();                                 \
      atomicAdd(&P.prof[ph], (unsigned long long)(_now - prof_t));      \
      prof_t = _now;                                                    \
    }                                                                   \
  } while (0)
// GLOB: candidate arrays in global slabs (else in shared memory: the template keeps the address space static)
template <bool GLOB>
__global__ void __launch_bounds__(SCAN_CTA_MAX_THREADS, 2) scan_cta_kernel(const ScanParams P) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint32_t s_item_next, s_total, s_ns, s_next, s_cnt, s_out0, s_wsum[SCAN_CTA_MAX_THREADS / 32];
  __shared__ uint32_t s_bcnt[64], s_boff[65];
  __shared__ int s_nanch, kbase[SHRIMP_MAX_SEEDS + 1];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nthr = blockDim.x, nwarps = nthr >> 5;
  constexpr bool glob = GLOB;
  const int cap = glob ? P.g_cap : P.cap;
  const int n_part = P.n_part;
  const CtaSmem L = cta_layout(P.cap, P.max_rl, P.k_cap, P.bm_log2, n_part, P.win, glob);
  const int win = P.win;
  uint32_t *kA = (uint32_t *)(smem_raw + L.kA), *kpre = (uint32_t *)(smem_raw + L.kpre);
  uint32_t *buf = (uint32_t *)(smem_raw + L.buf);
  uint16_t *bslot = (uint16_t *)(smem_raw + L.bslot);
  uint32_t *r2 = (uint32_t *)(smem_raw + L.r2);
  uint32_t *kst = (uint32_t *)(smem_raw + L.kst), *klen = (uint32_t *)(smem_raw + L.klen);
  uint32_t *cuts = (uint32_t *)(smem_raw + L.cuts);
  int32_t *cache = (int32_t *)(smem_raw + L.cache);
  unsigned long long *heap64 = (unsigned long long *)(smem_raw + L.heap);
  uint32_t *bm1 = (uint32_t *)(smem_raw + L.bm1), *bm2 = (uint32_t *)(smem_raw + L.bm2);
  const size_t slab = (size_t)blockIdx.x;
  // 16-byte aligned slabs (kpfx is 32-bit)
  unsigned long long *const ent = glob ? P.g_ent + slab * (size_t)(cap + 1) : (unsigned long long *)(smem_raw + L.ent);
  AnchorRec *const rec = glob ? P.g_rec + slab * (size_t)(cap + 1) : (AnchorRec *)(smem_raw + L.rec);
  uint16_t *const order16 = glob ? P.g_order + slab * (size_t)((cap + 15) & ~7) : (uint16_t *)(smem_raw + L.order);
  uint32_t *const keep = glob ? P.g_keep + slab * (size_t)(cap / 32 + 2) : (uint32_t *)(smem_raw + L.keep);
  uint32_t *kpfx = (uint32_t *)order16;   // compaction offsets; the pop order is written later
  const int bm_words = 1 << (P.bm_log2 - 5);
  const SeedTable &S = P.S;
  const MapParamsDev &M = P.M;
  const int mkp = M.colour_space ? 1 : 0;
  const uint32_t rmask = (1u << M.region_bits) - 1u;
  const bool filt = M.use_region_counts != 0;
  const bool hashed = P.bm_hashed != 0;   // hashed bitmaps (one partition): false positives only add candidates
  const uint32_t lt = (1u << lane) - 1u;
  const int glog = P.lanes_per_list_log2, g = 1 << glog, lpw = 32 >> glog, gl = lane & (g - 1);

  long long prof_t = clock64();
  if (tid == 0) s_item_next = atomicAdd(P.work_counter, 1u);
  uint32_t mp_epoch = (P.mp_mode && !P.resume) ? P.mp_epoch[blockIdx.x] : 0u;
  for (;;) {
    __syncthreads();
    PROF_MARK(15);
    const uint32_t item = s_item_next;
    __syncthreads();
    if (item >= P.n_work) break;
    if (tid == 0) {
      s_total = 0;
      s_ns = 0;
      s_cnt = 0;
    }
    // the next strand's ticket travels while this one is processed (the last warp waits for it at its next barrier)
    if (tid == nthr - 1) s_item_next = atomicAdd(P.work_counter, 1u);
    // Mate-pair region counts (P.mp_mode): a work item is a read PAIR.  Its four strands first mark
    // this CTA's exact region tables in global memory (sub-steps 0-3: read_get_region_counts), then each strand
    // keeps the entries that pass the paired rule of advance_index_in_genomemap (sub-steps 4-7), with the mate's
    // count taken over the region interval the insert-size limits allow (read_get_mp_region_counts).
    const bool mp = P.mp_mode != 0 && !P.resume;
    // with a work list (strands that overflowed a smaller slab) the pair's tables are rebuilt and one strand is kept
    const int n_sub = mp ? (P.work ? 5 : 8) : 1;
    const uint32_t mp_pair = mp ? (P.work ? P.work[item] >> 2 : item) : 0u;
    if (mp) mp_epoch++;
    for (int sub = 0; sub < n_sub; sub++) {
    if (sub > 0) {
      __syncthreads();
      if (tid == 0) {
        s_total = 0;
        s_ns = 0;
        s_cnt = 0;
      }
    }
    const bool mp_mark = mp && sub < 4;
    const int q = mp_mark ? sub : (P.work ? (int)(P.work[item] & 3u) : sub - 4);   // mp: strand of the pair, 2 * mate + strand
    const uint32_t rs = mp ? 4u * mp_pair + (uint32_t)q : P.resume ? P.tie_rec[item].x : P.work ? P.work[item] : item;
    const int r = (int)(rs >> 1);
    const int rl = P.read_len[r];
    const uint32_t *seq = P.reads + (size_t)rs * P.stride;
    int max_n_kmers = M.colour_space ? rl - S.min_span : rl - S.min_span + 1;
    if (max_n_kmers < 0) max_n_kmers = 0;
    if (!P.work && !P.resume && !mp_mark && tid == 0) P.rs_range[rs] = make_uint2(0u, 0u);
    if (rl <= 0 || max_n_kmers == 0) continue;
    uint32_t *tab_self = nullptr;
    const uint32_t *tab_mate = nullptr;
    int dr_min = 0, dr_max = 0;
    if (mp) {
      tab_self = P.mp_tab + ((size_t)blockIdx.x * 4 + (size_t)q) * (size_t)P.mp_regions;
      tab_mate = P.mp_tab + ((size_t)blockIdx.x * 4 + (size_t)(3 - q)) * (size_t)P.mp_regions;
      // readpair_compute_mp_ranges (mapping.c:2317-2430) for this mate and strand
      const int nip = (q >> 1) & 1, st = q & 1;
      const int rl1 = P.read_len[(rs >> 2) * 2], rl2 = P.read_len[(rs >> 2) * 2 + 1];
      const int wl1 = (int)(unsigned short)abs_or_pct_d(M.window_len, M.window_len_frac, (double)rl1);
      const int wl2 = (int)(unsigned short)abs_or_pct_d(M.window_len, M.window_len_frac, (double)rl2);
      int d1min[2], d1max[2];
      d1min[0] = P.min_insert - wl2;
      d1max[0] = P.max_insert + (wl1 - rl1) - rl2;
      d1min[1] = -P.max_insert + rl1 + (rl2 - wl2);
      d1max[1] = -P.min_insert + wl1;
      const int sh = P.pair_mode == 2 ? rl1 + rl2 : P.pair_mode == 3 ? rl2 : P.pair_mode == 4 ? rl1 : 0;
      d1min[0] += sh; d1max[0] += sh; d1min[1] -= sh; d1max[1] -= sh;
      int lo, hi;
      if (nip == 0) {
        lo = d1min[st]; hi = d1max[st];
      } else if (P.pair_mode == 1 || P.pair_mode == 2) {
        lo = -d1max[1 - st]; hi = -d1min[1 - st];
      } else {
        lo = -d1max[st]; hi = -d1min[st];
      }
      const int rsz = 1 << M.region_bits;
      dr_min = lo >= 0 ? lo / rsz : -1 - (-lo - 1) / rsz;
      dr_max = hi > 0 ? 1 + (hi - 1) / rsz : -(-hi / rsz);
    }
    // the paired keep rule for one region (mapping.c:704-712); also the hit list's heavy_mp (:1082-1094)
    auto mp_count = [&](uint32_t region) -> int {
      int first = (int)region + dr_min, last = (int)region + dr_max, mx = 0;
      if (first < 0) first = 0;
      if (last > P.mp_regions - 1) last = P.mp_regions - 1;
      for (int k = first; k <= last && mx < 2; k++) {
        const uint32_t c = tab_mate[k];
        if ((c >> 8) == mp_epoch) mx = (c & 2u) ? 2 : 1;
      }
      return mx;
    };
    auto mp_pass = [&](uint32_t region) -> bool {
      const int count_main = (tab_self[region] & 2u) ? 2 : 1, count_mp = mp_count(region);
      return (P.mp_mode == 1 && count_main >= 2 && count_mp >= 2) || (P.mp_mode == 2 && (count_main >= 2 || count_mp >= 2)) ||
             (P.mp_mode == 3 && count_mp >= 1 && count_main + count_mp >= 3);
    };

    int m_surv = 0;
    const unsigned long long *esrc = ent;   // candidates in pop order: esrc[order ? order[t] : t]
    const uint16_t *order = nullptr;
    if (P.resume) {
      // strands whose heap order was replayed by scan_replay_kernel: candidates and pop order come from the tie slab
      const uint4 tr = P.tie_rec[item];
      m_surv = (int)tr.z;
      if (m_surv <= P.resume_min || m_surv > cap) continue;
      esrc = P.tie_ent + tr.y;
      order = P.tie_order + tr.y;
    } else {
    PROF_MARK(0);
    // ---- 1. recode, project, bucket bounds ----------------------------------------------------------------
    const int nw2 = (rl + 15) / 16;
    for (int w = tid; w < nw2 + 3; w += nthr) {
      uint32_t v = 0;
      if (w < nw2) {
        v = squeeze8(seq[2 * w]);
        if (2 * w + 1 < P.stride) v |= squeeze8(seq[2 * w + 1]) << 16;
      }
      r2[w] = v;
    }
    int K = 0;
    for (int sn = 0; sn < S.n_seeds; sn++) {
      if (tid == 0) kbase[sn] = K;   // shared: the previous strand's readers passed the barrier at the loop top
      K += n_kmers_of(S, sn, rl, mkp);
    }
    if (tid == 0) kbase[S.n_seeds] = K;
    if (K > P.k_cap || S.n_seeds * max_n_kmers > P.k_cap) {
      if (tid == 0) atomicOr(P.status, 2u);
      continue;
    }
    __syncthreads();
    {
      uint32_t mytot = 0;
      for (int kk = tid; kk < K; kk += nthr) {
        int sn = 0;
        while (kk >= kbase[sn + 1]) sn++;
        const int i = kk - kbase[sn];
        const uint32_t m = S.n_runs[sn] ? mapidx_fast(S, sn, r2, mkp + i) : kmer_to_mapidx(S, sn, seq, (uint64_t)(mkp + i));
        const uint32_t start = __ldg(P.I.offs[sn] + m);
        uint32_t len = __ldg(P.I.offs[sn] + m + 1) - start;
        if (len > M.list_cutoff) len = 0;  // mapping.c:497,:889
        kst[kk] = start;
        klen[kk] = len;
        mytot += len;
      }
      mytot = (uint32_t)warp_sum((int)mytot);
      if (lane == 0 && mytot) atomicAdd(&s_total, mytot);
    }
    __syncthreads();
    const uint32_t total = s_total;
    if (total == 0) continue;
    PROF_MARK(1);
    if (P.walk && !mp) {
      // ---- 2./3. cursor walk: g lanes per index list step through it in ascending order, tile by tile (a tile = the
      // 2^bm_log2 regions one pair of exact bitmaps covers; one tile with hashed bitmaps or without a region filter).
      // Pass A marks the regions of the tile's entries, pass B walks the same entries again (L1 / L2) and keeps those of
      // regions marked twice; the cursor of every list then moves past the tile.  No staging buffer, no cut search:
      // the lists ascend, so the entries of a tile are the next ones below the tile's end.
      for (int kk = tid; kk < K; kk += nthr) kA[kk] = 0u;
      const int n_tiles = (filt && !hashed) ? n_part : 1;
      for (int part = 0; part < n_tiles; part++) {
        const uint32_t R0 = (uint32_t)part << P.bm_log2;
        const uint32_t Rlast = R0 + ((1u << P.bm_log2) - 1u);
        const unsigned long long pend64 = (unsigned long long)(part + 1) << (P.bm_log2 + M.region_bits);
        const bool last_tile = part == n_tiles - 1 || pend64 > 0xffffffffull;
        const uint32_t pend = last_tile ? 0xffffffffu : (uint32_t)pend64;
        if (filt)
          for (int w = tid; w < bm_words; w += nthr) {
            bm1[w] = 0u;
            bm2[w] = 0u;
          }
        __syncthreads();
        const int rb = M.region_bits, bml = P.bm_log2;
        const uint32_t ovl = (uint32_t)M.region_overlap;
        const bool edge_lo = n_tiles > 1 && part > 0, edge_hi = n_tiles > 1 && part < n_tiles - 1;
        for (int pass = filt ? 0 : 1; pass < 2; pass++) {
          for (int k0 = wid * lpw; k0 < K; k0 += nwarps * lpw) {   // lpw lists per warp and step, g lanes each
            const int kk = k0 + (lane >> glog);
            uint32_t c0 = 0, c = 0, len = 0, slot = 0;
            const uint32_t *p = nullptr;
            if (kk < K) {
              int sn = 0;
              while (kk >= kbase[sn + 1]) sn++;
              c0 = kA[kk] + (uint32_t)gl;
              len = klen[kk];
              p = P.I.pos[sn] + kst[kk];
              slot = (uint32_t)(sn * max_n_kmers + (kk - kbase[sn]));
            }
            c = c0;
            // every lane walks its own stride of the list; the next entry is already in flight
            uint32_t x = c < len ? __ldg(p + c) : 0u;
            if (pass == 0) {
              while (c < len && (last_tile || x < pend)) {
                const uint32_t cn = c + (uint32_t)g;
                const uint32_t xn = cn < len ? __ldg(p + cn) : 0u;
                const uint32_t region = x >> rb;
                const uint32_t idx = hashed ? region_hash(region, bml) : region - R0;
                const uint32_t bit = 1u << (idx & 31);
                if (atomicOr(&bm1[idx >> 5], bit) & bit) atomicOr(&bm2[idx >> 5], bit);
                if ((x & rmask) < ovl && (hashed ? region > 0 : idx > 0)) {
                  const uint32_t idx2 = hashed ? region_hash(region - 1, bml) : idx - 1;
                  const uint32_t bit2 = 1u << (idx2 & 31);
                  if (atomicOr(&bm1[idx2 >> 5], bit2) & bit2) atomicOr(&bm2[idx2 >> 5], bit2);
                }
                c = cn;
                x = xn;
              }
            } else {
              while (c < len && (last_tile || x < pend)) {
                const uint32_t cn = c + (uint32_t)g;
                const uint32_t xn = cn < len ? __ldg(p + cn) : 0u;
                bool kp = true;
                if (filt) {
                  const uint32_t region = x >> rb;
                  const uint32_t idx = hashed ? region_hash(region, bml) : region - R0;
                  kp = ((bm2[idx >> 5] >> (idx & 31)) & 1u) != 0;
                  if (!kp && (x & rmask) < ovl && (hashed ? region > 0 : idx > 0)) {
                    const uint32_t idx2 = hashed ? region_hash(region - 1, bml) : idx - 1;
                    kp = ((bm2[idx2 >> 5] >> (idx2 & 31)) & 1u) != 0;
                  }
                  // marks across a tile cut are not seen here: keep, the neighbour test decides
                  if (!kp && ((edge_lo && region == R0) || (edge_hi && region == Rlast))) kp = true;
                }
                if (kp) {   // survivors are few: one shared-memory atomic each
                  const uint32_t at = atomicAdd(&s_ns, 1u);
                  if (at < (uint32_t)cap) ent[at] = ((unsigned long long)x << 32) | slot;
                }
                c = cn;
                x = xn;
              }
              if (n_tiles > 1) {   // the list's cursor moves past the tile: the entries its g lanes consumed
                __syncwarp();
                uint32_t tot = (c - c0) >> glog;
                for (int o = 1; o < g; o <<= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                if (kk < K && gl == 0) kA[kk] += tot;
              }
            }
          }
          __syncthreads();   // all marks are in before pass B tests them
          PROF_MARK(3 + pass);
        }
      }
    } else {
    // cut every list at the partition boundaries
    if (n_part > 1) {
      const int nc = n_part - 1;
      for (int q = tid; q < K * nc; q += nthr) {
        const int kk = q / nc, c = q - kk * nc;
        int sn = 0;
        while (kk >= kbase[sn + 1]) sn++;
        const uint32_t *p = P.I.pos[sn] + kst[kk];
        const unsigned long long key = (unsigned long long)(c + 1) << (P.bm_log2 + M.region_bits);
        uint32_t lo = 0, hi = klen[kk];   // first entry >= key
        if (key > 0xffffffffull) lo = hi;
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if ((unsigned long long)__ldg(p + mid) < key) lo = mid + 1; else hi = mid;
        }
        cuts[q] = lo;
      }
      __syncthreads();
    }

    PROF_MARK(2);
    // ---- 2./3. per partition: the list entries are staged in shared memory (asynchronous 4-byte copies, several
    // lists per warp), then pass A marks the regions and pass B keeps the entries of regions marked twice as flat
    // loops over the staged window.  A partition with more entries than the window is staged window by window,
    // twice (the second time from L2).
    for (int part = 0; part < n_part; part++) {
      const uint32_t R0 = (uint32_t)part << P.bm_log2;
      const uint32_t Rlast = R0 + ((1u << P.bm_log2) - 1u);
      // sub-range of every list in this partition and the prefix sums of their lengths
      if (wid == 0) {
        uint32_t run = 0;
        for (int k0 = 0; k0 < K; k0 += 32) {
          const int kk = k0 + lane;
          uint32_t a = 0, b = 0;
          if (kk < K) {
            b = klen[kk];
            if (n_part > 1) {
              if (part > 0) a = cuts[kk * (n_part - 1) + part - 1];
              if (part < n_part - 1) b = cuts[kk * (n_part - 1) + part];
            }
            if (b < a) b = a;
          }
          const uint32_t len = b - a;
          const uint32_t inc = (uint32_t)warp_incl_scan((int)len, lane);
          if (kk < K) {
            kA[kk] = a;
            kpre[kk] = run + inc - len;
          }
          run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) kpre[K] = run;
      }
      if (filt)
        for (int w = tid; w < bm_words; w += nthr) {
          bm1[w] = 0u;
          bm2[w] = 0u;
        }
      __syncthreads();
      const uint32_t total_p = kpre[K];
      if (total_p == 0) continue;
      const bool one_window = total_p <= (uint32_t)win;
      for (int pass = mp ? (mp_mark ? 0 : 1) : filt ? 0 : 1; pass < (mp_mark ? 1 : 2); pass++) {
        for (uint32_t w0 = 0; w0 < total_p; w0 += (uint32_t)win) {
          const uint32_t w1 = min(w0 + (uint32_t)win, total_p), n = w1 - w0;
          if (!(pass == 1 && one_window && filt && !mp)) {
            // ---- stage the window: lists [klo, ...) whose flat range meets [w0, w1)
            int klo = 0;
            if (w0 > 0) {
              int hi = K;  // kpre[klo] <= w0 < kpre[hi]
              while (hi - klo > 1) {
                const int mid = (klo + hi) >> 1;
                if (kpre[mid] <= w0) klo = mid; else hi = mid;
              }
            }
            __syncthreads();   // the previous window's readers are done with buf
            if (tid == 0) s_next = (uint32_t)klo;
            __syncthreads();
            for (;;) {
              uint32_t kk = 0;
              if (lane == 0) kk = atomicAdd(&s_next, (uint32_t)lpw);
              kk = __shfl_sync(0xffffffffu, kk, 0);
              if (kk >= (uint32_t)K || kpre[kk] >= w1) break;
              kk += (uint32_t)(lane >> glog);
              if (kk >= (uint32_t)K) continue;
              const uint32_t f0 = kpre[kk], lo = max(f0, w0), hi = min(kpre[kk + 1], w1);
              if (lo >= hi) continue;
              int sn = 0;
              while ((int)kk >= kbase[sn + 1]) sn++;
              const uint32_t *p = P.I.pos[sn] + kst[kk] + kA[kk];
              const uint16_t slot = (uint16_t)(sn * max_n_kmers + ((int)kk - kbase[sn]));
              for (uint32_t f = lo + (uint32_t)gl; f < hi; f += (uint32_t)g) {
                const uint32_t d = f - w0;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(buf + d);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(p + (f - f0)) : "memory");
                bslot[d] = slot;
              }
            }
            asm volatile("cp.async.wait_all;\n" ::: "memory");
            __syncthreads();
          }
          if (pass == 0) {
            // ---- pass A: mark
            if (mp) {   // exact per-(mate, strand) tables in global memory: (epoch << 8) | has2 << 1 | touched
              const uint32_t fresh = (mp_epoch << 8) | 1u;
              for (uint32_t t = tid; t < n; t += nthr) {
                const uint32_t x = buf[t];
                uint32_t region = x >> M.region_bits;
                uint32_t old = atomicMax(&tab_self[region], fresh);
                if ((old >> 8) == mp_epoch) atomicOr(&tab_self[region], 2u);
                if ((x & rmask) < (uint32_t)M.region_overlap && region > 0) {
                  region--;
                  old = atomicMax(&tab_self[region], fresh);
                  if ((old >> 8) == mp_epoch) atomicOr(&tab_self[region], 2u);
                }
              }
            } else
            for (uint32_t t = tid; t < n; t += nthr) {
              const uint32_t x = buf[t];
              const uint32_t region = x >> M.region_bits;
              const uint32_t idx = hashed ? region_hash(region, P.bm_log2) : region - R0;
              const uint32_t bit = 1u << (idx & 31);
              const uint32_t old = atomicOr(&bm1[idx >> 5], bit);
              if (old & bit) atomicOr(&bm2[idx >> 5], bit);
              if ((x & rmask) < (uint32_t)M.region_overlap && (hashed ? region > 0 : idx > 0)) {
                const uint32_t idx2 = hashed ? region_hash(region - 1, P.bm_log2) : idx - 1;
                const uint32_t bit2 = 1u << (idx2 & 31);
                const uint32_t old2 = atomicOr(&bm1[idx2 >> 5], bit2);
                if (old2 & bit2) atomicOr(&bm2[idx2 >> 5], bit2);
              }
            }
          } else {
            // ---- pass B: keep
            for (uint32_t t0 = 0; t0 < n; t0 += nthr) {
              const uint32_t t = t0 + tid;
              bool kp = t < n;
              uint32_t x = 0;
              if (kp) {
                x = buf[t];
                if (mp) {
                  const uint32_t region = x >> M.region_bits;
                  kp = mp_pass(region) || ((x & rmask) < (uint32_t)M.region_overlap && region > 0 && mp_pass(region - 1));
                } else if (filt) {
                  const uint32_t region = x >> M.region_bits;
                  const uint32_t idx = hashed ? region_hash(region, P.bm_log2) : region - R0;
                  kp = ((bm2[idx >> 5] >> (idx & 31)) & 1u) != 0;
                  if (!kp && (x & rmask) < (uint32_t)M.region_overlap && (hashed ? region > 0 : idx > 0)) {
                    const uint32_t idx2 = hashed ? region_hash(region - 1, P.bm_log2) : idx - 1;
                    kp = ((bm2[idx2 >> 5] >> (idx2 & 31)) & 1u) != 0;
                  }
                  // marks across a partition cut are not seen here: keep, the neighbour test decides
                  if (!kp && n_part > 1 && ((part > 0 && region == R0) || (part < n_part - 1 && region == Rlast))) kp = true;
                }
              }
              const uint32_t bal = __ballot_sync(0xffffffffu, kp);
              if (bal) {
                uint32_t at = 0;
                if (lane == 0) at = atomicAdd(&s_ns, (uint32_t)__popc(bal));
                at = __shfl_sync(0xffffffffu, at, 0) + (uint32_t)__popc(bal & lt);
                if (kp && at < (uint32_t)cap) {
                  uint32_t lowword = bslot[t];
                  if (mp && M.match_mode == 3) {   // heavy_mp of the anchor this entry may head, bit 15 of the slot field
                    const uint32_t region = x >> M.region_bits;
                    if (mp_count(region) >= 2 ||
                        ((x & rmask) < (uint32_t)M.region_overlap && region > 0 && mp_count(region - 1) >= 2))
                      lowword |= 0x8000u;
                  }
                  ent[at] = ((unsigned long long)x << 32) | lowword;
                }
              }
            }
          }
        }
        __syncthreads();   // all marks are in before pass B tests them
        PROF_MARK(3 + pass);
      }
    }
    }
    if (mp_mark) continue;
    const int ns = (int)s_ns;
    if (ns > cap) {
      if (tid == 0) {
        if (glob || !P.overflow) atomicOr(P.status, 2u);
        else P.overflow[atomicAdd(P.n_overflow, 1u)] = rs;
      }
      continue;
    }
    if (ns == 0) {
      if (tid == 0) atomicAdd(&P.stats64[0], (unsigned long long)total);
      continue;
    }
    // ---- 4a. sort by position: the candidates are dealt into 64 bins by the high bits of their position (count,
    // prefix, scatter into the anchor array, which is still unused), every bin is sorted by one warp (a bitonic
    // network with the padding left virtual), and the bins are copied back in order.  A bin that is too large for
    // one warp (a tiny genome, a read of one repeat) sends the strand through the CTA-wide network below.
    bool binned = false;
    if (ns > 1024) {   // below that the CTA-wide network is as fast
      unsigned long long *const scratch = (unsigned long long *)rec;
      const int bshift = P.sort_shift;
      if (tid < 64) s_bcnt[tid] = 0u;
      __syncthreads();
      for (int t = tid; t < ns; t += nthr) atomicAdd(&s_bcnt[(uint32_t)min(63ull, (ent[t] >> 32) >> bshift)], 1u);
      __syncthreads();
      if (wid == 0) {
        const uint32_t c0 = s_bcnt[2 * lane], c1 = s_bcnt[2 * lane + 1];
        const uint32_t inc = (uint32_t)warp_incl_scan((int)(c0 + c1), lane);
        s_boff[2 * lane] = inc - c0 - c1;
        s_boff[2 * lane + 1] = inc - c1;
        if (lane == 31) s_boff[64] = inc;
        const uint32_t mx = max(c0, c1);
        const uint32_t big = __ballot_sync(0xffffffffu, mx > 1024u);
        if (lane == 0) s_next = big ? 1u : 0u;
        s_bcnt[2 * lane] = 0u;
        s_bcnt[2 * lane + 1] = 0u;
      }
      __syncthreads();
      if (s_next == 0u) {
        binned = true;
        for (int t = tid; t < ns; t += nthr) {
          const unsigned long long e = ent[t];
          const uint32_t bin = (uint32_t)min(63ull, (e >> 32) >> bshift);
          scratch[s_boff[bin] + atomicAdd(&s_bcnt[bin], 1u)] = e;
        }
        __syncthreads();
        // bins of up to 128 candidates: one warp each
        for (int bin = wid; bin < 64; bin += nwarps) {
          const int n = (int)(s_boff[bin + 1] - s_boff[bin]);
          if (n < 2 || n > 128) continue;
          unsigned long long *a = scratch + s_boff[bin];
          int P2 = 2;
          while (P2 < n) P2 <<= 1;
          for (int k = 2; k <= P2; k <<= 1) {
            const int h = k >> 1;
            for (int t = lane; t < (P2 >> 1); t += 32) {   // flip: i in the lower half of its block against its mirror
              const int blk = t / h, off = t - blk * h;
              const int i = blk * k + off, l = blk * k + (k - 1 - off);
              if (l < n) {
                const unsigned long long u = a[i], v = a[l];
                if (u > v) {
                  a[i] = v;
                  a[l] = u;
                }
              }
            }
            __syncwarp();
            for (int j = h >> 1; j > 0; j >>= 1) {
              for (int t = lane; t < (P2 >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                if (l < n) {
                  const unsigned long long u = a[i], v = a[l];
                  if (u > v) {
                    a[i] = v;
                    a[l] = u;
                  }
                }
              }
              __syncwarp();
            }
          }
        }
        // larger bins (the read's own locus: one entry per k-mer): the whole CTA, one bin after the other
        for (int bin = 0; bin < 64; bin++) {
          const int n = (int)(s_boff[bin + 1] - s_boff[bin]);
          if (n <= 128) continue;
          unsigned long long *const a_glob = scratch + s_boff[bin];
          // global slabs: the bin is sorted in shared memory (the bitmaps are dead by now) and written back
          const bool staged = glob && (size_t)n * 8 <= ((size_t)2 << P.bm_log2) / 8;
          unsigned long long *a = staged ? (unsigned long long *)bm1 : a_glob;
          if (staged) {
            for (int t = tid; t < n; t += nthr) a[t] = a_glob[t];
            __syncthreads();
          }
          int P2 = 256;
          while (P2 < n) P2 <<= 1;
          for (int k = 2; k <= P2; k <<= 1) {
            const int h = k >> 1;
            for (int t = tid; t < (P2 >> 1); t += nthr) {
              const int blk = t / h, off = t - blk * h;
              const int i = blk * k + off, l = blk * k + (k - 1 - off);
              if (l < n) {
                const unsigned long long u = a[i], v = a[l];
                if (u > v) {
                  a[i] = v;
                  a[l] = u;
                }
              }
            }
            __syncthreads();
            for (int j = h >> 1; j > 0; j >>= 1) {
              for (int t = tid; t < (P2 >> 1); t += nthr) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                if (l < n) {
                  const unsigned long long u = a[i], v = a[l];
                  if (u > v) {
                    a[i] = v;
                    a[l] = u;
                  }
                }
              }
              __syncthreads();
            }
          }
          if (staged) {
            for (int t = tid; t < n; t += nthr) a_glob[t] = a[t];
            __syncthreads();
          }
        }
        __syncthreads();
        for (int t = tid; t < ns; t += nthr) ent[t] = scratch[t];
        __syncthreads();
      }
    }
    int Pn = 32;
    while (Pn < ns) Pn <<= 1;
    if (binned) Pn = 1;   // sorted already
    for (int t = ns + tid; t < Pn; t += nthr) ent[t] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= Pn; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = tid; t < (Pn >> 1); t += nthr) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          const int l = i | j;
          const unsigned long long a = ent[i], b = ent[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            ent[i] = b;
            ent[l] = a;
          }
        }
        __syncthreads();
      }
    }
    PROF_MARK(5);
    // ---- 4b. region filter (RG_HAS_2 as a neighbour test) + ordered compaction ----------------------------
    m_surv = ns;
    if (filt && !mp) {
      const int n_words = (ns + 31) >> 5;
      for (int t0 = 0; t0 < n_words * 32; t0 += nthr) {
        const int t = t0 + tid;
        bool kp = false;
        if (t < ns) {
          const unsigned long long x = ent[t] >> 32;
          const unsigned long long xl = t > 0 ? (ent[t - 1] >> 32) : 0ull;
          const unsigned long long xr = t + 1 < ns ? (ent[t + 1] >> 32) : ~0ull;
          const unsigned long long region = x >> M.region_bits;
          const unsigned long long lo = region << M.region_bits;
          const unsigned long long hi = ((region + 1) << M.region_bits) + (unsigned long long)M.region_overlap;
          kp = (t > 0 && xl >= lo) || (t + 1 < ns && xr < hi);
          if (!kp && region > 0 && ((uint32_t)x & rmask) < (uint32_t)M.region_overlap) {
            const unsigned long long lo2 = (region - 1) << M.region_bits;
            const unsigned long long hi2 = lo + (unsigned long long)M.region_overlap;
            kp = (t > 0 && xl >= lo2) || (t + 1 < ns && xr < hi2);
          }
        }
        const uint32_t b = __ballot_sync(0xffffffffu, kp);
        if (lane == 0 && (t >> 5) < n_words) keep[t >> 5] = b;
      }
      __syncthreads();
      if (wid == 0) {  // exclusive prefix of the popcounts
        uint32_t run = 0;
        for (int w0 = 0; w0 < n_words; w0 += 32) {
          const int w = w0 + lane;
          const int c = w < n_words ? __popc(keep[w]) : 0;
          const int inc = warp_incl_scan(c, lane);
          if (w < n_words) kpfx[w] = run + (uint32_t)(inc - c);
          run += (uint32_t)__shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) s_cnt = run;
      }
      __syncthreads();
      m_surv = (int)s_cnt;
      for (int t0 = 0; t0 < n_words * 32; t0 += nthr) {
        const int t = t0 + tid;
        unsigned long long v = 0;
        uint32_t at = 0;
        bool kp = false;
        if (t < ns) {
          const uint32_t b = keep[t >> 5];
          kp = (b >> lane) & 1u;
          v = ent[t];
          at = kpfx[t >> 5] + (uint32_t)__popc(b & lt);
        }
        __syncthreads();
        if (kp) ent[at] = v;
        __syncthreads();
      }
      if (tid == 0) s_cnt = 0;
    }
    if (tid == 0) {
      atomicAdd(&P.stats64[0], (unsigned long long)total);
      atomicAdd(&P.stats64[1], (unsigned long long)m_surv);
    }
    if (m_surv == 0) continue;
    __syncthreads();

    PROF_MARK(6);
    // ---- 5. equal position on different read offsets -> replay the reference's heap (thread 0) -------------
    bool tie = false;
    for (int t = tid; t + 1 < m_surv; t += nthr) {
      const unsigned long long a = ent[t], b = ent[t + 1];
      if ((a >> 32) == (b >> 32) &&
          (((uint32_t)a & 0x7fffu) % (uint32_t)max_n_kmers) != (((uint32_t)b & 0x7fffu) % (uint32_t)max_n_kmers))
        tie = true;
    }
    tie = __syncthreads_or(tie) != 0;
    if (tie) {
      // the pop order of equal positions is the reference's binary heap order (SURVEY hard part 3a): park the
      // candidates in the tie slab; scan_replay_kernel replays heap_uu for all such strands at once (a warp each),
      // then this kernel resumes them at step 6
      if (tid == 0) {
        const uint32_t off = atomicAdd(P.tie_used, (uint32_t)m_surv);
        uint32_t idx = 0xffffffffu;
        if ((unsigned long long)off + (unsigned long long)m_surv <= (unsigned long long)P.tie_cap) {
          idx = atomicAdd(P.n_tie, 1u);
          if (idx < P.tie_rec_cap) P.tie_rec[idx] = make_uint4(rs, off, (uint32_t)m_surv, 0u);
          else idx = 0xffffffffu;
        }
        if (idx == 0xffffffffu) atomicOr(P.status, 4u);
        s_out0 = idx == 0xffffffffu ? idx : off;
        atomicAdd(&P.stats[0], 1u);
      }
      __syncthreads();
      const uint32_t off = s_out0;
      if (off != 0xffffffffu)
        for (int t = tid; t < m_surv; t += nthr) P.tie_ent[off + t] = ent[t];
      continue;
    }
    }  // !P.resume

    PROF_MARK(7);
    // ---- 6. anchors in pop order; colinear collapse (:941-971) in warp lockstep ---------------------------
    for (int t = tid; t < m_surv; t += nthr) {
      const unsigned long long e = esrc[order ? order[t] : t];
      const uint32_t slot = (uint32_t)e & 0x7fffu;
      const int sn = (int)(slot / (uint32_t)max_n_kmers), i = (int)(slot % (uint32_t)max_n_kmers);
      AnchorRec a;
      a.x = (uint32_t)(e >> 32);
      a.y = (int16_t)(mkp + i);
      a.len = (int16_t)S.span[sn];
      a.weight = 1 | (int)(((uint32_t)e >> 15) & 1u) << 30;
      a.cn = contig_of_dev(P.G.contig_off, P.G.num_contigs, a.x);
      rec[t] = a;
    }
    for (int t = tid; t < rl; t += nthr) cache[t] = -1;
    __syncthreads();
    PROF_MARK(8);
    if (wid == 0) {
      const int n_anch = collapse_lockstep(rec, cache, m_surv, rl, lane);
      if (lane == 0) s_nanch = n_anch;
    }
    __syncthreads();
    const int n_anch = s_nanch;

    PROF_MARK(9);
    // ---- 7. hit list: chain search on all threads, then ordered emission ------------------------------------
    const int window_len = (int)(unsigned short)abs_or_pct_d(M.window_len, M.window_len_frac, (double)rl);
    {
      int mine = 0;
      for (int i = tid; i < n_anch; i += nthr) {
        int max_idx, max_score;
        const bool emit = hit_chain(P, rec, i, rl, window_len, max_idx, max_score);
        ent[i] = hit_pack(emit, max_idx, max_score);
        mine += emit ? 1 : 0;
      }
      mine = warp_sum(mine);
      if (lane == 0 && mine) atomicAdd(&s_cnt, (uint32_t)mine);
    }
    __syncthreads();
    PROF_MARK(10);
    const int nh = (int)s_cnt;
    if (tid == 0) {
      atomicAdd(&P.stats64[2], (unsigned long long)n_anch);
      uint32_t o = 0;
      if (nh > 0) {
        o = atomicAdd(P.hits_used, (uint32_t)nh);
        if ((unsigned long long)o + (unsigned long long)nh > (unsigned long long)P.hits_cap) {
          atomicOr(P.status, 1u);
          o = 0xffffffffu;
        }
      }
      s_out0 = o;
    }
    __syncthreads();
    const uint32_t out0 = s_out0;
    if (nh == 0 || out0 == 0xffffffffu) continue;
    uint32_t basepos = 0;
    for (int i0 = 0; i0 < n_anch; i0 += nthr) {
      const int i = i0 + tid;
      const unsigned long long pk = i < n_anch ? ent[i] : 0ull;
      const bool emit = (pk & HITPACK_EMIT) != 0;
      const uint32_t b = __ballot_sync(0xffffffffu, emit);
      if (lane == 0) s_wsum[wid] = (uint32_t)__popc(b);
      __syncthreads();
      uint32_t before = 0, all = 0;
      for (int w = 0; w < nwarps; w++) {
        const uint32_t c = s_wsum[w];
        if (w < wid) before += c;
        all += c;
      }
      if (emit)
        P.hits[out0 + basepos + before + (uint32_t)__popc(b & lt)] =
            hit_make(P, rec, i, (int)((pk >> 32) & 0x7fffffffu), (int)(uint32_t)pk, rl, window_len);
      basepos += all;
      __syncthreads();
    }
    __threadfence_block();
    __syncthreads();
    if (tid == 0) {
      hit_sort_serial(P.hits + out0, nh);
      P.rs_range[rs] = make_uint2(out0, (uint32_t)nh);
    }
    PROF_MARK(11);
    }  // sub-steps
  }
  if (P.mp_mode && !P.resume && tid == 0) P.mp_epoch[blockIdx.x] = mp_epoch;
}

size_t scan_smem_bytes(int cap, int max_rl, int k_cap, int bm_log2, int warps, int stash) {
  return scan_layout(cap, max_rl, k_cap, bm_log2, stash).total * warps;
}
// ---- heap-order replay for the strands with equal positions on different read offsets -------------------------
// One warp per parked strand.  The k-way merge of the reference pops equal keys in binary-heap order
// (read_get_anchor_list_per_strand, mapping.c:913-1006; heap_uu, common/heap.h:43-113): lanes build the per-list
// chains in parallel (match_any), lane 0 replays the heap in shared memory.  A heap element is the candidate word with
// its slot field replaced by the candidate's index: position << 32 | next-in-list << 16 | index.
#define REPLAY_WARPS 8
#define REPLAY_PRE 512
#define REPLAY_SMEM_PER_WARP(ks_cap) ((((size_t)(ks_cap) * 18 + (size_t)REPLAY_PRE * 12 + 47)) & ~(size_t)15)
__global__ void __launch_bounds__(REPLAY_WARPS * 32) scan_replay_kernel(const ScanParams P, uint32_t n_rec, int ks_cap) {
  extern __shared__ unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t w = blockIdx.x * (blockDim.x >> 5) + wib;   // (fewer than REPLAY_WARPS warps per CTA for long reads)
  if (w >= n_rec) return;
  unsigned char *base = smem_raw + (size_t)wib * REPLAY_SMEM_PER_WARP(ks_cap);
  unsigned long long *heap = (unsigned long long *)base;
  unsigned long long *stage = heap + ks_cap;
  // successor cache: the element that follows a heap member in its list is copied to shared memory (cp.async, 8 bytes)
  // when the member enters the heap, so that the pop loop -- one lane, a chain of dependent steps -- never waits for
  // global memory.  Direct-mapped on the successor's index; a miss falls back to the global load.
  unsigned long long *pre_val = stage + ks_cap;
  uint32_t *pre_tag = (uint32_t *)(pre_val + REPLAY_PRE);
  uint16_t *first_of = (uint16_t *)(pre_tag + REPLAY_PRE);
  const uint4 tr = P.tie_rec[w];
  const uint32_t rs = tr.x;
  const int m = (int)tr.z;
  unsigned long long *ent = P.tie_ent + tr.y;
  uint16_t *order = P.tie_order + tr.y;
  const int rl = P.read_len[rs >> 1];
  int max_n_kmers = P.M.colour_space ? rl - P.S.min_span : rl - P.S.min_span + 1;
  const int Ks = P.S.n_seeds * max_n_kmers;
  const uint32_t lt = (1u << lane) - 1u;
  for (int k = lane; k < Ks; k += 32) first_of[k] = 0xffffu;
  for (int k = lane; k < REPLAY_PRE; k += 32) pre_tag[k] = 0xffffffffu;
  __syncwarp();
  // chains: next candidate of the same list, built from the back
  for (int t0 = ((m - 1) >> 5) << 5; t0 >= 0; t0 -= 32) {
    const int t = t0 + lane;
    const bool valid = t < m;
    const unsigned long long e = valid ? ent[t] : 0ull;
    const uint32_t slot = valid ? ((uint32_t)e & 0x7fffu) : (0x10000u + (uint32_t)lane);   // bit 15: heavy_mp flag
    const uint32_t grp = __match_any_sync(0xffffffffu, slot);
    const uint32_t above = grp & ~(lt | (1u << lane));
    uint32_t nx = 0xffffu;
    if (valid) nx = above ? (uint32_t)(t0 + __ffs(above) - 1) : (uint32_t)first_of[slot];
    __syncwarp();
    if (valid && !(grp & lt)) first_of[slot] = (uint16_t)t;
    if (valid) ent[t] = (e & 0xffffffff0000ffffull) | ((unsigned long long)nx << 16);
    __syncwarp();
  }
  // heads of the lists, in ascending slot order (the load order of mapping.c:913-935)
  for (int k = lane; k < Ks; k += 32) {
    const uint32_t t = first_of[k];
    unsigned long long el = ~0ull;
    if (t != 0xffffu) {
      const unsigned long long e = ent[t];
      el = (e & 0xffffffffffff0000ull) | t;
      stage[k] = el;
    } else {
      stage[k] = el;
    }
  }
  __syncwarp();
  uint32_t seq = 0;   // copy groups issued so far (lane 0)
  if (lane == 0) {   // successors of the heads; a slot that is taken is left alone (that successor is loaded when needed)
    for (int k = 0; k < Ks; k++) {
      const unsigned long long el = stage[k];
      if (el == ~0ull) continue;
      const uint32_t nx = ((uint32_t)el >> 16) & 0xffffu;
      if (nx != 0xffffu) {
        const uint32_t h = nx & (REPLAY_PRE - 1);
        if (pre_tag[h] == 0xffffffffu) {
          pre_tag[h] = (nx << 16) | (seq & 0xffffu);
          const uint32_t dst = (uint32_t)__cvta_generic_to_shared(pre_val + h);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(ent + nx) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    seq++;
  }
  __syncwarp();
  if (lane != 0) return;
  int load = 0;
  for (int k = 0; k < Ks; k++) {
    const unsigned long long el = stage[k];
    if (el == ~0ull) continue;
    heap[load++] = el;
    int node = load, parent = node / 2;  // percolate_up, heap.h:43-60
    while (node > 1 && (heap[node - 1] >> 32) < (heap[parent - 1] >> 32)) {
      const unsigned long long tmp = heap[parent - 1];
      heap[parent - 1] = heap[node - 1];
      heap[node - 1] = tmp;
      node = parent;
      parent = node / 2;
    }
  }
  int outn = 0;
  while (load > 0) {
    const unsigned long long root = heap[0];
    order[outn++] = (uint16_t)((uint32_t)root & 0xffffu);
    const uint32_t nx = ((uint32_t)root >> 16) & 0xffffu;
    if (nx != 0xffffu) {
      const uint32_t h = nx & (REPLAY_PRE - 1);
      const uint32_t tg = pre_tag[h];
      unsigned long long e;
      if ((tg >> 16) == nx) {
        // copy groups complete in issue order: the newest eight may stay in flight unless this one is among them
        if (((seq - tg) & 0xffffu) <= 8u) asm volatile("cp.async.wait_all;\n" ::: "memory");
        else asm volatile("cp.async.wait_group 8;\n" ::: "memory");
        e = pre_val[h];
        pre_tag[h] = 0xffffffffu;
      } else {
        e = ent[nx];
      }
      const uint32_t nx2 = ((uint32_t)e >> 16) & 0xffffu;
      if (nx2 != 0xffffu) {
        const uint32_t h2 = nx2 & (REPLAY_PRE - 1);
        if (pre_tag[h2] == 0xffffffffu) {
          pre_tag[h2] = (nx2 << 16) | (seq & 0xffffu);
          const uint32_t dst = (uint32_t)__cvta_generic_to_shared(pre_val + h2);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(ent + nx2) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;\n" ::: "memory");
      seq++;
      heap[0] = (e & 0xffffffffffff0000ull) | nx;  // heap_uu_replace_min
    } else {
      load--;  // heap_uu_extract_min
      if (load > 0) heap[0] = heap[load];
    }
    if (load > 0) {  // percolate_down, heap.h:62-89
      int node = 1;
      const unsigned long long cur = heap[0];
      for (;;) {
        const int left = node * 2, right = left + 1;
        int mn = node;
        unsigned long long mk = cur;
        if (left <= load) {
          const unsigned long long lk = heap[left - 1];
          if ((lk >> 32) < (mk >> 32)) { mn = left; mk = lk; }
        }
        if (right <= load) {
          const unsigned long long rk = heap[right - 1];
          if ((rk >> 32) < (mk >> 32)) { mn = right; mk = rk; }
        }
        if (mn == node) break;
        heap[node - 1] = mk;
        heap[mn - 1] = cur;
        node = mn;
      }
    }
  }
}

int launch_scan_replay(shrimp_gpu_ctx *ctx, ScanParams &P, uint32_t n_rec, int ks_cap) {
  // a warp's heap, stage and first-of table grow with the k-mers of the longest read (18 bytes per k-mer slot: 53 KB at
  // 1,000 bases): as many warps per CTA as the shared memory of an SM holds (eight until round 2, which made the launch
  // of a chunk with reads of more than about 450 bases and equal-position ties fail with "invalid argument")
  const size_t per_warp = REPLAY_SMEM_PER_WARP(ks_cap);
  int warps = REPLAY_WARPS;
  while (warps > 1 && per_warp * warps > (size_t)SHRIMP_MAX_DYN_SMEM - 1024) warps >>= 1;
  if (per_warp * warps > (size_t)SHRIMP_MAX_DYN_SMEM - 1024) {
    set_error("seed scan: reads of this length need %zu bytes of shared memory per warp in the heap replay", per_warp);
    return SHRIMP_E_RANGE;
  }
  const size_t smem = per_warp * warps;
  SH_OPT_IN_SMEM(scan_replay_kernel, ctx->device);
  scan_replay_kernel<<<(n_rec + warps - 1) / warps, warps * 32, smem, ctx->stream>>>(P, n_rec, ks_cap);
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_SCAN);
  return SHRIMP_OK;
}

size_t scan_cta_smem_bytes(int cap, int max_rl, int k_cap, int bm_log2, int n_part, int win, bool global_arrays) {
  return cta_layout(cap, max_rl, k_cap, bm_log2, n_part, win, global_arrays).total;
}

int launch_scan(shrimp_gpu_ctx *ctx, ScanParams &P, int warps_per_cta, int n_ctas) {
  const size_t smem = scan_smem_bytes(P.cap, P.max_rl, P.k_cap, P.bm_log2, warps_per_cta, P.stash);
  SH_OPT_IN_SMEM(scan_kernel, ctx->device);
  scan_kernel<<<n_ctas, warps_per_cta * 32, smem, ctx->stream>>>(P, warps_per_cta);
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_SCAN);
  return SHRIMP_OK;
}

int launch_scan_cta(shrimp_gpu_ctx *ctx, ScanParams &P, int n_ctas, int threads) {
  const size_t smem = scan_cta_smem_bytes(P.cap, P.max_rl, P.k_cap, P.bm_log2, P.n_part, P.win, P.g_ent != nullptr);
  if (P.g_ent) {
    SH_OPT_IN_SMEM(scan_cta_kernel<true>, ctx->device);
    scan_cta_kernel<true><<<n_ctas, threads, smem, ctx->stream>>>(P);
  } else {
    SH_OPT_IN_SMEM(scan_cta_kernel<false>, ctx->device);
    scan_cta_kernel<false><<<n_ctas, threads, smem, ctx->stream>>>(P);
  }
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_SCAN);
  return SHRIMP_OK;
}

}  // namespace shrimp
