// Read pairs: shrimp_gpu_map_pairs = handle_readpair (gmapper/mapping.c:2504-2650) for a chunk of pairs with
// the default paired option set (gmapper.c:2638-2714; match_mode 4, half-paired).
//
// Device stages on top of the unpaired ones (seed scan and sw_vector are shared, pipeline.cu):
//   pair_up_kernel        readpair_pair_up_hits :266-325 with the ranges of readpair_compute_mp_ranges :2317-2442
//   pass1_replay_kernel   read_pass1 with only_paired (pass1.cu)
//   select_pairs_kernel   readpair_get_vector_hits :1877-1932 incl. extheap_paired_pass1 (heap.h:226-327)
//   pair_tasks_*_kernel   the distinct hits of the selected pairs -> full-SW tasks (hit_run_full_sw at half the
//                         full threshold, gmapper.c:2677), sw_full_ring.cu does the alignments
// Host (readpair_pass2 after the DP, :2207-2314): pair scores, insert sizes, readpair_remove_duplicate_hits
// with glibc qsort on 48-byte records like struct read_hit_pair, ranking, `saved` marks.
// Then the half-paired fall-back (:2612-2616): the saved marks go back to the device, pass 1 is replayed for
// every read with only_paired off (hits that kept a positive score are not rescored, :1296), the unpaired
// top-k and full SW run as in pipeline.cu, and read_pass2 finishes on the host.
#include <math.h>
#include <stdlib.h>
#include <cub/cub.cuh>
#include "chunk.cuh"

namespace shrimp {

int launch_pass1_replay(shrimp_gpu_ctx *ctx, const Pass1Params &P);
int launch_select_unpaired(shrimp_gpu_ctx *ctx, const Pass1Params &P);

struct PairParamsDev {
  MapParamsDev M;
  const DevHit *hits;
  const uint2 *rs_range;
  const int32_t *read_len;
  int n_pairs;
  int pair_mode, min_insert, max_insert;
  int32_t *pair_min, *pair_max;  // per hit slot, index into the partner strand's list
  int2 *pairsel;                 // [n_pairs][num_tmp_outputs] (slot of mate 0, slot of mate 1), heap-array order
  int32_t *pairkey;              // same shape, heap keys
  int32_t *n_pairsel;
};

// readpair_compute_mp_ranges (mapping.c:2317-2442): g_off deltas of read 1 per strand
__device__ __forceinline__ void mp_ranges_dev(int pair_mode, int mn, int mx, int rl1, int wl1, int rl2, int wl2,
                                              int dmin[2], int dmax[2]) {
  dmin[0] = mn - wl2;
  dmax[0] = mx + (wl1 - rl1) - rl2;
  dmin[1] = -mx + rl1 + (rl2 - wl2);
  dmax[1] = -mn + wl1;
  int add = 0;
  if (pair_mode == 2) add = rl1 + rl2;   // PAIR_OPP_OUT
  else if (pair_mode == 3) add = rl2;    // PAIR_COL_FW
  else if (pair_mode == 4) add = rl1;    // PAIR_COL_BW
  dmin[0] += add;
  dmax[0] += add;
  dmin[1] -= add;
  dmax[1] -= add;
}

// one thread per pair: readpair_pair_up_hits (mapping.c:266-325); pair_min/pair_max start at -1
__global__ void pair_up_kernel(const PairParamsDev P) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.n_pairs) return;
  const int r1 = 2 * k, r2 = 2 * k + 1;
  const int rl1 = P.read_len[r1], rl2 = P.read_len[r2];
  const int wl1 = (int)(unsigned short)abs_or_pct_d(P.M.window_len, P.M.window_len_frac, (double)rl1);
  const int wl2 = (int)(unsigned short)abs_or_pct_d(P.M.window_len, P.M.window_len_frac, (double)rl2);
  int dmin[2], dmax[2];
  mp_ranges_dev(P.pair_mode, P.min_insert, P.max_insert, rl1, wl1, rl2, wl2, dmin, dmax);
  for (int st1 = 0; st1 < 2; st1++) {
    const int st2 = 1 - st1;  // the reference pairs opposite strands in every pair mode
    const uint2 A = P.rs_range[2 * r1 + st1], B = P.rs_range[2 * r2 + st2];
    uint32_t j = 0;
    for (uint32_t i = 0; i < A.y; i++) {
      const DevHit h1 = P.hits[A.x + i];
      while (j < B.y) {
        const DevHit h2 = P.hits[B.x + j];
        if (h2.cn < h1.cn || (h2.cn == h1.cn && (long long)h2.g_off < (long long)h1.g_off + (long long)dmin[st1])) j++;
        else break;
      }
      uint32_t e = j;
      while (e < B.y) {
        const DevHit h2 = P.hits[B.x + e];
        if (h2.cn == h1.cn && (long long)h2.g_off <= (long long)h1.g_off + (long long)dmax[st1]) e++;
        else break;
      }
      if (j == e) continue;
      P.pair_min[A.x + i] = (int32_t)j;
      P.pair_max[A.x + i] = (int32_t)e - 1;
      for (uint32_t l = j; l < e; l++) {
        if (P.pair_min[B.x + l] < 0) P.pair_min[B.x + l] = (int32_t)i;
        P.pair_max[B.x + l] = (int32_t)i;
      }
    }
  }
}

// one thread per pair: readpair_get_vector_hits (mapping.c:1877-1932), first pass (nothing saved yet)
__global__ void select_pairs_kernel(const PairParamsDev P) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.n_pairs) return;
  const MapParamsDev &M = P.M;
  const int NT = M.num_tmp_outputs;
  const bool absolute = M.vect_thr < 0;
  int2 *a = P.pairsel + (size_t)k * NT;
  int32_t *key = P.pairkey + (size_t)k * NT;
  int load = 0;
  const int r1 = 2 * k, r2 = 2 * k + 1;
  for (int st1 = 0; st1 < 2; st1++) {
    const int st2 = 1 - st1;
    const uint2 A = P.rs_range[2 * r1 + st1], B = P.rs_range[2 * r2 + st2];
    for (uint32_t i = 0; i < A.y; i++) {
      const int s1 = (int)(A.x + i);
      const int pmin = P.pair_min[s1];
      if (pmin < 0) continue;
      const int pmax = P.pair_max[s1];
      const DevHit h1 = P.hits[s1];
      for (int j = pmin; j <= pmax; j++) {
        const int s2 = (int)B.x + j;
        const DevHit h2 = P.hits[s2];
        const int score = h1.score_vector + h2.score_vector;
        const int score_max = h1.score_max + h2.score_max;
        const int pct = (1000 * 100 * score) / score_max;
        const int kk = absolute ? score : pct;
        if (score >= (int)abs_or_pct_d(M.vect_thr, M.vect_frac, (double)score_max) && (load < NT || kk > key[0])) {
          if (load < NT) {  // extheap insert + percolate_up (heap.h:43-60)
            a[load] = make_int2(s1, s2);
            key[load] = kk;
            load++;
            int node = load, parent = node / 2;
            while (node > 1 && key[node - 1] < key[parent - 1]) {
              const int2 ta = a[parent - 1]; a[parent - 1] = a[node - 1]; a[node - 1] = ta;
              const int tk = key[parent - 1]; key[parent - 1] = key[node - 1]; key[node - 1] = tk;
              node = parent;
              parent = node / 2;
            }
          } else {  // replace_min + percolate_down (heap.h:62-89)
            a[0] = make_int2(s1, s2);
            key[0] = kk;
            int node = 1;
            for (;;) {
              int left = node * 2, right = left + 1, mn = node;
              if (left <= load && key[left - 1] < key[node - 1]) mn = left;
              if (right <= load && key[right - 1] < key[mn - 1]) mn = right;
              if (mn == node) break;
              const int2 ta = a[mn - 1]; a[mn - 1] = a[node - 1]; a[node - 1] = ta;
              const int tk = key[mn - 1]; key[mn - 1] = key[node - 1]; key[node - 1] = tk;
              node = mn;
            }
          }
        }
      }
    }
  }
  P.n_pairsel[k] = load;
}

struct PairTaskParams {
  FullBuildParams FB;      // tasks/info filled by the second pass
  const int2 *pairsel;
  const int32_t *n_pairsel;
  int n_pairs, NT;
  int32_t *taskof;         // per hit slot: ordinal of its task inside the pair, -1 = none (memset 0xff)
  int32_t *n_tasks;        // [n_pairs + 1] distinct hits per pair
  const int32_t *task_off; // exclusive scan of n_tasks
  int2 *pairtask;          // [n_pairs][NT] task ids of the two mates of every selected pair
  DevHit *hits_rw;
};

// one thread per pair: number the distinct hits of its selected pairs in heap-array order
__global__ void pair_tasks_count_kernel(const PairTaskParams P) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.n_pairs) return;
  const int2 *a = P.pairsel + (size_t)k * P.NT;
  int count = 0;
  for (int i = 0; i < P.n_pairsel[k]; i++) {
    const int s[2] = {a[i].x, a[i].y};
    for (int j = 0; j < 2; j++)
      if (P.taskof[s[j]] < 0) P.taskof[s[j]] = count++;
  }
  P.n_tasks[k] = count;
}

__global__ void pair_tasks_fill_kernel(const PairTaskParams P) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.n_pairs) return;
  const int2 *a = P.pairsel + (size_t)k * P.NT;
  const int base = P.task_off[k];
  int next = 0;
  for (int i = 0; i < P.n_pairsel[k]; i++) {
    const int s[2] = {a[i].x, a[i].y};
    int t[2];
    for (int j = 0; j < 2; j++) {
      const int ord = P.taskof[s[j]];
      t[j] = base + ord;
      if (ord == next) {  // first occurrence: hit_run_full_sw (mapping.c:2207-2216)
        next++;
        FullTask T;
        SelInfo I;
        make_full_task(P.FB, 2 * k + j, s[j], T, I);
        P.FB.tasks[t[j]] = T;
        P.FB.info[t[j]] = I;
        // letter space: hit_run_full_sw overwrites score_vector with the uncached sw_vector score (:386)
        if (!P.FB.M.colour_space) P.hits_rw[s[j]].score_vector = T.maxscore;
      }
    }
    P.pairtask[(size_t)k * P.NT + i] = make_int2(t[0], t[1]);
  }
}

// ---- host: readpair_pass2 after the DP ------------------------------------------------------------
struct HostPair {  // struct read_hit_pair (gmapper-definitions.h:155-165): 48 bytes, moved by value by qsort
  HostHit *rh[2];
  int rh_idx[2];
  int score_max, score, pct_score, key, insert_size;
  int improper_mapping;
};
static_assert(sizeof(HostPair) == 48, "HostPair must have the size of struct read_hit_pair");

struct PairHostCtx {
  const uint32_t *contig_len;
  int pair_mode;
};
static thread_local PairHostCtx g_pc;

static int sam_start(const HostHit *h) {  // hit_output / get_insert_size coordinate (mapping.c:405-456)
  if (h->info.gen_st == 0) return h->res.genome_start + 1;
  const int read_start = h->res.read_start + 1, read_end = read_start + h->res.rmapped - 1;
  const int right = (int)g_pc.contig_len[h->info.cn] - h->res.genome_start;
  return right - (read_end - read_start - h->res.deletions + h->res.insertions);
}
static int insert_size_of(const HostHit *rh, const HostHit *mp) {
  if (rh->info.cn != mp->info.cn) return 0;
  const int gs_mp = sam_start(mp), ge_mp = gs_mp + mp->res.gmapped - 1;
  const int gs = sam_start(rh), ge = gs + rh->res.gmapped - 1;
  const int fivep = rh->info.gen_st == 1 ? ge : gs - 1;
  const int fivep_mp = mp->info.gen_st == 1 ? ge_mp : gs_mp - 1;
  return fivep_mp - fivep;
}
static void compute_paired_hit(HostHit *h1, HostHit *h2, bool absolute, HostPair *d) {  // mapping.c:2053-2080
  d->rh[0] = h1;
  d->rh[1] = h2;
  d->score_max = h1->info.score_max + h2->info.score_max;
  d->score = h1->score_full + h2->score_full;
  d->pct_score = (1000 * 100 * d->score) / d->score_max;
  d->key = absolute ? d->score : d->pct_score;
  const int ins = insert_size_of(h1, h2);
  int sign;
  if (g_pc.pair_mode == 1 || g_pc.pair_mode == 3) sign = h1->info.gen_st == 0 ? 1 : -1;
  else sign = h1->info.gen_st == 1 ? 1 : -1;
  d->insert_size = sign * ins;
  d->improper_mapping = 0;
}
static int hcmp_start(const HostHit *a, const HostHit *b) {  // pass2_read_hit_sfrp_gen_start_cmp_base
  if (a->info.cn != b->info.cn) return a->info.cn - b->info.cn;
  if (a->info.gen_st != b->info.gen_st) return a->info.gen_st - b->info.gen_st;
  return a->res.genome_start - b->res.genome_start;
}
static int hcmp_end(const HostHit *a, const HostHit *b) {  // pass2_read_hit_sfrp_gen_end_cmp_base
  if (a->info.cn != b->info.cn) return a->info.cn - b->info.cn;
  if (a->info.gen_st != b->info.gen_st) return a->info.gen_st - b->info.gen_st;
  return (-a->res.genome_start - a->res.rmapped + a->res.deletions - a->res.insertions) -
         (-b->res.genome_start - b->res.rmapped + b->res.deletions - b->res.insertions);
}
static int pcmp_start0(const void *a, const void *b) { return hcmp_start(((const HostPair *)a)->rh[0], ((const HostPair *)b)->rh[0]); }
static int pcmp_end0(const void *a, const void *b) { return hcmp_end(((const HostPair *)a)->rh[0], ((const HostPair *)b)->rh[0]); }
static int pcmp_start1(const void *a, const void *b) { return hcmp_start(((const HostPair *)a)->rh[1], ((const HostPair *)b)->rh[1]); }
static int pcmp_end1(const void *a, const void *b) { return hcmp_end(((const HostPair *)a)->rh[1], ((const HostPair *)b)->rh[1]); }
static int pcmp_pointer(const void *a, const void *b) {  // pass2_readpair_pointer_cmp, mapping.c:1990-2030
  const HostPair *x = (const HostPair *)a, *y = (const HostPair *)b;
  if (x->rh[0]->info.sort_idx != y->rh[0]->info.sort_idx) return x->rh[0]->info.sort_idx - y->rh[0]->info.sort_idx;
  return x->rh[1]->info.sort_idx - y->rh[1]->info.sort_idx;
}
static int pcmp_score(const void *a, const void *b) { return ((const HostPair *)b)->key - ((const HostPair *)a)->key; }

static void push_dominant(HostPair *h, int n, bool absolute, int nip, int (*cmp)(const void *, const void *)) {
  // readpair_push_dominant_single_hits, mapping.c:2083-2110
  qsort(h, n, sizeof(h[0]), cmp);
  int i = 0;
  while (i < n) {
    int max = h[i].rh[nip]->score_full, max_idx = i, j = i + 1;
    while (j < n && !cmp(&h[i], &h[j])) {
      if (h[j].rh[nip]->score_full > max) {
        max = h[j].rh[nip]->score_full;
        max_idx = j;
      }
      j++;
    }
    for (int q = i; q < j; q++)
      if (q != max_idx) {
        h[q].rh[nip] = h[max_idx].rh[nip];
        compute_paired_hit(h[q].rh[0], h[q].rh[1], absolute, &h[q]);
      }
    i = j;
  }
}

}  // namespace shrimp

using namespace shrimp;

static int map_pairs_impl(shrimp_gpu_ctx *ctx, const shrimp_map_params *mp, const shrimp_pair_params *pp,
                          int n_pairs, const uint32_t *reads, int stride, const int32_t *read_len,
                          const int8_t *initbp, shrimp_hit *hits_out, int64_t hits_cap, int64_t *n_hits,
                          shrimp_pair *pairs_out, int64_t pairs_cap, int64_t *n_pairs_out,
                          int32_t *n_pairs_per_pair, int32_t *n_unpaired_per_read, uint8_t *edits,
                          int64_t edits_cap, int64_t *edits_used, shrimp_map_stats *stats, bool resident) {
  const char *who = resident ? "shrimp_gpu_map_pairs_resident" : "shrimp_gpu_map_pairs";
  if (!pp || !hits_out || !n_hits || !pairs_out || !n_pairs_out || n_pairs < 0) {
    set_error("%s: invalid argument", who);
    return SHRIMP_E_ARG;
  }
  if (pp->pair_mode < 1 || pp->pair_mode > 4) {
    set_error("%s: pair_mode must be 1 (opp-in) .. 4 (col-bw)", who);
    return SHRIMP_E_ARG;
  }
  if (mp && (mp->match_mode < 2 || mp->match_mode > 4)) {
    set_error("%s: paired match_mode must be 2, 3 or 4 (gmapper.c:2495-2499)", who);
    return SHRIMP_E_ARG;
  }
  Chunk C;
  SH_TRY(chunk_begin(C, ctx, mp, 2 * n_pairs, reads, stride, read_len, initbp, resident, who));
  // per-read options of the paired set (gmapper.c:2652-2677): region counts unless -n 2; the mate-pair region
  // counts by match mode and half-pairing; hit-list mode 2 / 3 / 1 and pass-1 min_matches 2 / 1 / 1 for -n 4 / 3 / 2
  const int pmm = mp->match_mode;
  C.M.match_mode = pmm == 4 ? 2 : pmm == 3 ? 3 : 1;
  C.M.min_matches = pmm == 4 ? 2 : 1;
  C.M.use_region_counts = (mp->use_regions && pmm != 2) ? 1 : 0;
  C.mp_mode = !mp->use_regions ? 0 : (pmm == 4 && !pp->half_paired) ? 1 : (pmm == 3 && pp->half_paired) ? 2 :
              (pmm == 3 && !pp->half_paired) ? 3 : 0;
  C.pair_mode = pp->pair_mode;
  C.M.rev_mate[0] = (pp->pair_mode == 2 || pp->pair_mode == 4) ? 1 : 0;   // pair_reverse, gmapper-defaults.h:184-191
  C.M.rev_mate[1] = (pp->pair_mode == 2 || pp->pair_mode == 3) ? 1 : 0;
  C.min_insert = pp->min_insert_size;
  C.max_insert = pp->max_insert_size;
  const int n_reads = C.n_reads;
  read_len = C.read_len;
  *n_hits = 0;
  *n_pairs_out = 0;
  if (edits_used) *edits_used = 0;
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n_pairs_per_pair) memset(n_pairs_per_pair, 0, sizeof(int32_t) * (size_t)n_pairs);
  if (n_unpaired_per_read) memset(n_unpaired_per_read, 0, sizeof(int32_t) * (size_t)n_reads);
  if (n_pairs == 0) return SHRIMP_OK;
  if (pairs_cap < (int64_t)n_pairs * mp->num_outputs || hits_cap < (int64_t)n_pairs * mp->num_outputs * 4) {
    set_error("%s: pairs_cap must be >= n_pairs * num_outputs and hits_cap >= 4 * n_pairs * num_outputs", who);
    return SHRIMP_E_ARG;
  }
  Pipeline *pl = C.pl;
  cudaStream_t st = ctx->stream;
  pl->d2h_bytes = 0;
  SH_TRY(chunk_scan(C));
  SH_TRY(chunk_vector(C));

  const int NT = mp->num_tmp_outputs;
  const size_t HU = std::max<uint32_t>(C.hits_used, 1);
  SH_TRY(pl->d_sel.ensure((size_t)n_reads * NT * 4));
  SH_TRY(pl->d_nsel.ensure(((size_t)n_reads + 1) * 4));
  SH_TRY(pl->d_pair_min.ensure(HU * 4));
  SH_TRY(pl->d_pair_max.ensure(HU * 4));
  SH_TRY(pl->d_taskof.ensure(HU * 4));
  SH_TRY(pl->d_saved.ensure(HU));
  SH_TRY(pl->d_pairsel.ensure((size_t)n_pairs * NT * (sizeof(int2) * 2 + 4)));
  SH_TRY(pl->d_npairsel.ensure(((size_t)n_pairs + 1) * 4 * 2));
  SH_TRY(pl->d_pairoff.ensure(((size_t)n_pairs + 1) * 4));
  SH_CUDA(cudaMemsetAsync(pl->d_pair_min.p, 0xff, HU * 4, st));
  SH_CUDA(cudaMemsetAsync(pl->d_pair_max.p, 0xff, HU * 4, st));
  SH_CUDA(cudaMemsetAsync(pl->d_taskof.p, 0xff, HU * 4, st));
  int2 *d_pairsel = pl->d_pairsel.as<int2>();
  int2 *d_pairtask = d_pairsel + (size_t)n_pairs * NT;
  int32_t *d_pairkey = (int32_t *)(d_pairtask + (size_t)n_pairs * NT);
  int32_t *d_npairsel = pl->d_npairsel.as<int32_t>();
  int32_t *d_ntasks = d_npairsel + (n_pairs + 1);

  // ---- pair-up, only_paired pass 1, pair top-k ---------------------------------------------------
  PairParamsDev PP;
  memset(&PP, 0, sizeof(PP));
  PP.M = C.M;
  PP.hits = pl->d_hits.as<DevHit>();
  PP.rs_range = pl->d_rs_range.as<uint2>();
  PP.read_len = pl->d_read_len.as<int32_t>();
  PP.n_pairs = n_pairs;
  PP.pair_mode = pp->pair_mode;
  PP.min_insert = pp->min_insert_size;
  PP.max_insert = pp->max_insert_size;
  PP.pair_min = pl->d_pair_min.as<int32_t>();
  PP.pair_max = pl->d_pair_max.as<int32_t>();
  PP.pairsel = d_pairsel;
  PP.pairkey = d_pairkey;
  PP.n_pairsel = d_npairsel;
  {
    ScopedStage ss(ctx, ST_PASS1);
    pair_up_kernel<<<(n_pairs + 127) / 128, 128, 0, st>>>(PP);
    SH_CUDA(cudaGetLastError());
    SH_LAUNCHED(ctx, ST_PASS1);
    Pass1Params P1 = chunk_pass1_params(C);
    P1.pair_min = pl->d_pair_min.as<int32_t>();
    SH_TRY(launch_pass1_replay(ctx, P1));
    select_pairs_kernel<<<(n_pairs + 127) / 128, 128, 0, st>>>(PP);
    SH_CUDA(cudaGetLastError());
    SH_LAUNCHED(ctx, ST_PASS1);
  }

  // ---- full SW on the distinct hits of the selected pairs (threshold: half the full one) -----------
  const double half_thr = mp->sw_full_threshold * 0.5;  // gmapper.c:2677
  int n_slots = 0;
  PairTaskParams PT;
  memset(&PT, 0, sizeof(PT));
  {
    ScopedStage ss(ctx, ST_FULL);
    PT.FB.G = C.G;
    PT.FB.M = C.M;
    PT.FB.M.full_thr = half_thr;
    PT.FB.M.full_frac = half_thr / 100.0;
    PT.FB.hits = pl->d_hits.as<DevHit>();
    PT.FB.rs_range = pl->d_rs_range.as<uint2>();
    PT.FB.read_len = pl->d_read_len.as<int32_t>();
    PT.FB.vtrue0 = pl->d_vtrue[0].as<int32_t>();
    PT.FB.initbp = C.cs ? pl->d_initbp.as<int8_t>() : nullptr;
    PT.FB.n_reads = n_reads;
    PT.pairsel = d_pairsel;
    PT.n_pairsel = d_npairsel;
    PT.n_pairs = n_pairs;
    PT.NT = NT;
    PT.taskof = pl->d_taskof.as<int32_t>();
    PT.n_tasks = d_ntasks;
    PT.task_off = pl->d_pairoff.as<int32_t>();
    PT.pairtask = d_pairtask;
    PT.hits_rw = pl->d_hits.as<DevHit>();
    pair_tasks_count_kernel<<<(n_pairs + 127) / 128, 128, 0, st>>>(PT);
    SH_CUDA(cudaGetLastError());
    SH_LAUNCHED(ctx, ST_FULL);
    size_t tmp_bytes = 0;
    SH_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_ntasks, pl->d_pairoff.as<int32_t>(), n_pairs + 1, st));
    SH_TRY(pl->d_scan_tmp.ensure(tmp_bytes));
    SH_CUDA(cudaMemsetAsync(d_ntasks + n_pairs, 0, 4, st));
    SH_CUDA(cub::DeviceScan::ExclusiveSum(pl->d_scan_tmp.p, tmp_bytes, d_ntasks, pl->d_pairoff.as<int32_t>(),
                                          n_pairs + 1, st));
    ctx->launches += 1;
    SH_CUDA(cudaMemcpyAsync(&n_slots, pl->d_pairoff.as<int32_t>() + n_pairs, 4, cudaMemcpyDeviceToHost, st));
    SH_CUDA(cudaStreamSynchronize(st));
    SH_TRY(pl->d_ftasks.ensure((size_t)std::max(n_slots, 1) * sizeof(FullTask)));
    SH_TRY(pl->d_finfo.ensure((size_t)std::max(n_slots, 1) * sizeof(SelInfo)));
    SH_TRY(pl->d_fresults.ensure((size_t)std::max(n_slots, 1) * sizeof(FullResult)));
    PT.FB.tasks = pl->d_ftasks.as<FullTask>();
    PT.FB.info = pl->d_finfo.as<SelInfo>();
    pair_tasks_fill_kernel<<<(n_pairs + 127) / 128, 128, 0, st>>>(PT);
    SH_CUDA(cudaGetLastError());
    SH_LAUNCHED(ctx, ST_FULL);
  }
  SH_TRY(chunk_run_full(C, n_slots));
  SH_TRY(chunk_fetch_full(C, n_slots, false));
  SH_TRY(pl->h_pairsel.ensure((size_t)n_pairs * NT * sizeof(int2)));
  SH_TRY(pl->h_npairsel.ensure(((size_t)n_pairs + 1) * 4 * 2));
  SH_CUDA(cudaMemcpyAsync(pl->h_pairsel.p, d_pairtask, (size_t)n_pairs * NT * sizeof(int2), cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaMemcpyAsync(pl->h_npairsel.p, d_npairsel, ((size_t)n_pairs + 1) * 4 * 2, cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaStreamSynchronize(st));
  pl->d2h_bytes += (size_t)n_pairs * NT * sizeof(int2) + ((size_t)n_pairs + 1) * 8;

  // ---- host: readpair_pass2 after the DP (mapping.c:2207-2314) -------------------------------------
  SH_TRY(pl->h_saved.ensure(HU));
  uint8_t *saved = pl->h_saved.as<uint8_t>();
  memset(saved, 0, HU);
  g_pc.contig_len = C.g->h_len.data();
  g_pc.pair_mode = pp->pair_mode;
  const int2 *PTASK = pl->h_pairsel.as<int2>();
  const int32_t *NPSEL = pl->h_npairsel.as<int32_t>();
  const int32_t *NTASK = NPSEL + (n_pairs + 1);
  const bool absolute = mp->sw_full_threshold < 0;
  HostOut O;
  memset(&O, 0, sizeof(O));
  O.hits = hits_out;
  O.hits_cap = hits_cap;
  O.edits = edits;
  O.edits_cap = edits_cap;
  int64_t n_po = 0;
  std::vector<HostHit> hh;
  std::vector<HostPair> h2((size_t)NT + 1);
  int task_base = 0;
  for (int k = 0; k < n_pairs; k++) {
    const int n1 = NPSEL[k], nt = NTASK[k];
    hh.resize((size_t)std::max(nt, 1));
    for (int t = 0; t < nt; t++) {
      host_score_hit(C, task_base + t, hh[t]);
      if (!C.cs) {
        O.pass2_vector_calls++;
        O.pass2_vector_cells += (uint64_t)hh[t].info.w_len * (uint64_t)read_len[hh[t].info.read_idx];
      }
      if (hh[t].res.score > 0 || hh[t].res.ops_len > 0) O.full_calls++;
    }
    int n2 = 0;
    for (int i = 0; i < n1; i++) {
      const int2 tt = PTASK[(size_t)k * NT + i];
      HostHit *a = &hh[tt.x - task_base], *b = &hh[tt.y - task_base];
      if (a->score_full == 0 || b->score_full == 0) continue;
      const int smax = a->info.score_max + b->info.score_max;
      const double thr = absolute ? -mp->sw_full_threshold : smax * (mp->sw_full_threshold / 100.0);
      if (a->score_full + b->score_full >= (int)thr) compute_paired_hit(a, b, absolute, &h2[n2++]);
    }
    // readpair_remove_duplicate_hits, mapping.c:2114-2175
    push_dominant(h2.data(), n2, absolute, 0, pcmp_start0);
    push_dominant(h2.data(), n2, absolute, 0, pcmp_end0);
    push_dominant(h2.data(), n2, absolute, 1, pcmp_start1);
    push_dominant(h2.data(), n2, absolute, 1, pcmp_end1);
    qsort(h2.data(), n2, sizeof(HostPair), pcmp_pointer);
    {  // removedups, util.c:1239-1255
      int m = 0, i = 0;
      while (i < n2) {
        int j = i + 1;
        while (j < n2 && !pcmp_pointer(&h2[i], &h2[j])) j++;
        if (m < i) h2[m] = h2[i];
        m++;
        i = j;
      }
      n2 = m;
    }
    qsort(h2.data(), n2, sizeof(HostPair), pcmp_score);
    if (n2 > mp->num_outputs) n2 = mp->num_outputs;
    if (mp->strata && n2 > 0) {
      int i;
      for (i = 1; i < n2 && h2[0].score == h2[i].score; i++)
        ;
      n2 = i;
    }
    if (n2 > 0 && !(mp->max_alignments == 0 || n2 <= mp->max_alignments)) n2 = 0;
    for (int i = 0; i < n2; i++) {
      HostPair &p = h2[i];
      p.rh[0]->pass2_key = p.rh[1]->pass2_key = 0;
      shrimp_pair &o = pairs_out[n_po++];
      o.pair_idx = k;
      o.score = p.score;
      o.score_max = p.score_max;
      o.key = p.key;
      o.insert_size = p.insert_size;
      o.hit_idx[0] = (int32_t)O.n_out;
      host_fill_hit(C, *p.rh[0], 2 * k, O);
      o.hit_idx[1] = (int32_t)O.n_out;
      host_fill_hit(C, *p.rh[1], 2 * k + 1, O);
      saved[p.rh[0]->info.hit_slot] = 1;
      saved[p.rh[1]->info.hit_slot] = 1;
    }
    if (n_pairs_per_pair) n_pairs_per_pair[k] = n2;
    task_base += nt;
  }
  *n_pairs_out = n_po;
  uint32_t cntA[64];
  memcpy(cntA, (const char *)pl->h_nsel.p + (size_t)n_reads * 4, sizeof(cntA));

  const uint32_t *h_cnt = (const uint32_t *)((char *)pl->h_nsel.p + (size_t)n_reads * 4);
  if (pp->half_paired) {
    // ---- half-paired fall-back: handle_read for every mate, pass 1 + pass 2 only (gmapper.c:2694-2714;
    // stop_threshold 101 % means no pair ever stops it, :2686-2687) ----
    SH_CUDA(cudaMemcpyAsync(pl->d_saved.p, saved, HU, cudaMemcpyHostToDevice, st));
    pl->h2d_bytes += HU;
    {
      ScopedStage ss(ctx, ST_PASS1);
      Pass1Params P1 = chunk_pass1_params(C);
      P1.M.min_matches = 2;   // unpaired_mapping_options[.][0].pass1.min_matches, gmapper.c:2705
      P1.saved = pl->d_saved.as<uint8_t>();
      SH_TRY(launch_pass1_replay(ctx, P1));
      SH_TRY(launch_select_unpaired(ctx, P1));
    }
    int n_slots_b = 0;
    SH_TRY(chunk_full_tasks_unpaired(C, mp->sw_full_threshold, &n_slots_b));
    SH_TRY(chunk_run_full(C, n_slots_b));
    SH_TRY(chunk_fetch_full_packed(C, n_slots_b));
    const int32_t *NSEL = pl->h_nsel.as<int32_t>();
    SH_TRY(host_pass2_all(C, NSEL, mp->sw_full_threshold, O, n_unpaired_per_read));
  }
  *n_hits = O.n_out;
  if (edits_used) *edits_used = O.e_used;
  if (stats) {
    chunk_stats(C, h_cnt, stats);
    stats->vector_calls += O.pass2_vector_calls;
    stats->vector_cells += O.pass2_vector_cells;
    stats->full_calls = O.full_calls;
    (void)cntA;
  }
  if (O.hits_short) {
    set_error("%s: hits_cap too small", who);
    return SHRIMP_E_NOMEM;
  }
  if (O.edits_short && edits) {
    set_error("%s: edits_cap too small, %lld bytes needed", who, (long long)O.e_used);
    return SHRIMP_E_NOMEM;
  }
  return SHRIMP_OK;
}

extern "C" int shrimp_gpu_map_pairs(shrimp_gpu_ctx *ctx, const shrimp_map_params *mp, const shrimp_pair_params *pp,
                                    int n_pairs, const uint32_t *reads, int stride, const int32_t *read_len,
                                    const int8_t *initbp, shrimp_hit *hits_out, int64_t hits_cap, int64_t *n_hits,
                                    shrimp_pair *pairs_out, int64_t pairs_cap, int64_t *n_pairs_out,
                                    int32_t *n_pairs_per_pair, int32_t *n_unpaired_per_read, uint8_t *edits,
                                    int64_t edits_cap, int64_t *edits_used, shrimp_map_stats *stats) {
  return map_pairs_impl(ctx, mp, pp, n_pairs, reads, stride, read_len, initbp, hits_out, hits_cap, n_hits, pairs_out,
                        pairs_cap, n_pairs_out, n_pairs_per_pair, n_unpaired_per_read, edits, edits_cap, edits_used,
                        stats, false);
}

// Measurement entry: maps again the pairs the previous shrimp_gpu_map_pairs call left in HBM (no upload of the
// reads); the records go to library-owned scratch and only the statistics come back.  The mid-pipeline host
// stage (readpair_pass2) is part of the path and stays inside.
extern "C" int shrimp_gpu_map_pairs_resident(shrimp_gpu_ctx *ctx, const shrimp_map_params *mp,
                                             const shrimp_pair_params *pp, shrimp_map_stats *stats) {
  Pipeline *pl = ctx ? (Pipeline *)ctx->pipeline : nullptr;
  if (!pl || pl->res_n_reads <= 0 || (pl->res_n_reads & 1) || !mp) {
    set_error("shrimp_gpu_map_pairs_resident: no pairs resident; call shrimp_gpu_map_pairs first");
    return SHRIMP_E_STATE;
  }
  const int n_pairs = pl->res_n_reads / 2;
  static thread_local std::vector<shrimp_hit> hits;
  static thread_local std::vector<shrimp_pair> pairs;
  hits.resize((size_t)n_pairs * mp->num_outputs * 4);
  pairs.resize((size_t)n_pairs * mp->num_outputs);
  int64_t nh = 0, np = 0, eu = 0;
  return map_pairs_impl(ctx, mp, pp, n_pairs, nullptr, 0, nullptr, nullptr, hits.data(), (int64_t)hits.size(), &nh,
                        pairs.data(), (int64_t)pairs.size(), &np, nullptr, nullptr, nullptr, 0, &eu, stats, true);
}
