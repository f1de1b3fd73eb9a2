// Chunk state shared by the unpaired (pipeline.cu) and paired (pairs.cu) mapping entries.
#pragma once
#include <algorithm>
#include "band.cuh"

namespace shrimp {

struct SelInfo {
  int32_t hit_slot, read_idx, st, cn, gen_st, w_len;
  int32_t sort_idx;  // read_hit::sort_idx (mapping.c:1243-1246): index in the read's hit lists, strand 0 first
  uint32_t g_off;  // oriented (after reverse_hit)
  int32_t score_vector, score_max, matches;
  int32_t wg;  // read_hit::score_window_gen
};

struct FullBuildParams {
  GenomeView G;
  MapParamsDev M;
  const DevHit *hits;
  const uint2 *rs_range;
  const int32_t *read_len;
  const int32_t *sel;
  const int32_t *n_sel;
  const int32_t *vtrue0;
  const int8_t *initbp;
  const int32_t *task_off;  // exclusive prefix sum of n_sel: tasks of read r start at task_off[r]
  int n_reads;
  FullTask *tasks;   // [sum n_sel], dense
  SelInfo *info;
};

// hit_run_full_sw's orientation logic for one selected hit (mapping.c:353-361, reverse_hit :254-263,
// anchor_reverse anchors.h:30-34) -> the full-SW task and the bookkeeping record that goes back to the host
__device__ __forceinline__ void make_full_task(const FullBuildParams &P, int r, int hi, FullTask &T, SelInfo &I) {
  memset(&T, 0, sizeof(T));
  memset(&I, 0, sizeof(I));
  const DevHit h = P.hits[hi];
  const uint2 r0 = P.rs_range[2 * r], r1 = P.rs_range[2 * r + 1];
  const int st = ((uint32_t)hi >= r0.x && (uint32_t)hi < r0.x + r0.y) ? 0 : 1;
  const int rl = P.read_len[r];
  const uint32_t coff = P.G.contig_off[h.cn], clen = P.G.contig_len[h.cn];
  uint32_t g_off = h.g_off;
  int ax = h.ax, ay = h.ay;
  int gen_st = 0;
  const int in_st = P.M.rev_mate[r & 1];   // re->input_strand
  if (st != in_st) {  // reverse_hit: the read is always aligned in its input orientation
    g_off = clen - h.g_off - (uint32_t)h.w_len;
    ax = -h.ax + (h.w_len - 1) - (h.alen - 1) - (h.awidth - 1);
    ay = -h.ay + (rl - 1) - (h.alen - 1) + (h.awidth - 1);
    gen_st = 1;
  }
  T.goff_global = coff + g_off;
  T.goff_contig = g_off;
  T.glen = h.w_len;
  T.rlen = rl;
  T.ridx = 2 * r + in_st;
  T.ax = ax;
  T.ay = ay;
  T.alen = h.alen;
  T.awidth = h.awidth;
  T.thresh = (int)abs_or_pct_d(P.M.full_thr, P.M.full_frac, (double)h.score_max);
  T.gen_st = gen_st;
  if (!P.M.colour_space) {
    T.maxscore = P.vtrue0[hi];  // sw_vector re-run of mapping.c:386 (same score on the flipped window)
    T.run = T.maxscore >= T.thresh;
  } else {
    T.maxscore = h.score_vector;
    T.run = 1;
    T.initbp = P.initbp[r];
  }
  I.hit_slot = hi;
  I.read_idx = r;
  I.st = st;
  I.cn = h.cn;
  I.gen_st = gen_st;
  I.w_len = h.w_len;
  I.sort_idx = st == 0 ? (int)((uint32_t)hi - r0.x) : (int)(r0.y + ((uint32_t)hi - r1.x));
  I.g_off = g_off;
  I.score_vector = P.M.colour_space ? h.score_vector : T.maxscore;
  I.score_max = h.score_max;
  I.matches = h.matches;
  I.wg = h.wg;
}

#define RING_CLASSES 4

struct Pipeline {
  DevBuf d_in, d_reads, d_read_len, d_initbp, d_hits, d_rs_range, d_counters, d_overflow, d_overflow2, d_scan_slab, d_tie_ent, d_tie_order, d_tie_rec, d_prof, d_mp_tab, d_mp_epoch, d_xover, d_quals, d_fqual, d_pstab, d_gmtab, d_scratch;
  DevBuf d_task[2], d_vtrue[2], d_slot, d_writer, d_sel, d_nsel;
  DevBuf d_pk, d_pk_info, d_pk_res, d_pk_pool;
  DevBuf d_ftasks, d_finfo, d_fresults, d_frow, d_fbp[RING_CLASSES + 2], d_fops, d_taskoff, d_scan_tmp, d_perm;
  HostBuf h_info, h_results, h_ops, h_nsel, h_hits, h_range, h_xover, h_fqual;
  bool gm_tab_ready = false;
  int qual_stride = 0;      // quality strings of the resident chunk (0 = none)
  int xover_stride = 0;     // per-position crossover scores of the resident chunk (0 = none)
  uint32_t hits_cap = 0, tie_cap = 0;
  size_t mp_tab_ints = 0;   // size the region tables were zeroed for
  // reads left resident by the last upload (shrimp_gpu_map_resident)
  int res_n_reads = 0, res_stride = 0;
  std::vector<int32_t> res_read_len;
  size_t h2d_bytes = 0, d2h_bytes = 0;
  // paired mapping (pairs.cu)
  DevBuf d_pair_min, d_pair_max, d_saved, d_pairsel, d_npairsel, d_taskof, d_pairoff;
  HostBuf h_pairsel, h_npairsel, h_saved;
};

struct HostHit {
  SelInfo info;
  FullResult res;
  int score_full;
  double pct_score_full;
  int pass2_key;
  double posterior;
  int task_idx;
};


// One chunk of reads on its way through the device stages (pipeline.cu).
struct Chunk {
  shrimp_gpu_ctx *ctx = nullptr;
  Pipeline *pl = nullptr;
  DeviceGenome *g = nullptr;
  const shrimp_map_params *mp = nullptr;
  MapParamsDev M;
  GenomeView G;
  IndexView IV;
  int n_reads = 0, stride = 0, max_rl = 0, max_wl = 0;
  long long sum_rl = 0;
  bool cs = false;
  const int32_t *read_len = nullptr;   // host
  uint32_t *cnt = nullptr;             // device counters [64]
  uint32_t hits_used = 0;
  uint32_t scan_big = 0;               // read strands that went through scan_cta_kernel
  uint32_t scan_global = 0;            // ... of which with their candidate arrays in global slabs
  int n_ori = 1;
  // paired option sets that look at the mate's region counts (pairs.cu sets these; 0 = off)
  int mp_mode = 0, pair_mode = 0, min_insert = 0, max_insert = 0;
  size_t ops_stride = 0;
  bool post_sw = false;                // colour space with mapping qualities: post_sw runs after the full SW
  bool packed = false;                 // the host buffers hold the packed records of chunk_fetch_full_packed
  int packed_slots = 0;
  int64_t packed_kept = 0;
};

// stages (pipeline.cu)
int chunk_begin(Chunk &C, shrimp_gpu_ctx *ctx, const shrimp_map_params *mp, int n_reads, const uint32_t *reads,
                int stride, const int32_t *read_len, const int8_t *initbp, bool resident, const char *who);
int chunk_scan(Chunk &C);
int chunk_vector(Chunk &C);
Pass1Params chunk_pass1_params(Chunk &C);
void host_score_hit(const Chunk &C, int idx, HostHit &h);
int chunk_full_tasks_unpaired(Chunk &C, double full_thr, int *n_slots);
int chunk_run_full(Chunk &C, int n_slots);
int chunk_fetch_full(Chunk &C, int n_slots, bool with_nsel);
int chunk_fetch_full_packed(Chunk &C, int n_slots);
void chunk_stats(const Chunk &C, const uint32_t *h_cnt, shrimp_map_stats *stats);
// read_pass2 after the DP (mapping.c:1644-1722) for one read: tasks [task_base, task_base + n1) -> out
struct HostOut {
  shrimp_hit *hits;
  int64_t hits_cap, n_out;
  uint8_t *edits;
  int64_t edits_cap, e_used;
  bool edits_short, hits_short;
  uint64_t full_calls, pass2_vector_calls, pass2_vector_cells;
};
int host_pass2_all(const Chunk &C, const int32_t *n_sel, double full_thr, HostOut &O, int32_t *n_per_read);
void host_fill_hit(const Chunk &C, const HostHit &h, int r, HostOut &O);
int run_full_sw(shrimp_gpu_ctx *ctx, DevBuf &d_perm, DevBuf &d_row, DevBuf *d_bp, FullParams FP, int n, bool cs,
                uint32_t *d_cls_count);
void free_pipeline(shrimp_gpu_ctx *ctx);

}  // namespace shrimp
