// Parameter blocks of the pipeline stage kernels (scan.cu, pass1.cu, sw_full.cu), shared with the
// host orchestration in pipeline.cu.
#pragma once
#include "pipeline.cuh"

namespace shrimp {

#define SCAN_WARPS 8
#define SCAN_WARPS_HOST SCAN_WARPS

struct AnchorRec {
  uint32_t x;      // global position
  int32_t cn;
  int16_t y, len;  // one 32-bit word, len in the upper half (the CTA kernel's collapse does atomicMax on it)
  int32_t weight;
};

struct ScanParams {
  GenomeView G;
  IndexView I;
  SeedTable S;
  MapParamsDev M;
  const uint32_t *reads;  // [2 * n_reads][stride], row 2*r + st
  int stride;
  int n_reads;
  const int32_t *read_len;
  // work list: nullptr = all read strands, else explicit list (overflow pass)
  const uint32_t *work;
  uint32_t n_work;
  // outputs
  DevHit *hits;
  uint32_t hits_cap;
  uint32_t *hits_used;     // atomic cursor
  uint2 *rs_range;         // [2 * n_reads] (first hit, count)
  uint32_t *overflow;      // list of read strands that did not fit `cap`
  uint32_t *n_overflow;
  uint32_t *status;        // bit 0: hits_cap exhausted, bit 1: slab exhausted in the overflow pass
  uint32_t *stats;         // [0] heap replays
  unsigned long long *stats64;   // [0] gathered entries, [1] surviving entries, [2] anchors
  // per-warp global scratch for the heap replay: [n_warps][scratch_ints]
  int32_t *scratch;
  int scratch_ints;
  int k_max;               // n_seeds * max_n_kmers upper bound
  int cap;                 // candidate slots (after the bitmap filter) per warp / per CTA
  int max_rl;
  int k_cap;               // k-mers per read strand the shared-memory tables hold
  int bm_log2;             // log2 of the bits of each region bitmap
  int alias_rec;           // warp kernel: the anchors reuse the bitmaps (dead after pass B)
  int stream;              // warp kernel: lists long enough for one contiguous stream per lane
  int stash;               // warp kernel: list entries staged per strand in shared memory (0 = off)
  // CTA kernel (scan_cta_kernel)
  int n_part;              // genome partitions of 2^bm_log2 regions each (exact region bitmaps)
  uint32_t *work_counter;  // dynamic work distribution
  unsigned long long *g_ent;   // != nullptr: candidate arrays in a global slab per CTA of g_cap entries
  AnchorRec *g_rec;
  uint16_t *g_order;
  uint32_t *g_keep;
  int g_cap;
  // strands with equal positions on different read offsets: parked for scan_replay_kernel, resumed at step 6
  unsigned long long *tie_ent;   // candidate slab
  uint16_t *tie_order;           // pop order per parked strand (same offsets as tie_ent)
  uint4 *tie_rec;                // (read strand, slab offset, candidates, 0)
  uint32_t *tie_used, *n_tie;    // atomic cursors
  uint32_t tie_cap, tie_rec_cap;
  int resume, resume_min;        // resume launch: strands with resume_min < candidates <= cap
  unsigned long long *prof;      // optional phase cycle counters [16] (SHRIMP_SCAN_PROF)
  // mate-pair region counts (paired option sets that look at the mate)
  int mp_mode;                   // anchor_list.use_mp_region_counts: 0 off, 1, 2, 3 (gmapper.c:2659-2662); work item = pair
  int pair_mode, min_insert, max_insert;
  uint32_t *mp_tab;              // [n_ctas][4][mp_regions] region tables, zeroed once
  uint32_t *mp_epoch;            // [n_ctas] table epochs
  int mp_regions;
  int bm_hashed;                 // CTA kernel: hashed region bitmaps (one partition) instead of exact partitions
  int win;                       // entries of the shared-memory staging window
  int lanes_per_list_log2; // lanes that share one index list (2..5): short lists are streamed several per warp
  int walk;                // CTA kernel: cursor walk over the lists (no staging window, no partition cuts)
  int sort_shift;          // CTA kernel: candidates are binned by position >> sort_shift (64 bins) before the sort
};

struct TaskBuildParams {
  GenomeView G;
  MapParamsDev M;
  const DevHit *hits;
  const uint2 *rs_range;
  const int32_t *read_len;
  int n_reads;
  // dense task arrays, one set per genome orientation (colour space reverses the
  // strand-1 hits onto the reverse-complement genome, mapping.c:1303-1312)
  uint32_t *goff[2];
  int32_t *glen[2];
  int32_t *ridx[2];
  int32_t *rlen[2];
  int8_t *initbp_out[2];
  const int8_t *initbp;   // per read (colour space)
  uint32_t *out[2];       // hit slot of every dense task
  uint32_t *slot;         // f1 cache slot per hit (hash_filter_calls)
  uint32_t *task_stats;   // [0] eligible windows, [2..3] their cells (u64), [4], [5] dense task count per orientation
};

struct GaplessParams {
  GenomeView G;
  const DevHit *hits;
  const uint32_t *reads;
  int stride;
  const uint32_t *out;     // hit slot per dense task
  const int32_t *ridx, *rlen;
  const int8_t *initbp;    // colour space: per dense task
  uint32_t n_tasks;
  int match, mismatch;
  int cs, ori;             // colour space; orientation of this launch's tasks (1: reversed onto the rc arrays)
  int32_t *scores;         // per hit slot
};

struct Pass1Params {
  MapParamsDev M;
  DevHit *hits;
  const uint2 *rs_range;
  const int32_t *read_len;
  int n_reads;
  const int32_t *vtrue[2];   // true sw_vector scores per hit slot (per orientation launch)
  const uint32_t *slot;
  uint8_t *writer;           // scratch flag per hit slot
  int32_t *sel;              // [n_reads][num_tmp_outputs] hit slots in heap-array order
  int32_t *n_sel;            // [n_reads]
  uint32_t *stats;           // [4] vector calls the reference would make, [5] bypassed, [6] cells (lo), [7] cells (hi)
  const int32_t *pair_min;   // read pairs, only_paired pass: first partner index per hit slot (-1 = none); else nullptr
  const uint8_t *saved;      // half-paired second pass: hits kept by the paired pass 2; else nullptr
};

struct FullTask {
  uint32_t goff_global;  // window start in the chosen orientation array (global nibble coordinate)
  uint32_t goff_contig;  // same, relative to the contig (what sw_full_ls receives as goff)
  int32_t glen, rlen;
  int32_t ridx;          // row of the read in the reads array
  int32_t ax, ay, alen, awidth;
  int32_t thresh, maxscore;
  int32_t gen_st;        // 0 forward genome, 1 reverse-complement genome
  int32_t run;           // 0: below threshold, result score = 0 without DP (mapping.c:390-398)
  int32_t initbp;        // colour space: initial base of the read
};

struct FullResult {
  int32_t score;
  int32_t read_start, rmapped, genome_start, gmapped;
  int32_t matches, mismatches, insertions, deletions, crossovers;
  int32_t ops_start, ops_len;   // into this task's ops column
  double posterior;             // colour space with mapping qualities: sfrp->posterior of post_sw (post_sw.cu)
  // ... and the score hit_run_post_sw derives from it, (int)rint(alpha log2(posterior) + rmapped (2 alpha + beta))
  // (mapping.c:1619-1621), taken on the device with the same libm logarithm: one log() per alignment less on the host
  int32_t post_score, pad_;
};

struct FullParams {
  const uint32_t *genome_fwd, *genome_rc;
  const uint32_t *reads;
  int stride;
  const FullTask *tasks;
  FullResult *results;
  int n_tasks;      // in this launch
  int NT;           // scratch stride (>= n_tasks)
  int32_t *row;     // [3][max_glen+1][NT]
  uint8_t *bp;      // [max_rlen*max_glen][NT]
  uint8_t *ops;     // [NT][max_rlen+max_glen]
  int max_glen, max_rlen;
  int match, mismatch, a_open, a_ext, b_open, b_ext;
  int anchor_width, Tflag, local;
  unsigned long long *cells;
  // colour space (sw_full_cs.cu)
  int xover, indel_taboo_len;
  const int16_t *xover_pos;   // per read position crossover scores [n_reads][xover_stride] (reads with qualities), or nullptr
  int xover_stride;
  int32_t *row_cs;  // [12][max_glen+1][NT]
  uint8_t *bp_cs;   // [max_rlen*max_glen*12][NT]
  // band-ring kernels (sw_full_ring.cu): task ids of this launch (nullptr = identity), ring width,
  // packed colour-space back-pointers [max_rlen*W][NT]; `bp` is [max_rlen*W][NT] there
  const int32_t *perm;
  int rev;          // every task of this launch has gen_st && Tflag == rev (the tasks are grouped by it)
  int W;
  unsigned long long *bp64;
};

// post_sw over the full-SW tasks of a chunk (post_sw.cu)
struct PostParams {
  const uint32_t *genome_fwd, *genome_rc;
  const uint32_t *reads;
  int stride;
  const FullTask *tasks;
  FullResult *results;
  uint8_t *ops;            // edit scripts, rewritten with the corrected base calls
  int ops_stride;
  uint8_t *quals_out;      // [n_tasks][max_rlen] base qualities (33 + q) of the aligned read columns
  int max_rlen;
  int n_tasks;
  const uint8_t *read_quals;   // quality strings of the reads (gmapper -Q) or nullptr
  int qual_stride, qual_vector_offset;
  unsigned long long *columns;       // statistics: aligned read columns processed
  const unsigned long long *gm_tab;  // glibc exp/log tables (glibc_math.cuh): exp consts[8], exp tab[256], log consts[18], log tab[256]
  const double *lc1_tab, *lc2_tab;   // per quality character: log(1 - colour error rate), log(rate / 3) (host libm)
  double la1, la2;         // log(1 - pr_snp), log(pr_snp / 3)
  double lc1, lc2;         // reads without qualities: log(1 - pr_xover), log(pr_xover / 3)
  double ln1, ln2;         // colour N: rate .75
  double pr_del_open, pr_del_extend, pr_ins_open, pr_ins_extend;
  double score_alpha, score_2ab, log2v;   // alpha, 2 alpha + beta, the host's log(2.0)
};

}  // namespace shrimp
