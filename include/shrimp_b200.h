/*
 * shrimp_b200.h -- C ABI of the B200-native gmapper hot path.
 *
 * Plain C, POD arguments only (pointers + sizes); no CUDA/torch types cross this boundary.
 * Every entry point names the reference interface (file:line under compbio-UofT/shrimp 2.2.3)
 * it replaces.  All pointers are HOST pointers unless the name says "_dev"; the library does its
 * own H2D/D2H.  All functions return 0 on success and a negative SHRIMP_E_* code on failure;
 * shrimp_gpu_last_error() returns the message of the last failure on the calling thread.
 *
 * Data conventions are the reference's own (common/util.h:41-42): sequences are 4 bits per
 * base, 8 bases per uint32_t, base i in bits 4*(i%8) of word i/8; codes are fasta.h:26-42
 * (A,C,G,T = 0..3, N = 15); colour-space arrays hold colours 0..3 (N = 15).  Gap/mismatch
 * scores are passed with the reference's CLI sign (negative penalties), exactly as to
 * sw_vector_setup().
 *
 * There is no CPU fallback: with no usable CUDA device every call fails with SHRIMP_E_CUDA.
 */
#ifndef SHRIMP_B200_H
#define SHRIMP_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHRIMP_OK          0
#define SHRIMP_E_CUDA     -1   /* CUDA runtime / no device */
#define SHRIMP_E_ARG      -2   /* invalid argument */
#define SHRIMP_E_STATE    -3   /* call order (e.g. map before index build) */
#define SHRIMP_E_RANGE    -4   /* value outside what the int16 / packed kernels support */
#define SHRIMP_E_NOMEM    -5

typedef struct shrimp_gpu_ctx shrimp_gpu_ctx;

/* ------------------------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------------------------- */
int          shrimp_gpu_device_count(void);
int          shrimp_gpu_create(int device, shrimp_gpu_ctx **out);
void         shrimp_gpu_destroy(shrimp_gpu_ctx *ctx);
const char  *shrimp_gpu_last_error(void);
/* Number of kernels this library launched on ctx since creation (bench.py "gpu_launches"). */
uint64_t     shrimp_gpu_launch_count(const shrimp_gpu_ctx *ctx);
/* Device-side milliseconds (CUDA events on the library's stream) spent in each pipeline stage
 * since the last reset; names[] receives static strings.  Returns the number of stages. */
int          shrimp_gpu_stage_times(shrimp_gpu_ctx *ctx, const char **names, float *ms, uint64_t *launches, int max_n);
void         shrimp_gpu_stage_times_reset(shrimp_gpu_ctx *ctx);

/* ------------------------------------------------------------------------------------------
 * Scoring set-up.  Replaces sw_vector_setup (common/sw-vector.c:388-439), sw_gapless_setup
 * (sw-gapless.c:28-43), sw_full_ls_setup (sw-full-ls.c:573-623) and sw_full_cs_setup
 * (sw-full-cs.c:1076-1132): one parameter block for every kernel of the context.
 * ---------------------------------------------------------------------------------------- */
typedef struct shrimp_sw_params {
  int match;            /* > 0 */
  int mismatch;         /* < 0: mismatch_score as sw_full_*_setup get it.  In colour space the vector
                           filter scores a colour mismatch as match + crossover (gmapper.c:2935). */
  int a_gap_open;       /* <= 0, gap that consumes genome ("a", SAM D) */
  int a_gap_ext;
  int b_gap_open;       /* <= 0, gap that consumes read ("b", SAM I) */
  int b_gap_ext;
  int crossover;        /* < 0, colour space only (sw_full_cs) */
  int use_colours;      /* 0 letter space, 1 colour space */
  int anchor_width;     /* sw_full_*: band half-width around the anchor; < 0 = threshold band */
  int indel_taboo_len;  /* sw_full_cs */
  int max_read_len;     /* qrlen of the *_setup calls; match*max_read_len must be < 32768 */
  int max_window_len;   /* dblen of the *_setup calls */
} shrimp_sw_params;

int shrimp_gpu_sw_setup(shrimp_gpu_ctx *ctx, const shrimp_sw_params *p);

/* ------------------------------------------------------------------------------------------
 * Batched vector Smith-Waterman filter.  Replaces sw_vector (common/sw-vector.c:453-515), one
 * call per task: score-only affine-gap local alignment of read (rows) against a genome window
 * (columns).  Task t scores reads[read_idx[t]] (rlen[t] bases) against genome[goff[t] ..
 * goff[t]+glen[t]).  Colour space (use_colours): genome is the colour genome, genome_ls the
 * letter genome (same coordinates) and initbp[t] the read's initial base; row 0 then follows
 * sw-vector.c:116-146.  Letter space: genome_ls = NULL, initbp = NULL.
 * ---------------------------------------------------------------------------------------- */
int shrimp_gpu_sw_vector_batch(shrimp_gpu_ctx *ctx,
                               const uint32_t *genome, size_t genome_words,
                               const uint32_t *genome_ls,
                               const uint32_t *reads, int read_stride_words, int n_reads,
                               int n_tasks,
                               const uint32_t *goff, const int32_t *glen,
                               const int32_t *read_idx, const int32_t *rlen,
                               const int8_t *initbp,
                               int32_t *scores_out);

/* ------------------------------------------------------------------------------------------
 * Batched gapless filter.  Replaces sw_gapless (common/sw-gapless.c:57-117), one call per task: the best
 * ungapped segment score on the diagonal through (g_idx[t], r_idx[t]) of the genome piece genome[goff[t] ..
 * goff[t] + glen[t]) (f1_run passes a whole contig, f1-wrapper.h:121-124) against reads[read_idx[t]].  Scores are
 * the match / vector mismatch of the set-up (sw_gapless_setup gets the same two, f1-wrapper.h:56).  Colour space:
 * genome_ls + initbp[t] force the first colour of the read (:83-93); letter space: NULL, NULL.
 * ---------------------------------------------------------------------------------------- */
int shrimp_gpu_sw_gapless_batch(shrimp_gpu_ctx *ctx,
                                const uint32_t *genome, size_t genome_words,
                                const uint32_t *genome_ls,
                                const uint32_t *reads, int read_stride_words, int n_reads,
                                int n_tasks,
                                const uint32_t *goff, const int32_t *glen,
                                const int32_t *read_idx, const int32_t *rlen,
                                const int32_t *g_idx, const int32_t *r_idx,
                                const int8_t *initbp,
                                int32_t *scores_out);

/* ------------------------------------------------------------------------------------------
 * Batched full Smith-Waterman with traceback.  Replaces sw_full_ls (common/sw-full-ls.c:637-683) and,
 * after a colour-space set-up, sw_full_cs (common/sw-full-cs.c:1146-1236), one task per call the
 * reference would make: read reads[read_idx] (rlen bases / colours, initbp in colour space) against
 * genome[goff .. goff+glen) of ONE packed letter array, band from the anchor (x, y, length, width;
 * widened by anchor_width of the set-up, or the threshold band when that is < 0), tie-breaks
 * reversed when revcmpl.  Letter space: maxscore = the sw_vector score of the window (used by local
 * mode only).  local_alignment = !Gflag.  Results are the fields of struct sw_full_results
 * (sw-full-common.h:13-48); the alignment comes back as an edit script (see shrimp_hit below).
 * ---------------------------------------------------------------------------------------- */
typedef struct shrimp_full_task {
  uint32_t goff;
  int32_t glen, read_idx, rlen;
  int32_t threshscore, maxscore, revcmpl;
  int32_t ax, ay, alen, awidth;
  int32_t initbp;
} shrimp_full_task;

typedef struct shrimp_full_result {
  int32_t score, read_start, rmapped, genome_start, gmapped;
  int32_t matches, mismatches, insertions, deletions, crossovers;
  int32_t edit_len;
  int64_t edit_off;
} shrimp_full_result;

int shrimp_gpu_sw_full_batch(shrimp_gpu_ctx *ctx, const uint32_t *genome, size_t genome_words,
                             const uint32_t *reads, int read_stride_words, int n_reads, int n_tasks,
                             const shrimp_full_task *tasks, int local_alignment,
                             shrimp_full_result *results, uint8_t *edits, int64_t edits_cap,
                             int64_t *edits_used);
/* ... with sw_full_cs's last argument (sw-full-cs.c:1149, read_entry::crossover_score of reads with qualities):
 * row read_idx of crossover_scores[n_reads][crossover_stride], or NULL for the global crossover score. */
int shrimp_gpu_sw_full_batch_xover(shrimp_gpu_ctx *ctx, const uint32_t *genome, size_t genome_words,
                                   const uint32_t *reads, int read_stride_words, int n_reads, int n_tasks,
                                   const shrimp_full_task *tasks, int local_alignment,
                                   const int32_t *crossover_scores, int crossover_stride,
                                   shrimp_full_result *results, uint8_t *edits, int64_t edits_cap,
                                   int64_t *edits_used);

/* ------------------------------------------------------------------------------------------
 * Genome residency.  Replaces the arrays load_genome builds (gmapper/genome.c:1092-1124,
 * globals gmapper.h:264-275): takes the reference's own genome_contigs[] (packed letters, one
 * array per contig), genome_len[] and num_contigs, and derives genome_contigs_rc (util.c:541-598)
 * and, in colour space, genome_cs_contigs / genome_cs_contigs_rc (fasta.c:586-612) in HBM.
 * Global coordinates are contig_offsets[cn] + position, as in the reference.
 * ---------------------------------------------------------------------------------------- */
int shrimp_gpu_genome_load(shrimp_gpu_ctx *ctx, int num_contigs, const uint32_t *const *genome_contigs,
                           const uint32_t *genome_len, int colour_space);
/* One GPU, several host threads: every thread owns a context (stream, chunk buffers, scoring set-up) and they
 * share the genome arrays and the projection, as gmapper's -N threads share the globals of gmapper.h:262-275.
 * dst borrows what is resident in src (same device); src must outlive dst. */
int shrimp_gpu_share_genome(shrimp_gpu_ctx *dst, shrimp_gpu_ctx *src);
/* which: 0 letters fwd, 1 letters rc, 2 colours fwd, 3 colours rc; global packed coordinates */
int shrimp_gpu_genome_export(shrimp_gpu_ctx *ctx, int which, uint32_t *out_words, size_t n_words);

/* ------------------------------------------------------------------------------------------
 * Spaced-seed projection ("genome map").  Replaces the projection loop of load_genome
 * (genome.c:1138-1166) + KMER_TO_MAPIDX (gmapper.h:323-370) for seeds parsed as add_spaced_seed
 * does (seeds.c:9-42; masks[sn] bit 0 = rightmost seed character).  hflag = -H hashed k-mers.
 * Result: per seed a CSR (genomemap_len -> offsets, genomemap lists concatenated, ascending).
 * shrimp_gpu_index_export returns it in the layout of the reference's -S files (genome.c:37-63).
 * ---------------------------------------------------------------------------------------- */
int shrimp_gpu_index_build(shrimp_gpu_ctx *ctx, int n_seeds, const uint64_t *masks, const int32_t *spans,
                           const int32_t *weights, int hflag);
int shrimp_gpu_index_nbuckets(shrimp_gpu_ctx *ctx, int sn, uint32_t *nbuckets, uint64_t *total);
int shrimp_gpu_index_export(shrimp_gpu_ctx *ctx, int sn, uint32_t *lens_out, uint32_t *pos_out, uint64_t *total_out);

/* Projection save.  Replaces save_genome_map / save_genome_map_seed (gmapper/genome.c:185-272,
 * :15-66): writes <prefix>.genome and <prefix>.seed.N in the layout of `gmapper -S`, uncompressed
 * (the reference's loader reads through gzread, which passes plain files through), so that
 * `gmapper -L <prefix>` runs on exactly the projection held in HBM.  contig_names[num_contigs]. */
int shrimp_gpu_projection_save(shrimp_gpu_ctx *ctx, const char *prefix, const char *const *contig_names);
/* Projection load.  Replaces load_genome_map / load_genome_map_seed (gmapper/genome.c:670-832, :69-182) for the
 * device: <prefix>.genome and <prefix>.seed.N as `gmapper -S` (gzip) or shrimp_gpu_projection_save (plain) wrote them
 * go straight into HBM -- the letter contigs (reverse complement and colours are derived there) and per seed the CSR
 * (genomemap_len as offsets, the position lists in file order).  Replaces genome_load + index_build for a saved
 * projection; contig names as stored in the file. */
int shrimp_gpu_projection_load(shrimp_gpu_ctx *ctx, const char *prefix);
int shrimp_gpu_num_contigs(shrimp_gpu_ctx *ctx);
const char *shrimp_gpu_contig_name(shrimp_gpu_ctx *ctx, int cn);

/* ------------------------------------------------------------------------------------------
 * Chunk-level mapping.  Replaces handle_read (gmapper/mapping.c:1773-1868) for a whole chunk of
 * unpaired reads with the default single option set of gmapper.c:2601-2632, i.e. per read:
 *   read_get_mapidxs :76, read_get_region_counts :459, read_get_anchor_list :1008,
 *   read_get_hit_list :1232, read_pass1 :1345 (f1_run / sw_vector / sw_gapless),
 *   read_get_vector_hits :1376, read_pass2 :1631 (hit_run_full_sw -> sw_full_ls / sw_full_cs,
 *   hit_run_post_sw, read_remove_duplicate_hits, ranking).
 * What read_output (gmapper/output.c:955) would receive comes back as shrimp_hit records in read
 * order, then in the order of hits_pass2[].  The caller (the reference's unchanged output.c, or a
 * test) formats SAM from them.  Field names follow options/globals of gmapper.h:50-141.
 * ---------------------------------------------------------------------------------------- */
typedef struct shrimp_map_params {
  double window_len;            /* -w: > 0 percent of read length, < 0 absolute (util.h:48-53) */
  double window_overlap;        /* -l */
  double window_gen_threshold;  /* -r */
  double sw_vect_threshold;     /* -v */
  double sw_full_threshold;     /* -h */
  double score_alpha, score_beta; /* gmapper.c:2559-2568, used by hit_run_post_sw */
  int32_t match_mode;           /* -n: 1 or 2 (unpaired); 2, 3 or 4 (pairs, default 4) */
  int32_t num_outputs;          /* -o */
  int32_t num_tmp_outputs;      /* 20 + num_outputs */
  int32_t gapless;              /* -U / mirna: sw_gapless instead of sw_vector */
  int32_t hash_filter_calls;    /* 0 with -Z */
  int32_t use_regions;
  int32_t region_bits, region_overlap;
  int32_t Gflag, Tflag;         /* global-in-read alignment; reversed tie-breaks on the rc strand */
  int32_t strata, max_alignments;
  int32_t compute_mapping_qualities;
  uint32_t list_cutoff;         /* -z / automatic (gmapper.c:2811-2837) */
  /* colour-space reads with qualities (-Q without --ignore-qvs): read_entry::crossover_score of every read
   * (gmapper.c:532-543), rows of crossover_stride ints; NULL = the global crossover score for every position */
  const int32_t *crossover_scores;
  int32_t crossover_stride;
  /* colour space with compute_mapping_qualities: post_sw (common/sw-post.c) rescoring.  read_quals = the quality
   * strings of the reads as read (re->qual, gmapper -Q; rows of qual_stride bytes) or NULL for reads without
   * qualities; qual_delta / qual_vector_offset / use_sanger_qvs as post_sw_setup gets them (gmapper.c:2960-2962);
   * pr_xover as -X (0 = the default .03).  Every reported alignment then carries sfrp->posterior, the corrected base
   * calls (edit bytes with bit 3 set: bits 4-5 are the base itself, bit 2 = lower case) and, right after its edit
   * script in `edits`, rmapped bytes of base qualities (sfrp->qual, 33 + q). */
  const uint8_t *read_quals;
  int32_t qual_stride, qual_delta, qual_vector_offset, use_sanger_qvs;
  double pr_xover;
} shrimp_map_params;

/* One reported alignment: read_hit + sw_full_results (gmapper-definitions.h:125-153,
 * sw-full-common.h:13-48).  The alignment itself is an edit script in `edits`: one byte per
 * alignment column, the reference's backtrace codes (sw-full-ls.c:44-46) -- 1 BACK_INSERTION
 * (genome base against '-'), 2 BACK_DELETION (read base against '-'), 3 BACK_MATCH_MISMATCH in
 * bits 0-1; colour space sets bit 2 when the column carries a crossover and bits 4-5 to the layer
 * (letter-space translation starting from (layer + initbp) % 4, sw-full-cs.c:1181-1196) whose
 * letter pretty_print shows.  dbalign/qralign follow from it. */
typedef struct shrimp_hit {
  int32_t read_idx, cn, gen_st, w_len;
  int64_t g_off;                /* rh->g_off, in the orientation of gen_st */
  int32_t score_vector, score_full, pass2_key, score_max, matches, sw_score;
  double  posterior;
  int32_t read_start, rmapped, genome_start, gmapped;
  int32_t sfr_matches, mismatches, insertions, deletions, crossovers;
  int32_t edit_len;
  int64_t edit_off;
  /* identity of the read_hit inside the chunk (its slot in the chunk's hit lists): two records with the same
   * hit_slot are the SAME struct read_hit in the reference -- readpair_save_final_hits (mapping.c:2446-2500) pools
   * the members of the pairs by that identity before compute_paired_mqv (output.c:811-941) sums over them.
   * st = the read strand the hit was found on before hit_run_full_sw re-oriented it (mapping.c:349-351). */
  int32_t hit_slot, st;
  int32_t score_window_gen;     /* rh->score_window_gen (mapping.c:1185), the ZR field of --extra-sam-fields */
  int32_t reserved;
} shrimp_hit;

/* A read_hit after read_pass1 (stage-level parity with DEBUG_HIT_LIST_PASS1 dumps) */
typedef struct shrimp_stage_hit {
  int32_t read_idx, st, cn, w_len;
  int64_t g_off;                /* g_off_pos_strand */
  int32_t score_window_gen, matches, score_max, score_vector, pct_score_vector;
  int32_t ax, ay, alen, awidth;
} shrimp_stage_hit;

typedef struct shrimp_map_stats {
  uint64_t list_entries;        /* index positions gathered */
  uint64_t surviving_entries;   /* after the region filter */
  uint64_t anchors, hits;
  uint64_t heap_replays;        /* read strands that needed exact heap-order emulation */
  uint64_t vector_tasks;        /* windows scored on the device (superset) */
  uint64_t vector_calls;        /* sw_vector calls the reference makes (pass 1 + pass 2) */
  uint64_t vector_cells;        /* sum glen*rlen over those calls (sw-vector.c:509) */
  uint64_t vector_bypassed;     /* f1 cache hits */
  uint64_t full_calls, full_cells;
  uint64_t device_vector_cells; /* sum glen*rlen over the windows the device scored */
  uint64_t scan_big_strands;    /* read strands served by the CTA-per-strand scan kernel (long index lists) */
  uint64_t scan_global_strands; /* ... of which needed candidate arrays in global memory (more than a shared-memory slab) */
  uint64_t post_sw_columns;     /* aligned read columns post_sw ran its 16-node forward-backward over */
} shrimp_map_stats;

/* initbp: per-read initial base (colour space) or NULL.  hits_cap >= n_reads*num_outputs always
 * suffices.  Returns SHRIMP_E_NOMEM with *edits_used = required bytes if edits_cap is too small.
 * stage / stage_cap / n_stage are optional (NULL / 0). */
int shrimp_gpu_map_reads(shrimp_gpu_ctx *ctx, const shrimp_map_params *mp, int n_reads,
                         const uint32_t *reads, int read_stride_words, const int32_t *read_len,
                         const int8_t *initbp,
                         shrimp_hit *hits, int64_t hits_cap, int32_t *n_hits_per_read,
                         uint8_t *edits, int64_t edits_cap, int64_t *n_hits, int64_t *edits_used,
                         shrimp_stage_hit *stage, int64_t stage_cap, int64_t *n_stage,
                         shrimp_map_stats *stats);

/* ------------------------------------------------------------------------------------------
 * Chunk-level mapping of read pairs.  Replaces handle_readpair (gmapper/mapping.c:2504-2650) for a chunk
 * of pairs with the paired option sets of gmapper.c:2638-2714: mp->match_mode 4 (default), 3 or 2, with or without
 * half-pairing.  readpair_compute_mp_ranges :2317; per-read region counts; for -n 3 and for --no-half-paired the
 * mate-pair region counts (read_get_mp_region_counts :546, the paired rules of advance_index_in_genomemap :667-728,
 * hit-list mode 3 :1082-1094); anchors / hit lists; readpair_pair_up_hits :266, read_pass1 with only_paired :1261, readpair_get_vector_hits
 * :1877, readpair_pass2 :2181 (hit_run_full_sw at half the full threshold, readpair_remove_duplicate_hits,
 * ranking), then (half_paired) the half-paired fall-back into handle_read :1773 for both mates (pass 1 + pass 2).
 * reads rows 2k and 2k+1 are the mates of pair k (gmapper -1 / -2).  What readpair_output
 * (gmapper/output.c:1071) would receive comes back as:
 *   pairs[]  final_paired_hits in pair order, each naming its two shrimp_hit records (mate 0, mate 1);
 *   hits[]   first the members of the pairs, then every read's final_unpaired_hits in read order
 *            (n_unpaired_per_read[2k + mate]).
 * ---------------------------------------------------------------------------------------- */
typedef struct shrimp_pair_params {
  int32_t pair_mode;            /* 1 opp-in, 2 opp-out, 3 col-fw, 4 col-bw (gmapper-definitions.h:42-47) */
  int32_t min_insert_size, max_insert_size;   /* -I */
  int32_t half_paired;          /* 1 (default) / 0 = --no-half-paired */
} shrimp_pair_params;

typedef struct shrimp_pair {    /* struct read_hit_pair, gmapper-definitions.h:155-165 */
  int32_t pair_idx;
  int32_t score, score_max, key, insert_size;
  int32_t hit_idx[2];
} shrimp_pair;

int shrimp_gpu_map_pairs(shrimp_gpu_ctx *ctx, const shrimp_map_params *mp, const shrimp_pair_params *pp,
                         int n_pairs, const uint32_t *reads, int read_stride_words, const int32_t *read_len,
                         const int8_t *initbp,
                         shrimp_hit *hits, int64_t hits_cap, int64_t *n_hits,
                         shrimp_pair *pairs, int64_t pairs_cap, int64_t *n_pairs_out,
                         int32_t *n_pairs_per_pair, int32_t *n_unpaired_per_read,
                         uint8_t *edits, int64_t edits_cap, int64_t *edits_used, shrimp_map_stats *stats);

/* Measurement entries (no reference counterpart).  shrimp_gpu_map_resident re-runs every device
 * stage on the reads the previous shrimp_gpu_map_reads call left in HBM and keeps the results on
 * the device: the "inputs already resident" throughput of bench.py.  shrimp_gpu_last_transfer_bytes
 * reports the host<->device bytes of the last shrimp_gpu_map_reads call. */
int shrimp_gpu_map_resident(shrimp_gpu_ctx *ctx, const shrimp_map_params *mp, shrimp_map_stats *stats);
/* same for pairs: maps again the pairs the previous shrimp_gpu_map_pairs call left in HBM; the mid-pipeline host
 * stage (readpair_pass2) is part of the path and stays inside, the records go to library-owned scratch */
int shrimp_gpu_map_pairs_resident(shrimp_gpu_ctx *ctx, const shrimp_map_params *mp, const shrimp_pair_params *pp,
                                  shrimp_map_stats *stats);
int shrimp_gpu_last_transfer_bytes(shrimp_gpu_ctx *ctx, uint64_t *h2d, uint64_t *d2h);

/* CUDA-event bracket on the library's stream (which = 0 start, 1 stop) and an L2 flush (writes a
 * 256 MB buffer), for timing hygiene in bench.py. */
int shrimp_gpu_event_record(shrimp_gpu_ctx *ctx, int which);
int shrimp_gpu_event_elapsed_ms(shrimp_gpu_ctx *ctx, float *ms);
int shrimp_gpu_flush_l2(shrimp_gpu_ctx *ctx);

/* ------------------------------------------------------------------------------------------
 * Measurement helper (no reference counterpart): integer-pipe peak, in giga thread-level
 * VIADDMNMX.S16x2 instructions per second, measured with a register-resident micro-benchmark.
 * bench.py uses it as the denominator of the sw_vector roofline.
 * ---------------------------------------------------------------------------------------- */
int shrimp_gpu_dpx_peak(shrimp_gpu_ctx *ctx, double *ginstr_per_s);
/* ... and the FP64 issue peak (giga thread-level DFMA per second, register-resident chains), the denominator of the
 * post_sw roofline. */
int shrimp_gpu_fp64_peak(shrimp_gpu_ctx *ctx, double *ginstr_per_s);

/* Host threads of the stages that stay on the CPU as in the reference (read_pass2's duplicate removal and ranking,
 * the -N threads of gmapper.c:2907): n > 0 fixes the count, 0 = the OpenMP default. Process-wide. */
int shrimp_gpu_set_host_threads(int n);

/* Diagnostic: exp() and log() of n doubles as post_sw's kernel computes them (a transcription of the libm the
 * reference runs on), so that a test can compare them bit for bit with the host's libm. */
int shrimp_gpu_glibc_explog(shrimp_gpu_ctx *ctx, const double *x, int n, double *exp_out, double *log_out);

/* Diagnostic: hash_genome_window (common/util.h:220-241, the key of the f1 window cache, f1-wrapper.h:97-134) of n
 * windows [goff, goff + glen) of a packed genome (8 four-bit codes per word, host memory) as the device computes it. */
int shrimp_gpu_hash_windows(shrimp_gpu_ctx *ctx, const uint32_t *genome_words, uint64_t n_words, const uint32_t *goff,
                            const int32_t *glen, int n, uint32_t *hash_out);

#ifdef __cplusplus
}
#endif
#endif /* SHRIMP_B200_H */
