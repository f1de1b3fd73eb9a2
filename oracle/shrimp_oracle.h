/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the gmapper hot path of
 * compbio-UofT/shrimp 2.2.3.  Plain scalar C written from the reference's algorithms, each
 * function citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load this; the product (libshrimp_b200.so)
 * never does.
 *
 * Pinning: tests/test_oracle_*.py check every function here against the reference's own objects
 * (oracle/_ref/libshrimp_ref.so, built from /root/reference by oracle/Makefile) on random cases
 * and against golden vectors generated from the reference binary (tests/golden/, made by
 * tests/golden/make_golden.py).
 */
#ifndef SHRIMP_ORACLE_H
#define SHRIMP_ORACLE_H
#include <stdint.h>

typedef struct orc_scores {
  int match, mismatch;           /* CLI sign: mismatch < 0 */
  int a_gap_open, a_gap_ext;     /* <= 0 */
  int b_gap_open, b_gap_ext;     /* <= 0 */
  int crossover;                 /* < 0 (colour space) */
} orc_scores;

int orc_sw_vector(const uint32_t *genome, int goff, int glen, const uint32_t *read, int rlen,
                  const uint32_t *genome_ls, int initbp, const orc_scores *sc);
int orc_sw_gapless(const uint32_t *genome, int glen, const uint32_t *read, int rlen, int g_idx, int r_idx,
                   const uint32_t *genome_ls, int initbp, const orc_scores *sc);
uint32_t orc_hash_genome_window(const uint32_t *genome, uint32_t goff, uint32_t glen);


/* ---------------------------------------------------------------------------------------------
 * Genome arrays (gmapper/genome.c:1092-1124, util.c:541-598, fasta.c:586-605)
 * ------------------------------------------------------------------------------------------- */
typedef struct orc_genome {
  int num_contigs;
  int colour_space;
  uint32_t *contig_offsets; /* global coordinate of each contig start */
  uint32_t *genome_len;
  uint32_t **ls, **ls_rc;   /* 4-bit packed letters, forward and reverse complement */
  uint32_t **cs, **cs_rc;   /* colour arrays (colour space only) */
  uint64_t total_len;
} orc_genome;

orc_genome *orc_genome_create(int num_contigs, const uint8_t *codes, const uint32_t *lens, int colour_space);
void orc_genome_destroy(orc_genome *g);

/* ---------------------------------------------------------------------------------------------
 * Seeds and the projection (gmapper/seeds.c:9-42, gmapper.h:349-368, genome.c:1138-1166)
 * ------------------------------------------------------------------------------------------- */
#define ORC_MAX_SEEDS 16
typedef struct orc_index {
  int n_seeds;
  uint64_t mask[ORC_MAX_SEEDS];
  int span[ORC_MAX_SEEDS], weight[ORC_MAX_SEEDS];
  int max_span, min_span;
  int hflag;                       /* -H: hashed k-mers into 4^12 buckets */
  uint32_t nbuckets[ORC_MAX_SEEDS];
  uint32_t *len[ORC_MAX_SEEDS];    /* genomemap_len[sn][mapidx] */
  uint32_t *start[ORC_MAX_SEEDS];  /* offset of the bucket in pos[sn] */
  uint32_t *pos[ORC_MAX_SEEDS];    /* genomemap[sn][mapidx][..], ascending global start positions */
  uint64_t total[ORC_MAX_SEEDS];
} orc_index;

orc_index *orc_index_build(const orc_genome *g, int n_seeds, const uint64_t *masks, const int *spans,
                           const int *weights, int hflag);
void orc_index_destroy(orc_index *ix);
uint32_t orc_kmer_to_mapidx(const orc_index *ix, int sn, const uint32_t *seq, int start);

/* ---------------------------------------------------------------------------------------------
 * Full smith-waterman (common/sw-full-ls.c, sw-full-cs.c) and anchors (common/anchors.c)
 * ------------------------------------------------------------------------------------------- */
#define ORC_ALN_CAP 640
typedef struct orc_sfr {
  int read_start, rmapped, genome_start, gmapped;
  int matches, mismatches, insertions, deletions, score, crossovers;
  char dbalign[ORC_ALN_CAP];
  char qralign[ORC_ALN_CAP];
  char qual[ORC_ALN_CAP];      /* sfrp->qual: base qualities from post_sw (colour space with mapping qualities) */
} orc_sfr;

void orc_sw_full_ls(const uint32_t *genome, int goff, int glen, const uint32_t *read, int rlen,
                    int threshscore, int maxscore, int revcmpl,
                    long long ax, long long ay, int alen, int awidth, int anchor_width,
                    int local_alignment, const orc_scores *sc, orc_sfr *out);

void orc_sw_full_cs(const uint32_t *genome_ls, int goff, int glen, const uint32_t *read, int rlen, int initbp,
                    int threshscore, int revcmpl, long long ax, long long ay, int alen, int awidth,
                    int anchor_width, int indel_taboo_len, int local_alignment, const int *crossover_scores,
                    const orc_scores *sc, orc_sfr *out, uint64_t *cells);

/* ---------------------------------------------------------------------------------------------
 * The per-read pipeline (gmapper/mapping.c handle_read :1773 with the default unpaired option
 * set of gmapper.c:2601-2632)
 * ------------------------------------------------------------------------------------------- */
typedef struct orc_params {
  orc_scores sc;               /* sc.mismatch = LS mismatch score; CS vector filter uses match+crossover */
  int colour_space;
  double window_len;           /* >0: percent of read length, <0: -absolute (util.h:48-53) */
  double window_overlap;
  double window_gen_threshold;
  double sw_vect_threshold;
  double sw_full_threshold;
  int match_mode;              /* 1 or 2 unpaired */
  int num_outputs;             /* 10 */
  int num_tmp_outputs;         /* 30 */
  int anchor_width;            /* 8 */
  int indel_taboo_len;
  int gapless;                 /* gapless_sw (-U / mirna) */
  int hash_filter_calls;       /* f1 window cache on (default) */
  int use_regions;
  int region_bits, region_overlap;
  int Gflag, Tflag;
  int strata, max_alignments;
  int compute_mapping_qualities;
  uint32_t list_cutoff;
  double score_alpha, score_beta;
} orc_params;

/* pairing options (gmapper.h:140-142, gmapper.c:2638-2714); match_mode 4 / default paired set only */
typedef struct orc_pair_params {
  int pair_mode;               /* 1 opp-in, 2 opp-out, 3 col-fw, 4 col-bw (gmapper-definitions.h:42-47) */
  int min_insert_size, max_insert_size;
  int half_paired;
} orc_pair_params;

typedef struct orc_stage_hit {   /* a read_hit after read_pass1 (mapping.c:1345) */
  int read_idx, st, cn, w_len;
  long long g_off;               /* g_off_pos_strand */
  int score_window_gen, matches, score_max;
  int score_vector, pct_score_vector;
  int ax, ay, alen, awidth;      /* anchor relative to the window, positive-strand orientation */
} orc_stage_hit;

typedef struct orc_hit_out {     /* one reported alignment (a hits_pass2 entry handed to read_output) */
  int read_idx, cn, gen_st, st;
  long long g_off;
  int w_len, score_vector, score_full, pass2_key, score_max, matches;
  int sw_score;                  /* sfrp->score */
  double posterior;
  orc_sfr sfr;
} orc_hit_out;

typedef struct orc_stats {
  uint64_t vector_calls, vector_cells, vector_bypassed, full_calls, full_cells;
  uint64_t n_anchors, n_hits, eq_x_ties;
} orc_stats;

/* Maps n_reads unpaired reads.  out[] receives the reported hits in output order (read order,
 * then the order read_output would print them); stage[] (optional) the post-pass1 hit lists.
 * Returns the number of out entries, or -1 if a capacity is too small. */
long long orc_map_reads(const orc_genome *g, const orc_index *ix, const orc_params *p, int n_reads,
                        const uint32_t *reads, int stride_words, const int *read_len, const int8_t *initbp,
                        orc_hit_out *out, long long out_cap, int *n_out_per_read,
                        orc_stage_hit *stage, long long stage_cap, long long *n_stage, orc_stats *stats);

/* handle_readpair (mapping.c:2504-2650) for n_pairs pairs; reads 2k, 2k+1 are mates. */
long long orc_map_pairs(const orc_genome *g, const orc_index *ix, const orc_params *p, const orc_pair_params *pp,
                        int n_pairs, const uint32_t *reads, int stride_words, const int *read_len,
                        const int8_t *initbp, orc_hit_out *pairs_out, int *pair_info, long long pairs_cap,
                        int *n_pairs_per_pair, orc_hit_out *unp_out, long long unp_cap, int *n_unp_per_read,
                        long long *n_unp, orc_stats *stats);

#endif
