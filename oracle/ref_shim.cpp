// TEST INFRASTRUCTURE ONLY (oracle/): extern "C" wrappers around the *unmodified* reference
// objects (common/sw-vector.c, sw-gapless.c, sw-full-ls.c, sw-full-cs.c, anchors.c, util.c)
// compiled in place from /root/reference by oracle/Makefile into oracle/_ref/libshrimp_ref.so.
// Used by tests/ to pin the C restatement in oracle/shrimp_oracle.c and to generate the golden
// vectors under tests/golden/.  Nothing in the product path may link or load this file.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "common/sw-vector.h"
#include "common/sw-gapless.h"
#include "common/sw-full-common.h"
#include "common/sw-full-ls.h"
#include "common/sw-full-cs.h"
#include "common/anchors.h"
#include "common/util.h"

// POD mirror of the fields of sw_full_results (sw-full-common.h:13-48) the path produces.
struct ref_sfr {
  int read_start, rmapped, genome_start, gmapped;
  int matches, mismatches, insertions, deletions, score, crossovers;
  char dbalign[4096];
  char qralign[4096];
};

extern "C" {

int ref_sw_vector_setup(int dblen, int qrlen, int a_open, int a_ext, int b_open, int b_ext,
                        int match, int mismatch, int use_colours) {
  return sw_vector_setup(dblen, qrlen, a_open, a_ext, b_open, b_ext, match, mismatch, use_colours, true);
}

int ref_sw_vector(uint32_t *genome, int goff, int glen, uint32_t *read, int rlen,
                  uint32_t *genome_ls, int initbp) {
  return sw_vector(genome, goff, glen, read, rlen, genome_ls, initbp, false);
}

int ref_sw_gapless_setup(int match, int mismatch) { return sw_gapless_setup(match, mismatch, true); }

int ref_sw_gapless(uint32_t *genome, int glen, uint32_t *read, int rlen, int g_idx, int r_idx,
                   uint32_t *genome_ls, int initbp) {
  return sw_gapless(genome, glen, read, rlen, g_idx, r_idx, genome_ls, initbp, false);
}

int ref_sw_full_ls_setup(int dblen, int qrlen, int a_open, int a_ext, int b_open, int b_ext,
                         int match, int mismatch, int anchor_width) {
  return sw_full_ls_setup(dblen, qrlen, a_open, a_ext, b_open, b_ext, match, mismatch, true, anchor_width);
}

static void copy_out(struct sw_full_results *sfr, struct ref_sfr *out) {
  out->read_start = sfr->read_start; out->rmapped = sfr->rmapped;
  out->genome_start = sfr->genome_start; out->gmapped = sfr->gmapped;
  out->matches = sfr->matches; out->mismatches = sfr->mismatches;
  out->insertions = sfr->insertions; out->deletions = sfr->deletions;
  out->score = sfr->score; out->crossovers = sfr->crossovers;
  out->dbalign[0] = out->qralign[0] = 0;
  if (sfr->dbalign) { strncpy(out->dbalign, sfr->dbalign, sizeof(out->dbalign) - 1); free(sfr->dbalign); }
  if (sfr->qralign) { strncpy(out->qralign, sfr->qralign, sizeof(out->qralign) - 1); free(sfr->qralign); }
}

// anchor passed as (x, y, length, width)
void ref_sw_full_ls(uint32_t *genome, int goff, int glen, uint32_t *read, int rlen,
                    int threshscore, int maxscore, int revcmpl,
                    long long ax, long long ay, int alen, int awidth,
                    int local_alignment, struct ref_sfr *out) {
  struct sw_full_results sfr;
  struct anchor a;
  memset(&sfr, 0, sizeof(sfr));
  memset(&a, 0, sizeof(a));
  a.x = ax; a.y = ay; a.length = alen; a.width = awidth; a.weight = 1;
  sw_full_ls(genome, goff, glen, read, rlen, threshscore, maxscore, &sfr, revcmpl != 0, &a, 1, local_alignment);
  copy_out(&sfr, out);
}

int ref_sw_full_cs_setup(int dblen, int qrlen, int a_open, int a_ext, int b_open, int b_ext,
                         int match, int mismatch, int xover, int anchor_width, int indel_taboo_len) {
  return sw_full_cs_setup(dblen, qrlen, a_open, a_ext, b_open, b_ext, match, mismatch, xover, true,
                          anchor_width, indel_taboo_len);
}

void ref_sw_full_cs(uint32_t *genome, int goff, int glen, uint32_t *read, int rlen, int initbp,
                    int threshscore, int revcmpl,
                    long long ax, long long ay, int alen, int awidth,
                    int local_alignment, int *crossover_scores, struct ref_sfr *out) {
  struct sw_full_results sfr;
  struct anchor a;
  memset(&sfr, 0, sizeof(sfr));
  memset(&a, 0, sizeof(a));
  a.x = ax; a.y = ay; a.length = alen; a.width = awidth; a.weight = 1;
  sw_full_cs(genome, goff, glen, read, rlen, initbp, threshscore, &sfr, revcmpl != 0, false, &a, 1,
             local_alignment, crossover_scores);
  copy_out(&sfr, out);
}

void ref_anchor_get_x_range(long long ax, long long ay, int alen, int awidth, int x_len, int y_len, int y,
                            int *x_min, int *x_max) {
  struct anchor a; memset(&a, 0, sizeof(a));
  a.x = ax; a.y = ay; a.length = alen; a.width = awidth;
  anchor_get_x_range(&a, x_len, y_len, y, x_min, x_max);
}

uint32_t ref_hash_genome_window(uint32_t *genome, unsigned goff, unsigned glen) {
  return hash_genome_window(genome, goff, glen);
}

void ref_reverse_complement_read_ls(uint32_t *read, uint32_t len, uint32_t *out) {
  uint32_t *r = reverse_complement_read_ls(read, len, false);
  memcpy(out, r, sizeof(uint32_t) * BPTO32BW(len));
  free(r);
}

void ref_reverse_complement_read_cs(uint32_t *read, int initbp, uint32_t len, uint32_t *out) {
  uint32_t *r = reverse_complement_read_cs(read, (int8_t)initbp, (int8_t)initbp, len, false);
  memcpy(out, r, sizeof(uint32_t) * BPTO32BW(len));
  free(r);
}

}  // extern "C"
