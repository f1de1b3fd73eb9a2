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

#endif
