/* TEST INFRASTRUCTURE ONLY -- see shrimp_oracle.h. */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "shrimp_oracle.h"

#define EX4(a, i) ((int)(((a)[(i) >> 3] >> (4 * ((i) & 7))) & 0xf)) /* EXTRACT, common/util.h:41 */
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* lstocs, common/util.h:182-207 (DNA only: no uracil handling) */
static inline int orc_lstocs(int first, int second) {
  if (first > 3 || second > 3) return 15;
  return first ^ second;
}

/*
 * sw_vector: common/sw-vector.c:453-515 with vect_sw_diff_gap :68-217 (the same_gap copy :228-377
 * is the a==b special case).  The SSE code is Wozniak's anti-diagonal layout of a plain Gotoh
 * recurrence; restated here row by row.  nogap[] = H of the previous row, b_gap[] = F.
 * Colour space: row 0 compares lstocs(letter genome, initbp) with read colour 0 (:116-146).
 */
int orc_sw_vector(const uint32_t *genome, int goff, int glen, const uint32_t *read, int rlen,
                  const uint32_t *genome_ls, int initbp, const orc_scores *sc) {
  const int ao = -sc->a_gap_open, ae = -sc->a_gap_ext, bo = -sc->b_gap_open, be = -sc->b_gap_ext;
  int *H = (int *)malloc(sizeof(int) * (glen + 1));
  int *F = (int *)malloc(sizeof(int) * (glen + 1));
  int score = 0;
  for (int j = 0; j <= glen; j++) {
    H[j] = 0;
    F[j] = -bo;
  }
  for (int i = 0; i < rlen; i++) {
    int q = EX4(read, i);
    int E = -ao, hleft = 0, hdiag = 0;
    for (int j = 0; j < glen; j++) {
      int d = (genome_ls != NULL && i == 0) ? orc_lstocs(EX4(genome_ls, goff + j), initbp) : EX4(genome, goff + j);
      int ms = (d == q) ? sc->match : sc->mismatch;
      E = imax(hleft - ao - ae, E - ae);
      F[j] = imax(H[j] - bo - be, F[j] - be);
      int h = imax(hdiag + ms, 0);
      h = imax(h, E);
      h = imax(h, F[j]);
      hdiag = H[j];
      H[j] = h;
      hleft = h;
      score = imax(score, h);
    }
  }
  free(H);
  free(F);
  return score;
}

/* sw_gapless: common/sw-gapless.c:57-117 */
int orc_sw_gapless(const uint32_t *genome, int glen, const uint32_t *read, int rlen, int g_idx, int r_idx,
                   const uint32_t *genome_ls, int initbp, const orc_scores *sc) {
  int g_left, r_left, g_right, r_right, score = 0, max_score;
  if (g_idx < r_idx) {
    g_left = 0;
    r_left = r_idx - g_idx;
  } else {
    g_left = g_idx - r_idx;
    r_left = 0;
  }
  g_right = g_left;
  r_right = r_left;
  if (genome_ls != NULL && r_left == 0) { /* first colour is forced through the letter genome (:83-93) */
    int real_colour = orc_lstocs(EX4(genome_ls, g_right), initbp);
    if (real_colour == EX4(read, 0)) score = sc->match;
    r_right++;
    g_right++;
  }
  max_score = score;
  while (g_right < glen && r_right < rlen) {
    score += (EX4(genome, g_right) == EX4(read, r_right)) ? sc->match : sc->mismatch;
    if (score > max_score) max_score = score;
    g_right++;
    r_right++;
    if (score < 0) score = 0;
  }
  return max_score;
}

/* hash_accumulate / hash_finalize (common/hash.h:69-92) and hash_genome_window (util.h:220-241) */
static inline void orc_hash_acc(uint32_t *key, uint32_t val) {
  uint32_t tmp;
  *key += (val >> 16);
  tmp = ((val & 0xFFFF) << 11) ^ *key;
  *key = (*key << 16) ^ tmp;
  *key += *key >> 11;
}
static inline void orc_hash_fin(uint32_t *key) {
  *key ^= *key << 3;
  *key += *key >> 5;
  *key ^= *key << 4;
  *key += *key >> 17;
  *key ^= *key << 25;
  *key += *key >> 6;
}
uint32_t orc_hash_genome_window(const uint32_t *genome, uint32_t goff, uint32_t glen) {
  uint32_t key = 0;
  for (uint32_t i = 0; i < (glen + 15) / 16; i++) {
    uint32_t buffer = 0;
    for (uint32_t j = 0; j < 16 && i * 16 + j < glen; j++) {
      buffer <<= 2;
      buffer |= (uint32_t)EX4(genome, goff + i * 16 + j) & 3u;
    }
    orc_hash_acc(&key, buffer);
  }
  orc_hash_fin(&key);
  return key;
}
