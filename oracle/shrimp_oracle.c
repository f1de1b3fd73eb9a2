/* TEST INFRASTRUCTURE ONLY -- see shrimp_oracle.h. */
#include <ctype.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
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

/* =============================================================================================
 * Genome arrays
 * ===========================================================================================*/
#define WORDS4(n) (((n) + 7) / 8) /* BPTO32BW, util.h:42 */
static inline void put4(uint32_t *a, uint64_t i, uint32_t v) { /* bitfield_append, util.c:362-371 */
  uint32_t w = a[i >> 3];
  w &= ~(0xfu << (4 * (i & 7)));
  w |= (v & 0xfu) << (4 * (i & 7));
  a[i >> 3] = w;
}
/* complement table, util.h:128-145 (DNA: U complements to A, never produced) */
static const uint8_t orc_cmpl[16] = {3, 2, 1, 0, 0, 10, 9, 7, 8, 6, 5, 14, 13, 12, 11, 15};

orc_genome *orc_genome_create(int num_contigs, const uint8_t *codes, const uint32_t *lens, int colour_space) {
  orc_genome *g = (orc_genome *)calloc(1, sizeof(*g));
  g->num_contigs = num_contigs;
  g->colour_space = colour_space;
  g->contig_offsets = (uint32_t *)calloc(num_contigs, sizeof(uint32_t));
  g->genome_len = (uint32_t *)calloc(num_contigs, sizeof(uint32_t));
  g->ls = (uint32_t **)calloc(num_contigs, sizeof(uint32_t *));
  g->ls_rc = (uint32_t **)calloc(num_contigs, sizeof(uint32_t *));
  g->cs = (uint32_t **)calloc(num_contigs, sizeof(uint32_t *));
  g->cs_rc = (uint32_t **)calloc(num_contigs, sizeof(uint32_t *));
  uint64_t off = 0;
  for (int c = 0; c < num_contigs; c++) {
    uint32_t n = lens[c];
    g->contig_offsets[c] = (uint32_t)off; /* genome.c:1072 */
    g->genome_len[c] = n;
    size_t w = WORDS4(n) + 2;
    g->ls[c] = (uint32_t *)calloc(w, 4);
    g->ls_rc[c] = (uint32_t *)calloc(w, 4);
    for (uint32_t i = 0; i < n; i++) {
      put4(g->ls[c], i, codes[off + i]);
      put4(g->ls_rc[c], n - 1 - i, orc_cmpl[codes[off + i] & 15]); /* reverse_complement_read_ls, util.c:541 */
    }
    if (colour_space) { /* bitfield_to_colourspace, fasta.c:592-612: first colour against T */
      g->cs[c] = (uint32_t *)calloc(w, 4);
      g->cs_rc[c] = (uint32_t *)calloc(w, 4);
      int last = 3, last_rc = 3;
      for (uint32_t i = 0; i < n; i++) {
        int a = EX4(g->ls[c], i), b = EX4(g->ls_rc[c], i);
        put4(g->cs[c], i, orc_lstocs(last, a));
        put4(g->cs_rc[c], i, orc_lstocs(last_rc, b));
        last = a;
        last_rc = b;
      }
    }
    off += n;
  }
  g->total_len = off;
  return g;
}

void orc_genome_destroy(orc_genome *g) {
  if (!g) return;
  for (int c = 0; c < g->num_contigs; c++) {
    free(g->ls[c]);
    free(g->ls_rc[c]);
    free(g->cs[c]);
    free(g->cs_rc[c]);
  }
  free(g->ls);
  free(g->ls_rc);
  free(g->cs);
  free(g->cs_rc);
  free(g->contig_offsets);
  free(g->genome_len);
  free(g);
}

/* =============================================================================================
 * Seeds / projection
 * ===========================================================================================*/
/* integer hash of gmapper.h:309-319 */
static inline uint32_t orc_hash32(uint32_t a) {
  a = (a + 0x7ed55d16) + (a << 12);
  a = (a ^ 0xc761c23c) ^ (a >> 19);
  a = (a + 0x165667b1) + (a << 5);
  a = (a + 0xd3a2646c) ^ (a << 9);
  a = (a + 0xfd7046c5) + (a << 3);
  a = (a ^ 0xb55a4f09) ^ (a >> 16);
  return a;
}

/*
 * KMER_TO_MAPIDX (gmapper.h:370) of the k-mer of seed sn that STARTS at base `start` of seq.
 * The reference keeps a sliding window whose nibble 0 is the newest base (bitfield_prepend,
 * util.c:335-348); mask bit i therefore selects base start+span-1-i.
 *  - kmer_to_mapidx_orig (gmapper.h:349-368): low 2 bits of the selected bases, bit 0 first and
 *    ending up most significant;
 *  - kmer_to_mapidx_hash (gmapper.h:323-336): iterated hash over the masked 4-bit window words.
 */
uint32_t orc_kmer_to_mapidx(const orc_index *ix, int sn, const uint32_t *seq, int start) {
  const int span = ix->span[sn];
  const uint64_t mask = ix->mask[sn];
  if (!ix->hflag) {
    uint32_t m = 0;
    uint64_t a = mask;
    int i = 0;
    do {
      if (a & 1) m = (m << 2) | ((uint32_t)EX4(seq, start + span - 1 - i) & 3u);
      a >>= 1;
      i++;
    } while (a != 0);
    return m;
  } else {
    uint32_t m = 0;
    int nw = WORDS4(ix->max_span);
    for (int w = 0; w < nw; w++) {
      uint32_t word = 0;
      for (int n = 0; n < 8; n++) {
        int i = 8 * w + n;
        if (i < span && ((mask >> i) & 1)) word |= (uint32_t)EX4(seq, start + span - 1 - i) << (4 * n);
      }
      m = orc_hash32(word ^ m);
    }
    return m & ((1u << 24) - 1); /* 4^HASH_TABLE_POWER - 1 */
  }
}

orc_index *orc_index_build(const orc_genome *g, int n_seeds, const uint64_t *masks, const int *spans,
                           const int *weights, int hflag) {
  orc_index *ix = (orc_index *)calloc(1, sizeof(*ix));
  ix->n_seeds = n_seeds;
  ix->hflag = hflag;
  ix->min_span = 64;
  for (int sn = 0; sn < n_seeds; sn++) {
    ix->mask[sn] = masks[sn];
    ix->span[sn] = spans[sn];
    ix->weight[sn] = weights[sn];
    if (spans[sn] > ix->max_span) ix->max_span = spans[sn];
    if (spans[sn] < ix->min_span) ix->min_span = spans[sn];
    ix->nbuckets[sn] = 1u << (2 * (hflag ? 12 : weights[sn]));
  }
  /* genome.c:1138-1166: for every contig position i (last base of the k-mer), any N/X resets the
   * load counter; a seed fires once `load >= span`; stored value is the global START position.
   * Colour space projects the colour genome (:1125-1135).  Two counting passes give the same
   * per-bucket ascending lists the reference builds with one realloc per position. */
  for (int sn = 0; sn < n_seeds; sn++) {
    ix->len[sn] = (uint32_t *)calloc(ix->nbuckets[sn], 4);
    ix->start[sn] = (uint32_t *)calloc((size_t)ix->nbuckets[sn] + 1, 4);
  }
  for (int pass = 0; pass < 2; pass++) {
    for (int c = 0; c < g->num_contigs; c++) {
      const uint32_t *seq = g->colour_space ? g->cs[c] : g->ls[c];
      int load = 0;
      for (uint32_t i = 0; i < g->genome_len[c]; i++) {
        int base = EX4(seq, i);
        if (base == 15)
          load = 0;
        else if (load < ix->max_span)
          load++;
        for (int sn = 0; sn < n_seeds; sn++) {
          if (load < ix->span[sn]) continue;
          uint32_t st = i - ix->span[sn] + 1;
          uint32_t m = orc_kmer_to_mapidx(ix, sn, seq, (int)st);
          if (pass == 0)
            ix->len[sn][m]++;
          else
            ix->pos[sn][ix->start[sn][m] + ix->len[sn][m]++] = g->contig_offsets[c] + st;
        }
      }
    }
    if (pass == 0) {
      for (int sn = 0; sn < n_seeds; sn++) {
        uint64_t tot = 0;
        for (uint32_t m = 0; m < ix->nbuckets[sn]; m++) {
          ix->start[sn][m] = (uint32_t)tot;
          tot += ix->len[sn][m];
          ix->len[sn][m] = 0;
        }
        ix->start[sn][ix->nbuckets[sn]] = (uint32_t)tot;
        ix->total[sn] = tot;
        ix->pos[sn] = (uint32_t *)calloc(tot + 1, 4);
      }
    }
  }
  return ix;
}

void orc_index_destroy(orc_index *ix) {
  if (!ix) return;
  for (int sn = 0; sn < ix->n_seeds; sn++) {
    free(ix->len[sn]);
    free(ix->start[sn]);
    free(ix->pos[sn]);
  }
  free(ix);
}

/* =============================================================================================
 * Anchors (common/anchors.c, anchors.h)
 * ===========================================================================================*/
typedef struct o_anchor {
  long long x, y;
  int length, width, weight, cn;
} o_anchor;

static void o_anchor_join(const o_anchor *a, int n, o_anchor *dest) { /* anchors.c:9-54 */
  long long nw_min = INT_MAX, sw_min = INT_MAX, ne_max = INT_MIN, se_max = INT_MIN;
  dest->weight = 0;
  dest->cn = a[0].cn;
  for (int i = 0; i < n; i++) {
    long long nw = a[i].x + a[i].y, sw = a[i].x - a[i].y;
    long long ne = sw + 2 * (a[i].width - 1), se = nw + 2 * (a[i].length - 1);
    if (nw < nw_min) nw_min = nw;
    if (sw < sw_min) sw_min = sw;
    if (ne > ne_max) ne_max = ne;
    if (se > se_max) se_max = se;
    dest->weight += a[i].weight;
  }
  if ((nw_min + sw_min) % 2 != 0) nw_min--;
  dest->x = (nw_min + sw_min) / 2;
  dest->y = nw_min - dest->x;
  if ((ne_max - sw_min) % 2 != 0) ne_max++;
  dest->width = (int)((ne_max - sw_min) / 2 + 1);
  if ((se_max - nw_min) % 2 != 0) se_max++;
  dest->length = (int)((se_max - nw_min) / 2 + 1);
}

static void o_anchor_widen(o_anchor *a, int width) { /* anchors.c:57-63 */
  a->x -= width / 2;
  a->y += width / 2;
  a->width += width;
}

static void o_anchor_x_range(const o_anchor *a, int x_len, int y_len, int y, int *x_min, int *x_max) {
  /* anchors.c:66-95 */
  (void)y_len;
  if (y < a->y)
    *x_min = 0;
  else if (y <= a->y + (a->length - 1))
    *x_min = (int)(a->x + (y - a->y));
  else
    *x_min = (int)(a->x + a->length);
  if (*x_min < 0) *x_min = 0;
  if (*x_min >= x_len) *x_min = x_len - 1;
  if (y < a->y - (a->width - 1))
    *x_max = (int)(a->x + (a->width - 1) - 1);
  else if (y <= a->y - (a->width - 1) + (a->length - 1))
    *x_max = (int)(a->x + (a->width - 1) + (y - (a->y - (a->width - 1))));
  else
    *x_max = x_len - 1;
  if (*x_max < 0) *x_max = 0;
  if (*x_max >= x_len) *x_max = x_len - 1;
}

static inline void o_anchor_reverse(o_anchor *a, int x_len, int y_len) { /* anchors.h:30-34 */
  a->x = -a->x + (x_len - 1) - (a->length - 1) - (a->width - 1);
  a->y = -a->y + (y_len - 1) - (a->length - 1) + (a->width - 1);
}

/* =============================================================================================
 * sw_full_ls (common/sw-full-ls.c): banded 3-state affine DP with back-pointers + traceback.
 * Back-pointer codes are the reference's FROM_* values (:36-42).
 * ===========================================================================================*/
enum { FR_N_N = 1, FR_N_NW = 2, FR_W_NW = 3, FR_W_W = 4, FR_NW_N = 5, FR_NW_NW = 6, FR_NW_W = 7 };
typedef struct o_cell {
  int n, w, nw;
  int8_t bn, bw, bnw;
} o_cell;
static const char o_ls_letters[] = "ACGTUMRWSYKVHDBN"; /* base_translate, fasta.c:694-696 */

static inline void o_init_cell(o_cell *c, int local, int bo, int ao) { /* init_cell :66-81 */
  if (local) {
    c->nw = 0;
    c->n = -bo;
    c->w = -ao;
  } else {
    c->nw = c->n = c->w = -INT_MAX / 2;
  }
  c->bnw = c->bn = c->bw = 0;
}

static int o_full_sw_ls(o_cell *mat, const int8_t *db, int lena, const int8_t *qr, int lenb, int threshscore,
                        int maxscore, int *iret, int *jret, int revcmpl, const o_anchor *anchor, int anchor_width,
                        int local, const orc_scores *sc, uint64_t *cells) {
  const int ao = -sc->a_gap_open, ae = -sc->a_gap_ext, bo = -sc->b_gap_open, be = -sc->b_gap_ext;
  const int match = sc->match, mismatch = sc->mismatch;
  int max_i = 0, max_j = 0, score = 0;
  o_anchor rect;
  if (anchor != NULL && anchor_width >= 0) { /* :175-177 */
    o_anchor_join(anchor, 1, &rect);
    o_anchor_widen(&rect, anchor_width);
  } else { /* threshold band :178-191 */
    o_anchor t[2];
    memset(t, 0, sizeof(t));
    t[0].x = 0;
    t[0].y = (lenb * match - threshscore) / match;
    t[0].length = 1;
    t[0].width = 1;
    t[1].x = lena - 1;
    t[1].y = lenb - 1 - t[0].y;
    t[1].length = 1;
    t[1].width = 1;
    o_anchor_join(t, 2, &rect);
  }
  const int W = lena + 1;
  for (int j = 0; j < lena + 1; j++) o_init_cell(&mat[j], 1, bo, ao); /* :194-196 */
  int i, j = 0;
  for (i = 0; i < lenb; i++) {
    int x_min, x_max;
    o_anchor_x_range(&rect, lena, lenb, i, &x_min, &x_max);
    o_init_cell(&mat[(i + 1) * W + x_min], local ? 1 : 0, bo, ao); /* :228-233 */
    if (cells) *cells += (uint64_t)(x_max - x_min + 1);
    for (j = x_min; j <= x_max; j++) {
      const o_cell *cnw = &mat[i * W + j], *cn = cnw + 1, *cw = cnw + W;
      o_cell *cur = &mat[(i + 1) * W + j + 1];
      int ms = (db[j] == qr[i]) ? match : mismatch;
      int tmp;
      int8_t t2;
      /* northwest :261-296 */
      if (!revcmpl) {
        tmp = cnw->nw + ms; t2 = FR_NW_NW;
        if (cnw->n + ms > tmp) { tmp = cnw->n + ms; t2 = FR_NW_N; }
        if (cnw->w + ms > tmp) { tmp = cnw->w + ms; t2 = FR_NW_W; }
      } else {
        tmp = cnw->w + ms; t2 = FR_NW_W;
        if (cnw->n + ms > tmp) { tmp = cnw->n + ms; t2 = FR_NW_N; }
        if (cnw->nw + ms > tmp) { tmp = cnw->nw + ms; t2 = FR_NW_NW; }
      }
      if (tmp <= 0 && local) tmp = t2 = 0;
      cur->nw = tmp; cur->bnw = t2;
      /* north :299-324 */
      if (!revcmpl) {
        tmp = cn->nw - bo - be; t2 = FR_N_NW;
        if (cn->n - be > tmp) { tmp = cn->n - be; t2 = FR_N_N; }
      } else {
        tmp = cn->n - be; t2 = FR_N_N;
        if (cn->nw - bo - be > tmp) { tmp = cn->nw - bo - be; t2 = FR_N_NW; }
      }
      if (tmp <= 0 && local) tmp = t2 = 0;
      cur->n = tmp; cur->bn = t2;
      /* west :327-352 */
      if (!revcmpl) {
        tmp = cw->nw - ao - ae; t2 = FR_W_NW;
        if (cw->w - ae > tmp) { tmp = cw->w - ae; t2 = FR_W_W; }
      } else {
        tmp = cw->w - ae; t2 = FR_W_W;
        if (cw->nw - ao - ae > tmp) { tmp = cw->nw - ao - ae; t2 = FR_W_NW; }
      }
      if (tmp <= 0 && local) tmp = t2 = 0;
      cur->w = tmp; cur->bw = t2;
      /* max score :357-368 */
      if (local || i == lenb - 1) {
        int t = imax(cur->n, cur->nw);
        t = imax(t, cur->w);
        if (t > score) { score = t; max_i = i; max_j = j; }
      }
      if (score == maxscore && local) break;
    }
    if (score == maxscore && local) break;
    if (i + 1 < lenb) { /* :376-383 */
      int nmin, nmax;
      o_anchor_x_range(&rect, lena, lenb, i + 1, &nmin, &nmax);
      for (j = x_max + 1; j <= nmax; j++) o_init_cell(&mat[(i + 1) * W + (j + 1)], local, bo, ao);
    }
  }
  *iret = max_i;
  *jret = max_j;
  if (score == maxscore || !local) return score;
  if (anchor != NULL) /* local retry with the threshold band :395-398 */
    return o_full_sw_ls(mat, db, lena, qr, lenb, threshscore, maxscore, iret, jret, revcmpl, NULL, anchor_width, local,
                        sc, cells);
  return 0;
}

static void o_sw_full_ls_core(const uint32_t *genome, int goff, int glen, const uint32_t *read, int rlen,
                              int threshscore, int maxscore, int revcmpl, const o_anchor *anchor, int anchor_width,
                              int local, const orc_scores *sc, orc_sfr *sfr, uint64_t *cells) {
  int8_t *db = (int8_t *)malloc(glen), *qr = (int8_t *)malloc(rlen);
  o_cell *mat = (o_cell *)malloc(sizeof(o_cell) * (size_t)(glen + 1) * (rlen + 1));
  int8_t *bt = (int8_t *)calloc(glen + rlen, 1);
  /* cells outside the band are never read (x_min/x_max are monotone); poison them to be sure */
  for (size_t k = 0; k < (size_t)(glen + 1) * (rlen + 1); k++) o_init_cell(&mat[k], 0, 0, 0);
  for (int i = 0; i < glen; i++) db[i] = (int8_t)EX4(genome, goff + i);
  for (int i = 0; i < rlen; i++) qr[i] = (int8_t)EX4(read, i);
  memset(sfr, 0, sizeof(*sfr));
  int i, j;
  sfr->score = o_full_sw_ls(mat, db, glen, qr, rlen, threshscore, maxscore, &i, &j, revcmpl, anchor, anchor_width,
                            local, sc, cells);
  /* do_backtrace :413-516 */
  const int W = glen + 1;
  const int ie = i, je = j;
  o_cell *cell = &mat[(i + 1) * W + j + 1];
  int from = cell->bnw, fromscore = cell->nw;
  if (cell->w > fromscore) { from = cell->bw; fromscore = cell->w; }
  if (cell->n > fromscore) from = cell->bn;
  int k = (glen + rlen) - 1;
  if (from != 0) {
    while (i >= 0 && j >= 0) {
      switch (from) {
        case FR_N_N: case FR_N_NW:
          bt[k] = 2; sfr->deletions++; sfr->read_start = i--; break;
        case FR_W_W: case FR_W_NW:
          bt[k] = 1; sfr->insertions++; sfr->genome_start = j--; break;
        default:
          bt[k] = 3;
          if (db[j] == qr[i]) sfr->matches++; else sfr->mismatches++;
          sfr->read_start = i--; sfr->genome_start = j--; break;
      }
      cell = &mat[(i + 1) * W + j + 1];
      switch (from) {
        case FR_N_N: from = cell->bn; break;
        case FR_N_NW: from = cell->bnw; break;
        case FR_W_W: from = cell->bw; break;
        case FR_W_NW: from = cell->bnw; break;
        case FR_NW_N: from = cell->bn; break;
        case FR_NW_NW: from = cell->bnw; break;
        default: from = cell->bw; break;
      }
      k--;
      if (from == 0) break;
    }
  }
  /* pretty_print :524-560 */
  {
    char *d = sfr->dbalign, *q = sfr->qralign;
    int ri = sfr->read_start, gj = sfr->genome_start, n = 0;
    for (int l = k + 1; l < glen + rlen && n < ORC_ALN_CAP - 1; l++, n++) {
      if (bt[l] == 2) { *d++ = '-'; *q++ = o_ls_letters[qr[ri++] & 15]; }
      else if (bt[l] == 1) { *d++ = o_ls_letters[db[gj++] & 15]; *q++ = '-'; }
      else if (bt[l] == 3) { *d++ = o_ls_letters[db[gj++] & 15]; *q++ = o_ls_letters[qr[ri++] & 15]; }
      else break;
    }
    *d = *q = 0;
  }
  sfr->gmapped = je - sfr->genome_start + 1;
  sfr->genome_start += goff;
  sfr->rmapped = ie - sfr->read_start + 1;
  free(db); free(qr); free(mat); free(bt);
}

void orc_sw_full_ls(const uint32_t *genome, int goff, int glen, const uint32_t *read, int rlen, int threshscore,
                    int maxscore, int revcmpl, long long ax, long long ay, int alen, int awidth, int anchor_width,
                    int local_alignment, const orc_scores *sc, orc_sfr *out) {
  o_anchor a;
  memset(&a, 0, sizeof(a));
  a.x = ax; a.y = ay; a.length = alen; a.width = awidth; a.weight = 1;
  o_sw_full_ls_core(genome, goff, glen, read, rlen, threshscore, maxscore, revcmpl, &a, anchor_width, local_alignment,
                    sc, out, NULL);
}

/* =============================================================================================
 * sw_full_cs (common/sw-full-cs.c): the letter genome against the four letter-space translations
 * of a colour read ("layers"), with crossovers between layers.
 * ===========================================================================================*/
typedef struct o_cscell {
  int n[4], w[4], nw[4];
  int8_t bn[4], bw[4], bnw[4]; /* FROM_x(layer, dir) = dir << 2 | layer, sw-full-cs.c:53 */
} o_cscell;
#define CSFROM(mat, dir) ((int8_t)(((dir) << 2) | (mat)))

static void o_cs_init_cell(o_cscell *c, int local, int xover, int bo, int ao) { /* init_cell :198-242 */
  for (int k = 0; k < 4; k++) {
    if (local) {
      int add = k == 0 ? 0 : xover;
      c->nw[k] = add;
      c->n[k] = -bo + add;
      c->w[k] = -ao + add;
    } else {
      c->nw[k] = c->n[k] = c->w[k] = -INT_MAX / 2;
    }
    c->bnw[k] = c->bn[k] = c->bw[k] = 0;
  }
}

static int o_full_sw_cs(o_cscell *mat, const int8_t *db, int lena, int8_t *const qr[4], int lenb, int threshscore,
                        int *iret, int *jret, int *kret, int revcmpl, const o_anchor *anchor, int anchor_width,
                        int local, const int *crossover_score, int global_xover, int indel_taboo_len,
                        const orc_scores *sc, uint64_t *cells) {
  const int ao = -sc->a_gap_open, ae = -sc->a_gap_ext, bo = -sc->b_gap_open, be = -sc->b_gap_ext;
  const int match = sc->match, mismatch = sc->mismatch;
  int max_i = 0, max_j = 0, max_k = 0, score = 0;
  const int W = lena + 1;
  o_anchor rect;
  for (int j = 0; j < lena + 1; j++) o_cs_init_cell(&mat[j], 1, global_xover, bo, ao); /* :268-270 */
  if (anchor != NULL && anchor_width >= 0) {
    o_anchor_join(anchor, 1, &rect);
    o_anchor_widen(&rect, anchor_width);
  } else {
    o_anchor t[2];
    memset(t, 0, sizeof(t));
    t[0].x = 0; t[0].y = (lenb * match - threshscore) / match; t[0].length = 1; t[0].width = 1;
    t[1].x = lena - 1; t[1].y = lenb - 1 - t[0].y; t[1].length = 1; t[1].width = 1;
    o_anchor_join(t, 2, &rect);
  }
  for (int i = 0; i < lenb; i++) {
    int x_min, x_max;
    const int xp = crossover_score == NULL ? global_xover : crossover_score[i];
    const int nt = i < lenb - indel_taboo_len; /* not in the indel taboo zone */
    o_anchor_x_range(&rect, lena, lenb, i, &x_min, &x_max);
    o_cs_init_cell(&mat[(i + 1) * W + x_min], local ? 1 : 0, xp, bo, ao); /* :317-325 */
    if (cells) *cells += (uint64_t)(x_max - x_min + 1);
    for (int j = x_min; j <= x_max; j++) {
      const o_cscell *cnw = &mat[i * W + j], *cn = cnw + 1, *cw = cnw + W;
      o_cscell *cur = &mat[(i + 1) * W + j + 1];
      for (int k = 0; k < 4; k++) {
        const int resetval = k != 0 ? xp : 0;
        int ms, tmp;
        int8_t t2;
        if (db[j] == 15 || qr[k][i] == 15) ms = 0;
        else ms = (db[j] == qr[k][i]) ? match : mismatch;
        /* northwest :362-437 */
        if (!revcmpl) {
          tmp = cnw->nw[k] + ms; t2 = CSFROM(k, FR_NW_NW);
          if (nt && cnw->n[k] + ms > tmp) { tmp = cnw->n[k] + ms; t2 = CSFROM(k, FR_NW_N); }
          if (cnw->w[k] + ms > tmp) { tmp = cnw->w[k] + ms; t2 = CSFROM(k, FR_NW_W); }
        } else {
          tmp = cnw->w[k] + ms; t2 = CSFROM(k, FR_NW_W);
          if (nt && cnw->n[k] + ms > tmp) { tmp = cnw->n[k] + ms; t2 = CSFROM(k, FR_NW_N); }
          if (cnw->nw[k] + ms > tmp) { tmp = cnw->nw[k] + ms; t2 = CSFROM(k, FR_NW_NW); }
        }
        for (int l = 0; l < 4; l++) {
          if (l == k) continue;
          if (!revcmpl) {
            if (cnw->nw[l] + ms + xp > tmp) { tmp = cnw->nw[l] + ms + xp; t2 = CSFROM(l, FR_NW_NW); }
            if (nt && cnw->n[l] + ms + xp > tmp) { tmp = cnw->n[l] + ms + xp; t2 = CSFROM(l, FR_NW_N); }
            if (cnw->w[l] + ms + xp > tmp) { tmp = cnw->w[l] + ms + xp; t2 = CSFROM(l, FR_NW_W); }
          } else {
            if (cnw->w[l] + ms + xp > tmp) { tmp = cnw->w[l] + ms + xp; t2 = CSFROM(l, FR_NW_W); }
            if (nt && cnw->n[l] + ms + xp > tmp) { tmp = cnw->n[l] + ms + xp; t2 = CSFROM(l, FR_NW_N); }
            if (cnw->nw[l] + ms + xp > tmp) { tmp = cnw->nw[l] + ms + xp; t2 = CSFROM(l, FR_NW_NW); }
          }
        }
        if (tmp <= resetval && local) { tmp = resetval; t2 = 0; }
        cur->nw[k] = tmp; cur->bnw[k] = t2;
        /* north :447-501 */
        if (!revcmpl) {
          tmp = cn->nw[k] - bo - be; t2 = CSFROM(k, FR_N_NW);
          if (!nt || cn->n[k] - be > tmp) { tmp = cn->n[k] - be; t2 = CSFROM(k, FR_N_N); }
        } else {
          tmp = cn->n[k] - be; t2 = CSFROM(k, FR_N_N);
          if (nt && cn->nw[k] - bo - be > tmp) { tmp = cn->nw[k] - bo - be; t2 = CSFROM(k, FR_N_NW); }
        }
        for (int l = 0; l < 4; l++) {
          if (l == k) continue;
          if (!revcmpl) {
            if (nt && cn->nw[l] - bo - be + xp > tmp) { tmp = cn->nw[l] - bo - be + xp; t2 = CSFROM(l, FR_N_NW); }
            if (cn->n[l] - be + xp > tmp) { tmp = cn->n[l] - be + xp; t2 = CSFROM(l, FR_N_N); }
          } else {
            if (cn->n[l] - be + xp > tmp) { tmp = cn->n[l] - be + xp; t2 = CSFROM(l, FR_N_N); }
            if (nt && cn->nw[l] - bo - be + xp > tmp) { tmp = cn->nw[l] - bo - be + xp; t2 = CSFROM(l, FR_N_NW); }
          }
        }
        if (tmp <= resetval && local) { tmp = resetval; t2 = 0; }
        cur->n[k] = tmp; cur->bn[k] = t2;
        /* west :511-545: no crossover on a genomic gap */
        if (!revcmpl) {
          tmp = cw->nw[k] - ao - ae; t2 = CSFROM(k, FR_W_NW);
          if (!nt || cw->w[k] - ae > tmp) { tmp = cw->w[k] - ae; t2 = CSFROM(k, FR_W_W); }
        } else {
          tmp = cw->w[k] - ae; t2 = CSFROM(k, FR_W_W);
          if (nt && cw->nw[k] - ao - ae > tmp) { tmp = cw->nw[k] - ao - ae; t2 = CSFROM(k, FR_W_NW); }
        }
        if (tmp <= resetval && local) { tmp = resetval; t2 = 0; }
        cur->w[k] = tmp; cur->bw[k] = t2;
        /* max score :552-580 */
        if (local || i == lenb - 1) {
          if (!revcmpl) {
            if (cur->nw[k] > score) { score = cur->nw[k]; max_i = i; max_j = j; max_k = k; }
            if (cur->n[k] > score) { score = cur->n[k]; max_i = i; max_j = j; max_k = k; }
            if (cur->w[k] > score) { score = cur->w[k]; max_i = i; max_j = j; max_k = k; }
          } else {
            if (cur->w[k] > score) { score = cur->w[k]; max_i = i; max_j = j; max_k = k; }
            if (cur->n[k] > score) { score = cur->n[k]; max_i = i; max_j = j; max_k = k; }
            if (cur->nw[k] > score) { score = cur->nw[k]; max_i = i; max_j = j; max_k = k; }
          }
        }
      }
    }
    if (i + 1 < lenb) { /* :604-612, still with the crossover penalty of colour i */
      int nmin, nmax;
      o_anchor_x_range(&rect, lena, lenb, i + 1, &nmin, &nmax);
      for (int j = x_max + 1; j <= nmax; j++) o_cs_init_cell(&mat[(i + 1) * W + (j + 1)], local, xp, bo, ao);
    }
  }
  *iret = max_i; *jret = max_j; *kret = max_k;
  return score;
}

void orc_sw_full_cs(const uint32_t *genome_ls, int goff, int glen, const uint32_t *read, int rlen, int initbp,
                    int threshscore, int revcmpl, long long ax, long long ay, int alen, int awidth,
                    int anchor_width, int indel_taboo_len, int local_alignment, const int *crossover_scores,
                    const orc_scores *sc, orc_sfr *sfr, uint64_t *cells) {
  int8_t *db = (int8_t *)malloc(glen);
  int8_t *qr[4];
  o_cscell *mat = (o_cscell *)malloc(sizeof(o_cscell) * (size_t)(glen + 1) * (rlen + 1));
  uint8_t *bt = (uint8_t *)calloc(glen + rlen + 1, 1);
  o_anchor a;
  memset(&a, 0, sizeof(a));
  a.x = ax; a.y = ay; a.length = alen; a.width = awidth; a.weight = 1;
  for (size_t q = 0; q < (size_t)(glen + 1) * (rlen + 1); q++) o_cs_init_cell(&mat[q], 0, 0, 0, 0);
  for (int i = 0; i < glen; i++) db[i] = (int8_t)EX4(genome_ls, goff + i);
  for (int k = 0; k < 4; k++) { /* :1182-1196: layer k starts from letter (k + initbp) % 4 */
    qr[k] = (int8_t *)malloc(rlen);
    int letter = (k + initbp) % 4;
    for (int j = 0; j < rlen; j++) {
      int base = EX4(read, j);
      if (base == 15) {
        qr[k][j] = 15;
        letter = (k + initbp) % 4;
      } else {
        /* cstols, util.h:157-180 */
        int r = (letter == 15 || base > 3) ? 15 : ((letter % 2 == 0) ? (4 + letter + base) % 4 : (4 + letter - base) % 4);
        qr[k][j] = (int8_t)r;
        letter = r;
      }
    }
  }
  memset(sfr, 0, sizeof(*sfr));
  int i, j, k;
  sfr->score = o_full_sw_cs(mat, db, glen, qr, rlen, threshscore, &i, &j, &k, revcmpl, &a, anchor_width,
                            local_alignment, crossover_scores, sc->crossover, indel_taboo_len, sc, cells);
  if (sfr->score >= 0 && sfr->score >= threshscore) {
    /* do_backtrace :633-937 */
    const int W = glen + 1;
    const int ie = i, je = j;
    int off = (glen + rlen) - 1;
    o_cscell *cell = &mat[(i + 1) * W + j + 1];
    int from = cell->bnw[k], fromscore = cell->nw[k];
    if (cell->w[k] > fromscore) { from = cell->bw[k]; fromscore = cell->w[k]; }
    if (cell->n[k] > fromscore) from = cell->bn[k];
    if (from != 0) {
      while (i >= 0 && j >= 0) {
        const int dir = from >> 2, lay = from & 3;
        if (dir == FR_N_N || dir == FR_N_NW) {
          sfr->deletions++; sfr->read_start = i--;
          bt[off] = (uint8_t)(2 + k); /* BACK_A_DELETION + k */
        } else if (dir == FR_W_W || dir == FR_W_NW) {
          sfr->insertions++; sfr->genome_start = j--;
          bt[off] = 1; /* BACK_INSERTION */
        } else {
          if (db[j] == qr[k][i] || db[j] == 15 || qr[k][i] == 15) sfr->matches++; else sfr->mismatches++;
          sfr->read_start = i--; sfr->genome_start = j--;
          bt[off] = (uint8_t)(6 + k); /* BACK_A_MATCH_MISMATCH + k */
        }
        if (k != lay) { bt[off] |= 0x80; sfr->crossovers++; k = lay; }
        cell = &mat[(i + 1) * W + j + 1];
        switch (dir) {
          case FR_N_N: from = cell->bn[k]; break;
          case FR_N_NW: from = cell->bnw[k]; break;
          case FR_W_W: from = cell->bw[k]; break;
          case FR_W_NW: from = cell->bnw[k]; break;
          case FR_NW_N: from = cell->bn[k]; break;
          case FR_NW_NW: from = cell->bnw[k]; break;
          default: from = cell->bw[k]; break;
        }
        off--;
        if (from == 0) break;
      }
    }
    off++;
    if (k != 0) { bt[off] |= 0x80; sfr->crossovers++; }
    /* pretty_print :945-1060 */
    {
      char *d = sfr->dbalign, *q = sfr->qralign;
      int ri = sfr->read_start, gj = sfr->genome_start, n = 0;
      for (int l = off; l < glen + rlen && n < ORC_ALN_CAP - 1; l++, n++) {
        const int ty = bt[l] & 0x0f, xo = bt[l] & 0x80;
        if (ty >= 2 && ty <= 5) {
          char c = o_ls_letters[qr[ty - 2][ri++] & 15];
          *d++ = '-';
          *q++ = xo ? (char)(c | 0x20) : c;
        } else if (ty == 1) {
          *d++ = o_ls_letters[db[gj++] & 15];
          *q++ = '-';
        } else if (ty >= 6 && ty <= 9) {
          char c = o_ls_letters[qr[ty - 6][ri++] & 15];
          *d++ = o_ls_letters[db[gj++] & 15];
          *q++ = xo ? (char)(c | 0x20) : c;
          if (*(q - 1) == 'n' || *(q - 1) == 'N') *(q - 1) = xo ? (char)(*(d - 1) | 0x20) : *(d - 1);
        } else {
          break;
        }
      }
      *d = *q = 0;
    }
    sfr->gmapped = je - sfr->genome_start + 1;
    sfr->genome_start += goff;
    sfr->rmapped = ie - sfr->read_start + 1;
  } else {
    sfr->score = 0;
  }
  for (int q = 0; q < 4; q++) free(qr[q]);
  free(db); free(mat); free(bt);
}

#include "oracle_pipeline.inc"
