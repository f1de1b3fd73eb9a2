/*
 * sw_shims.cpp -- replaces common/sw-vector.o, sw-gapless.o, sw-full-ls.o and sw-full-cs.o on the link line of
 * `gmapper` (reference Makefile:52-57).  Every symbol those four objects export is here, with the reference's C++
 * linkage (their headers have no extern "C") and exact signatures, served by libshrimp_b200.so:
 *
 *   sw_vector_setup / sw_vector / sw_vector_stats / sw_vector_cleanup        common/sw-vector.c:388,:453,:441,:379
 *   sw_gapless_setup / sw_gapless / sw_gapless_stats                         common/sw-gapless.c:28,:57,:46
 *   sw_full_ls_setup / sw_full_ls / sw_full_ls_stats / sw_full_ls_cleanup    common/sw-full-ls.c:573,:637,:625,:562
 *   sw_full_cs_setup / sw_full_cs / sw_full_cs_stats / sw_full_cs_cleanup    common/sw-full-cs.c:1076,:1146,:1134,:1062
 *
 * The *_setup / *_cleanup / *_stats functions are what gmapper.c itself calls (gmapper.c:2907-2965, :741-760,
 * :3031-3049).  The per-window entry points are a batch of ONE on the device: they exist so that the symbols the
 * reference's other callers expect resolve and so that the kernels can be checked call by call against the
 * reference objects (tests/test_gpu_shims.py); the production path is the chunk-level handle_read of
 * mapping_shim.cpp, whose counts the *_stats functions report.
 *
 * State is per thread, as in the reference (`#pragma omp threadprivate`, sw-vector.c:40).
 */
#include <vector>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common/sw-full-common.h"
#include "common/sw-full-cs.h"
#include "common/sw-full-ls.h"
#include "common/sw-gapless.h"
#include "common/sw-vector.h"

#include "shrimp_b200.h"
#include "shim_align.h"
#include "shim_state.h"

namespace shrimp_shim {

thread_local ThreadStats tstats;
shrimp_gpu_ctx *(*chunk_ctx_hook)() = nullptr;
void (*chunk_init_hook)() = nullptr;

struct CallState {
  shrimp_gpu_ctx *ctx = nullptr;
  shrimp_sw_params p;
  bool dirty = true;
  bool vector_set = false, gapless_set = false, full_ls_set = false, full_cs_set = false;
  int gapless_match = 0, gapless_mismatch = 0;
  uint64_t vec_invocs = 0, vec_cells = 0, gl_invocs = 0, gl_cells = 0, ls_invocs = 0, cs_invocs = 0;
  CallState() { memset(&p, 0, sizeof(p)); p.anchor_width = 8; }
};
static thread_local CallState S;

static void die(const char *what) {
  fprintf(stderr, "gmapper-b200: %s: %s\n", what, shrimp_gpu_last_error());
  exit(1);
}

// the device context of the per-call path, created on the first call that needs it
static shrimp_gpu_ctx *call_ctx() {
  if (!S.ctx && shrimp_gpu_create(0, &S.ctx) != SHRIMP_OK) die("shrimp_gpu_create");
  if (S.dirty) {
    const int rc = shrimp_gpu_sw_setup(S.ctx, &S.p);
    if (rc != SHRIMP_OK) die("shrimp_gpu_sw_setup");
    S.dirty = false;
  }
  return S.ctx;
}

static double stage_secs(const char *name) {
  shrimp_gpu_ctx *ctx = chunk_ctx_hook ? chunk_ctx_hook() : nullptr;
  double ms = 0;
  for (shrimp_gpu_ctx *c : {ctx, S.ctx}) {
    if (!c) continue;
    const char *names[16];
    float t[16];
    uint64_t l[16];
    const int n = shrimp_gpu_stage_times(c, names, t, l, 16);
    for (int i = 0; i < n; i++)
      if (!strcmp(names[i], name)) ms += t[i];
  }
  return ms / 1e3;
}

static void common_scores(int dblen, int qrlen, int a_open, int a_ext, int b_open, int b_ext, int match) {
  S.p.max_window_len = dblen;
  S.p.max_read_len = qrlen;
  S.p.a_gap_open = a_open;
  S.p.a_gap_ext = a_ext;
  S.p.b_gap_open = b_open;
  S.p.b_gap_ext = b_ext;
  S.p.match = match;
  S.dirty = true;
}

}  // namespace shrimp_shim

using namespace shrimp_shim;

// ================================================================================================
// sw-vector.o
// ================================================================================================
int sw_vector_setup(int dblen, int qrlen, int a_gap_open, int a_gap_ext, int b_gap_open, int b_gap_ext, int match,
                    int mismatch, int use_colours, bool reset_stats) {   // sw-vector.c:388
  if (match * qrlen >= 32768) {   // :393-398, same message, same exit
    fprintf(stderr, "Error: Match Value is too high/reads are too long. "
                    "Please ensure that (Match_Value x your_longest_read_length)"
                    " is less than 32768! Try using smaller S-W values.");
    exit(1);
  }
  common_scores(dblen, qrlen, a_gap_open, a_gap_ext, b_gap_open, b_gap_ext, match);
  S.p.use_colours = use_colours ? 1 : 0;
  if (use_colours) {
    // the colour filter is handed match + crossover as its mismatch (gmapper.c:2935); the library derives exactly that
    // from the crossover score, sw_full_cs_setup supplies the letter mismatch
    S.p.crossover = mismatch - match;
    if (!S.full_cs_set) S.p.mismatch = mismatch;
  } else {
    S.p.mismatch = mismatch;
  }
  if (reset_stats) S.vec_invocs = S.vec_cells = 0;
  S.vector_set = true;
  return 0;
}

int sw_vector_cleanup(void) {   // sw-vector.c:379
  S.vector_set = false;
  if (getenv("SHRIMP_B200_VERBOSE") && tstats.batches)
    fprintf(stderr,
            "[gmapper-b200] thread: %llu reads in %llu device batches (%llu re-mapped alone), %llu records; host seconds: "
            "look-ahead %.3f, device calls %.3f (the first %.3f), record rebuild %.3f, output %.3f\n",
            (unsigned long long)tstats.reads, (unsigned long long)tstats.batches, (unsigned long long)tstats.mispredicted,
            (unsigned long long)tstats.records, tstats.t_prep, tstats.t_device, tstats.t_first_device, tstats.t_build, tstats.t_output);
  return 0;
}

void sw_vector_stats(uint64_t *invocs, uint64_t *cells, double *secs) {   // sw-vector.c:441
  if (invocs) *invocs = S.vec_invocs + tstats.vector_calls;
  if (cells) *cells = S.vec_cells + tstats.vector_cells;
  if (secs) *secs = stage_secs("sw_vector");
}

int sw_vector(uint32_t *genome, int goff, int glen, uint32_t *read, int rlen, uint32_t *genome_ls, int initbp,
              bool is_rna) {   // sw-vector.c:453
  if (!S.vector_set) abort();   // :462
  if (is_rna) {
    fprintf(stderr, "gmapper-b200: sw_vector: RNA genomes are not served by the GPU path\n");
    exit(1);
  }
  shrimp_gpu_ctx *ctx = call_ctx();
  const uint32_t go = (uint32_t)goff;
  const int32_t gl = glen, ri = 0, rl = rlen;
  const int8_t ib = (int8_t)initbp;
  int32_t score = 0;
  if (shrimp_gpu_sw_vector_batch(ctx, genome, ((size_t)goff + glen + 7) / 8, genome_ls, read, (rlen + 7) / 8, 1, 1, &go,
                                 &gl, &ri, &rl, genome_ls ? &ib : nullptr, &score) != SHRIMP_OK)
    die("sw_vector");
  S.vec_invocs++;
  S.vec_cells += (uint64_t)glen * rlen;   // :509
  return score;
}

// ================================================================================================
// sw-gapless.o
// ================================================================================================
int sw_gapless_setup(int match, int mismatch, bool reset_stats) {   // sw-gapless.c:28
  S.gapless_match = match;
  S.gapless_mismatch = mismatch;
  if (reset_stats) S.gl_invocs = S.gl_cells = 0;
  S.gapless_set = true;
  return 0;
}

void sw_gapless_stats(uint64_t *invocs, uint64_t *cells, uint64_t *ticks) {   // sw-gapless.c:46
  if (invocs) *invocs = S.gl_invocs + tstats.vector_calls;
  if (cells) *cells = S.gl_cells + tstats.vector_cells;
  if (ticks) *ticks = 0;
}

int sw_gapless(uint32_t *genome, int glen, uint32_t *read, int rlen, int g_idx, int r_idx, uint32_t *genome_ls,
               int init_bp, bool is_rna) {   // sw-gapless.c:57
  if (!S.gapless_set) abort();   // :67
  if (is_rna) {
    fprintf(stderr, "gmapper-b200: sw_gapless: RNA genomes are not served by the GPU path\n");
    exit(1);
  }
  // the gapless kernel scores with the set-up's match and vector mismatch: make them sw_gapless_setup's
  if (S.p.match != S.gapless_match || (S.p.use_colours ? S.p.match + S.p.crossover : S.p.mismatch) != S.gapless_mismatch ||
      S.p.max_read_len < rlen) {
    S.p.match = S.gapless_match;
    if (S.p.use_colours)
      S.p.crossover = S.gapless_mismatch - S.gapless_match;
    else
      S.p.mismatch = S.gapless_mismatch;
    if (S.p.max_read_len < rlen) S.p.max_read_len = rlen;
    if (S.p.max_window_len < 1) S.p.max_window_len = rlen;
    S.dirty = true;
  }
  shrimp_gpu_ctx *ctx = call_ctx();
  const uint32_t go = 0;
  const int32_t gl = glen, ri = 0, rl = rlen, gi = g_idx, rj = r_idx;
  const int8_t ib = (int8_t)init_bp;
  int32_t score = 0;
  if (shrimp_gpu_sw_gapless_batch(ctx, genome, ((size_t)glen + 7) / 8, genome_ls, read, (rlen + 7) / 8, 1, 1, &go, &gl,
                                  &ri, &rl, &gi, &rj, genome_ls ? &ib : nullptr, &score) != SHRIMP_OK)
    die("sw_gapless");
  S.gl_invocs++;
  S.gl_cells += (uint64_t)rlen;   // :111
  return score;
}

// ================================================================================================
// sw-full-ls.o / sw-full-cs.o
// ================================================================================================
int sw_full_ls_setup(int dblen, int qrlen, int a_gap_open, int a_gap_ext, int b_gap_open, int b_gap_ext, int match,
                     int mismatch, bool reset_stats, int anchor_width) {   // sw-full-ls.c:573
  common_scores(dblen, qrlen, a_gap_open, a_gap_ext, b_gap_open, b_gap_ext, match);
  if (!S.p.use_colours) S.p.mismatch = mismatch;
  S.p.anchor_width = anchor_width;
  if (reset_stats) S.ls_invocs = 0;
  S.full_ls_set = true;
  if (chunk_init_hook) chunk_init_hook();
  return 0;
}

int sw_full_ls_cleanup(void) {   // sw-full-ls.c:562
  S.full_ls_set = false;
  return 0;
}

int sw_full_cs_setup(int dblen, int qrlen, int a_gap_open, int a_gap_ext, int b_gap_open, int b_gap_ext, int match,
                     int mismatch, int crossover, bool reset_stats, int anchor_width, int indel_taboo_len) {
  // sw-full-cs.c:1076
  common_scores(dblen, qrlen, a_gap_open, a_gap_ext, b_gap_open, b_gap_ext, match);
  S.p.use_colours = 1;
  S.p.mismatch = mismatch;
  S.p.crossover = crossover;
  S.p.anchor_width = anchor_width;
  S.p.indel_taboo_len = indel_taboo_len;
  if (reset_stats) S.cs_invocs = 0;
  S.full_cs_set = true;
  if (chunk_init_hook) chunk_init_hook();
  return 0;
}

int sw_full_cs_cleanup(void) {   // sw-full-cs.c:1062
  S.full_cs_set = false;
  return 0;
}

// In gmapper exactly one of the two runs (shrimp_mode); both report the chunk path's full-SW counts.
void sw_full_ls_stats(uint64_t *invocs, uint64_t *cells, double *secs) {   // sw-full-ls.c:625
  if (invocs) *invocs = S.ls_invocs + (S.p.use_colours ? 0 : tstats.full_calls);
  if (cells) *cells = S.p.use_colours ? 0 : tstats.full_cells;
  if (secs) *secs = S.p.use_colours ? 0.0 : stage_secs("sw_full");
}

void sw_full_cs_stats(uint64_t *invocs, uint64_t *cells, double *secs) {   // sw-full-cs.c:1134
  if (invocs) *invocs = S.cs_invocs + (S.p.use_colours ? tstats.full_calls : 0);
  if (cells) *cells = S.p.use_colours ? tstats.full_cells : 0;
  if (secs) *secs = S.p.use_colours ? stage_secs("sw_full") : 0.0;
}

static void full_call(bool cs, uint32_t *genome, int goff, int glen, uint32_t *read, int rlen, int initbp,
                      int threshscore, int maxscore, struct sw_full_results *sfr, bool revcmpl, struct anchor *anchors,
                      int anchors_cnt, int local_alignment, int *crossover_score) {
  struct sw_full_results scratch;
  if (sfr == NULL) {   // sw-full-ls.c:654-657
    sfr = &scratch;
    memset(sfr, 0, sizeof(*sfr));
  }
  if (S.p.use_colours != (cs ? 1 : 0)) {
    S.p.use_colours = cs ? 1 : 0;
    S.dirty = true;
  }
  shrimp_gpu_ctx *ctx = call_ctx();
  struct anchor a;
  anchor_join(anchors, anchors_cnt, &a);   // the reference's own, anchors.c:9-54 (full_sw does the same, sw-full-ls.c:170)
  shrimp_full_task t;
  memset(&t, 0, sizeof(t));
  t.goff = (uint32_t)goff;
  t.glen = glen;
  t.read_idx = 0;
  t.rlen = rlen;
  t.threshscore = threshscore;
  t.maxscore = maxscore;
  t.revcmpl = revcmpl ? 1 : 0;
  t.ax = (int32_t)a.x;
  t.ay = (int32_t)a.y;
  t.alen = a.length;
  t.awidth = a.width;
  t.initbp = initbp;
  shrimp_full_result r;
  std::vector<uint8_t> ed((size_t)glen + rlen + 16);
  int64_t used = 0;
  if (shrimp_gpu_sw_full_batch_xover(ctx, genome, ((size_t)goff + glen + 7) / 8, read, (rlen + 7) / 8, 1, 1, &t,
                                     local_alignment, cs ? crossover_score : nullptr, rlen, &r, ed.data(),
                                     (int64_t)ed.size(), &used) != SHRIMP_OK)
    die(cs ? "sw_full_cs" : "sw_full_ls");
  sfr->score = r.score;
  if (cs && !(r.score > 0 || r.edit_len > 0)) {   // below the threshold: no traceback, score 0 (sw-full-cs.c:1216-1226)
    sfr->score = 0;
    return;
  }
  sfr->read_start = r.read_start;
  sfr->rmapped = r.rmapped;
  sfr->genome_start = r.genome_start;
  sfr->gmapped = r.gmapped;
  sfr->matches = r.matches;
  sfr->mismatches = r.mismatches;
  sfr->insertions = r.insertions;
  sfr->deletions = r.deletions;
  sfr->crossovers = r.crossovers;
  char *db = (char *)xmalloc((size_t)r.edit_len + 1), *qr = (char *)xmalloc((size_t)r.edit_len + 1);
  edit_to_strings(ed.data() + r.edit_off, r.edit_len, genome, r.genome_start, read, r.read_start, cs, initbp, rlen, db, qr);
  sfr->dbalign = db;
  sfr->qralign = qr;
}

void sw_full_ls(uint32_t *genome, int goff, int glen, uint32_t *read, int rlen, int threshscore, int maxscore,
                struct sw_full_results *sfr, bool revcmpl, struct anchor *anchors, int anchors_cnt,
                int local_alignment) {   // sw-full-ls.c:637
  if (!S.full_ls_set) abort();   // :647
  S.ls_invocs++;
  full_call(false, genome, goff, glen, read, rlen, 0, threshscore, maxscore, sfr, revcmpl, anchors, anchors_cnt,
            local_alignment, nullptr);
}

void sw_full_cs(uint32_t *genome_ls, int goff, int glen, uint32_t *read, int rlen, int initbp, int threshscore,
                struct sw_full_results *sfr, bool revcmpl, bool is_rna, struct anchor *anchors, int anchors_cnt,
                int local_alignment, int *crossover_score) {   // sw-full-cs.c:1146
  if (!S.full_cs_set) abort();   // :1157
  if (is_rna) {
    fprintf(stderr, "gmapper-b200: sw_full_cs: RNA genomes are not served by the GPU path\n");
    exit(1);
  }
  S.cs_invocs++;
  full_call(true, genome_ls, goff, glen, read, rlen, initbp, threshscore, 0, sfr, revcmpl, anchors, anchors_cnt,
            local_alignment, crossover_score);
}
