/*
 * mapping_shim.cpp -- replaces gmapper/mapping.o on the link line of `gmapper` (reference Makefile:52-57).
 *
 * It carries the three symbols mapping.o exports, with the reference's C++ linkage and signatures:
 *     void handle_read(read_entry *, read_mapping_options_t *, int)          gmapper/mapping.c:1773
 *     void handle_readpair(pair_entry *, readpair_mapping_options_t *, int)  gmapper/mapping.c:2504
 *     int  get_insert_size(read_hit *, read_hit *)                           gmapper/mapping.c:405
 * and serves them from libshrimp_b200.so (include/shrimp_b200.h).  gmapper.c, genome.c, output.c, fasta.c ... are
 * the reference's own, compiled unchanged from where they lie (integration/Makefile); SAM emission is the
 * unchanged read_output / readpair_output / hit_output of gmapper/output.c:955,:1070,:227.
 *
 * How a per-read interface feeds a batched device path without touching gmapper.c: launch_scan_threads
 * (gmapper.c:286-645) loads a whole chunk of `-K` reads into a contiguous re_buffer[] before it walks it and
 * calls handle_read for one entry after the other.  The first call of a chunk therefore sees the rest of the
 * chunk behind its own entry: it looks ahead, prepares the entries that follow exactly as the loop at
 * gmapper.c:411-553 will (trimming, length and quality filters, packing, crossover scores), maps the whole
 * look-ahead on the GPU in one shrimp_gpu_map_reads / shrimp_gpu_map_pairs call and keeps the records; every
 * later call of the chunk only picks its records up, rebuilds struct read_hit / sw_full_results from them and
 * hands them to the unchanged output code.  When an entry arrives whose actual state differs from the
 * prediction, it is mapped again on its own (a batch of one): still the device path, never a CPU fall-back.
 *
 * One context per OpenMP thread (`-N`); thread t works on device t % n_devices; the first thread on a device
 * uploads the genome arrays load_genome built (gmapper.h:264-275) and builds the projection in HBM, the other
 * threads on that device share it (shrimp_gpu_share_genome).
 */
#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include <execinfo.h>
#include <malloc.h>
#include <math.h>
#include <signal.h>
#include <unistd.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "gmapper/mapping.h"
#include "gmapper/output.h"
#include "common/sw-full-common.h"

#include "shrimp_b200.h"
#include "shim_align.h"
#include "shim_lookahead.h"
#include "shim_state.h"

namespace shrimp_shim {

static void die(const char *what) {
  fprintf(stderr, "gmapper-b200: %s: %s\n", what, shrimp_gpu_last_error());
  exit(1);
}

static void unsupported(const char *what) {
  fprintf(stderr, "gmapper-b200: %s is not served by the GPU path (there is no CPU fall-back)\n", what);
  exit(1);
}

// ------------------------------------------------------------------------------------------------
// Contexts
// ------------------------------------------------------------------------------------------------
static std::mutex g_dev_mu[64];
static shrimp_gpu_ctx *g_dev_owner[64];
static int g_ndev = -1;
static std::once_flag g_ndev_once;

static thread_local shrimp_gpu_ctx *t_ctx;
static shrimp_gpu_ctx *ctx_if_any() { return t_ctx; }
shrimp_gpu_ctx *thread_ctx();
static void warm_up(shrimp_gpu_ctx *ctx);
static void init_thread_ctx() {
  // only once the genome is in memory and this is a mapping run (gmapper -S exits before the set-up calls)
  if (genome_contigs != NULL && num_contigs > 0 && n_seeds > 0) {
    const bool fresh = t_ctx == nullptr;
    shrimp_gpu_ctx *ctx = thread_ctx();
    if (fresh) warm_up(ctx);
  }
}
// gmapper.c allocates a 10 MB output buffer per chunk and grows it in 10 MB steps (gmapper.c:403, output.c:246-268):
// at GPU rates that is an mmap / page-fault / munmap cycle of tens of megabytes per thread every few milliseconds, all
// of them serialised on the process's address-space lock.  Keep such blocks on the heap and never trim it.
static void segv_backtrace(int sig) {   // SHRIMP_B200_BACKTRACE=1: where a crash happened (no debugger on the GPU boxes)
  void *frames[64];
  const int n = backtrace(frames, 64);
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}
static bool tune_malloc() {
  if (getenv("SHRIMP_B200_BACKTRACE")) signal(SIGSEGV, segv_backtrace);
  if (!getenv("SHRIMP_B200_NO_MALLOPT")) {
    mallopt(M_MMAP_THRESHOLD, 32 << 20);   // the largest value glibc accepts
    mallopt(M_TRIM_THRESHOLD, -1);
    mallopt(M_TOP_PAD, 64 << 20);
  }
  return true;
}
static const bool g_hooked = ((chunk_ctx_hook = ctx_if_any), (chunk_init_hook = init_thread_ctx), tune_malloc());

static shrimp_sw_params sw_params_from_globals() {
  shrimp_sw_params sp;
  memset(&sp, 0, sizeof(sp));
  const bool cs = shrimp_mode == MODE_COLOUR_SPACE;
  sp.match = match_score;
  sp.mismatch = mismatch_score;  // sw_full_*_setup's value; the library derives the colour filter's match + crossover
  sp.a_gap_open = a_gap_open_score;
  sp.a_gap_ext = a_gap_extend_score;
  sp.b_gap_open = b_gap_open_score;
  sp.b_gap_ext = b_gap_extend_score;
  sp.crossover = cs ? crossover_score : 0;
  sp.use_colours = cs ? 1 : 0;
  sp.anchor_width = anchor_width;
  sp.indel_taboo_len = cs ? indel_taboo_len : 0;
  sp.max_read_len = longest_read_len;
  sp.max_window_len = (int)abs_or_pct(window_len, longest_read_len);  // gmapper.c:2865
  return sp;
}

shrimp_gpu_ctx *thread_ctx() {
  if (t_ctx) return t_ctx;
  std::call_once(g_ndev_once, [] {
    // every kernel of the library is loaded when the context is created (in the set-up phase), not at its first launch
    setenv("CUDA_MODULE_LOADING", "EAGER", 0);
    g_ndev = shrimp_gpu_device_count();
    if (const char *e = getenv("SHRIMP_B200_GPUS")) g_ndev = std::min(g_ndev, std::max(1, atoi(e)));
    if (g_ndev > 64) g_ndev = 64;
    if (g_ndev <= 0) {
      fprintf(stderr, "gmapper-b200: no usable CUDA device (there is no CPU fall-back)\n");
      exit(1);
    }
  });
  if (genome_is_rna) unsupported("an RNA genome (uracil)");
  if (!Cflag || !Fflag) unsupported("-C / -F (one strand only)");
  const int dev = omp_get_thread_num() % g_ndev;
  shrimp_gpu_ctx *ctx = nullptr;
  if (shrimp_gpu_create(dev, &ctx) != SHRIMP_OK) die("shrimp_gpu_create");
  shrimp_sw_params sp = sw_params_from_globals();
  const int rc = shrimp_gpu_sw_setup(ctx, &sp);
  if (rc != SHRIMP_OK) die("shrimp_gpu_sw_setup");
  {
    std::lock_guard<std::mutex> lk(g_dev_mu[dev]);
    if (!g_dev_owner[dev]) {
      // the genome arrays of load_genome / load_genome_map (genome.c:1092-1124, :670-832) -> HBM, then the
      // projection of genome.c:1138-1166 built there with the seeds of seeds.c
      if (shrimp_gpu_genome_load(ctx, num_contigs, genome_contigs, genome_len, shrimp_mode == MODE_COLOUR_SPACE) !=
          SHRIMP_OK)
        die("shrimp_gpu_genome_load");
      std::vector<uint64_t> masks(n_seeds);
      std::vector<int32_t> spans(n_seeds), weights(n_seeds);
      for (int sn = 0; sn < n_seeds; sn++) {
        masks[sn] = seed[sn].mask[0];
        spans[sn] = seed[sn].span;
        weights[sn] = seed[sn].weight;
      }
      if (shrimp_gpu_index_build(ctx, n_seeds, masks.data(), spans.data(), weights.data(), Hflag ? 1 : 0) != SHRIMP_OK)
        die("shrimp_gpu_index_build");
      g_dev_owner[dev] = ctx;
    } else if (shrimp_gpu_share_genome(ctx, g_dev_owner[dev]) != SHRIMP_OK) {
      die("shrimp_gpu_share_genome");
    }
  }
  t_ctx = ctx;
  return ctx;
}

// ------------------------------------------------------------------------------------------------
// Set-up phase: the context's chunk buffers.  The reference allocates each thread's DP matrices, region maps and
// window caches inside its set-up parallel region, before the mapping clock starts (gmapper.c:2907-2965); the device
// path's counterpart are the chunk buffers of the context (device slabs, pinned staging), which the library sizes at
// its first batch.  So that they, too, exist before "Read Mapping Time" starts, every thread maps one synthetic chunk
// here: chunk_size reads of the read file's read length cut from the genome itself (they map, so every stage sizes
// its buffers as a real chunk will).  Nothing of it is output.  SHRIMP_B200_NO_WARM_UP=1 skips it.
// ------------------------------------------------------------------------------------------------
static int first_read_length() {
  const char *fn = single_reads_file ? reads_filename : left_reads_filename;
  if (!fn || !strcmp(fn, "-")) return 0;
  gzFile f = gzopen(fn, "r");
  if (!f) return 0;
  char line[1 << 16];
  int len = 0;
  bool named = false;
  while (gzgets(f, line, sizeof(line))) {
    if (line[0] == '#') continue;
    if (!named) {
      if (line[0] != '>' && line[0] != '@') break;
      named = true;
      continue;
    }
    len = (int)strcspn(line, "\r\n");
    break;
  }
  gzclose(f);
  return len;
}

static shrimp_map_params map_params_from_globals();
static void warm_up(shrimp_gpu_ctx *ctx) {
  if (getenv("SHRIMP_B200_NO_WARM_UP")) return;
  const bool cs = shrimp_mode == MODE_COLOUR_SPACE, paired = pair_mode != PAIR_NONE;
  if (cs && (paired || (Qflag && !ignore_qvs))) return;   // (per-position crossover scores: sized at the first real chunk)
  int L = first_read_length();
  if (cs) L--;
  if (L < 15 || L > longest_read_len) return;
  int n = std::min(chunk_size, 1 << 17) & ~1;
  std::vector<int> usable;
  for (int cn = 0; cn < num_contigs; cn++)
    if ((long long)genome_len[cn] > 2LL * L + 400) usable.push_back(cn);
  if (usable.empty() || n < 2) return;
  const int stride = BPTO32BW(L) + 1;
  std::vector<uint32_t> reads((size_t)n * stride, 0);
  std::vector<int32_t> rlen((size_t)n, L);
  std::vector<int8_t> initbp((size_t)n, BASE_T);
  unsigned long long seed = 0x9E3779B97F4A7C15ull * (unsigned long long)(omp_get_thread_num() + 1);
  auto rnd = [&seed]() {
    seed ^= seed << 13;
    seed ^= seed >> 7;
    seed ^= seed << 17;
    return seed;
  };
  auto put = [&](int row, int j, uint32_t v) { reads[(size_t)row * stride + (j >> 3)] |= (v & 15u) << (4 * (j & 7)); };
  for (int row = 0; row < n; row++) {
    const int cn = usable[rnd() % usable.size()];
    const uint32_t *g = genome_contigs[cn];
    const bool second = paired && (row & 1);
    // a pair: the mates 200 bases apart on opposite strands; the second mate is the reverse complement
    static thread_local long long pos1;
    long long pos = second ? pos1 + 200 : 1 + (long long)(rnd() % (unsigned long long)(genome_len[cn] - 2 * L - 300));
    if (!second) pos1 = pos;
    if (pos + L >= (long long)genome_len[cn]) pos = (long long)genome_len[cn] - L - 1;
    if (!cs) {
      for (int j = 0; j < L; j++) {
        const uint32_t b = EXTRACT(g, pos + (second ? L - 1 - j : j));
        put(row, j, second ? (uint32_t)complement_base((int)b, false) : b);
      }
    } else {   // T + the colours between consecutive genome letters, the first one against T (fasta.c:587-605)
      int prev = BASE_T;
      for (int j = 0; j < L; j++) {
        const int b = (int)EXTRACT(g, pos + j);
        put(row, j, (uint32_t)lstocs(prev, b, false));
        prev = b;
      }
    }
  }
  shrimp_map_params mp = map_params_from_globals();
  shrimp_map_stats st;
  memset(&st, 0, sizeof(st));
  int64_t n_hits = 0, e_used = 0;
  std::vector<int32_t> n_unp((size_t)n, 0);
  std::vector<uint8_t> edits((size_t)n * 2 * (size_t)L + 4096);
  int rc;
  if (!paired) {
    std::vector<shrimp_hit> hits((size_t)n * num_outputs);
    rc = shrimp_gpu_map_reads(ctx, &mp, n, reads.data(), stride, rlen.data(), cs ? initbp.data() : nullptr, hits.data(),
                              (int64_t)hits.size(), n_unp.data(), edits.data(), (int64_t)edits.size(), &n_hits, &e_used,
                              nullptr, 0, nullptr, &st);
  } else {
    const int np = n / 2;
    shrimp_pair_params pp = {pair_mode, min_insert_size, max_insert_size, half_paired ? 1 : 0};
    std::vector<shrimp_hit> hits((size_t)np * num_outputs * 4);
    std::vector<shrimp_pair> pairs((size_t)np * num_outputs);
    std::vector<int32_t> n_pairs((size_t)np, 0);
    int64_t n_pairs_out = 0;
    rc = shrimp_gpu_map_pairs(ctx, &mp, &pp, np, reads.data(), stride, rlen.data(), nullptr, hits.data(),
                              (int64_t)hits.size(), &n_hits, pairs.data(), (int64_t)pairs.size(), &n_pairs_out,
                              n_pairs.data(), n_unp.data(), edits.data(), (int64_t)edits.size(), &e_used, &st);
  }
  if (rc != SHRIMP_OK && rc != SHRIMP_E_NOMEM && getenv("SHRIMP_B200_VERBOSE"))
    fprintf(stderr, "[gmapper-b200] warm-up chunk: %s (ignored)\n", shrimp_gpu_last_error());
  shrimp_gpu_stage_times_reset(ctx);   // the statistics of print_statistics count the mapping only
}

// ------------------------------------------------------------------------------------------------
// Options: the reference's option structs -> shrimp_map_params.  Only the option sets gmapper.c builds by
// default (gmapper.c:2588-2718) are served; anything else stops the run.
// ------------------------------------------------------------------------------------------------
static shrimp_map_params map_params_from_globals() {
  shrimp_map_params mp;
  memset(&mp, 0, sizeof(mp));
  mp.window_len = window_len;
  mp.window_overlap = window_overlap;
  mp.window_gen_threshold = window_gen_threshold;
  mp.sw_vect_threshold = sw_vect_threshold;
  mp.sw_full_threshold = sw_full_threshold;
  mp.score_alpha = score_alpha;
  mp.score_beta = score_beta;
  mp.match_mode = match_mode;
  mp.num_outputs = num_outputs;
  mp.num_tmp_outputs = num_tmp_outputs;
  mp.gapless = gapless_sw ? 1 : 0;
  mp.hash_filter_calls = hash_filter_calls ? 1 : 0;
  mp.use_regions = use_regions ? 1 : 0;
  mp.region_bits = region_bits;
  mp.region_overlap = region_overlap;
  mp.Gflag = Gflag ? 1 : 0;
  mp.Tflag = Tflag ? 1 : 0;
  mp.strata = strata_flag ? 1 : 0;
  mp.max_alignments = max_alignments;
  mp.compute_mapping_qualities = compute_mapping_qualities ? 1 : 0;
  mp.list_cutoff = list_cutoff;
  mp.qual_delta = qual_delta;
  mp.qual_vector_offset = qual_vector_offset;
  mp.use_sanger_qvs = use_sanger_qvs ? 1 : 0;
  mp.pr_xover = pr_xover;
  return mp;
}

// One option set of an --unpaired-options list (gmapper.c:1632-1717) as parameters of a device call.  Served: the
// sets that compute every structure afresh (regions when they are used, anchor list, hit list, pass 1) -- what the
// reference's own default is, with any thresholds, match mode, window overlap and output counts -- because a device
// call keeps no per-read state from one option set to the next.  A set that reuses the previous set's anchor or hit
// list (`recompute` 0) stops the run.
static shrimp_map_params map_params_from_globals();
static shrimp_map_params stage_params(const read_mapping_options_t &o, int k) {
  if (!o.anchor_list.recompute || !o.hit_list.recompute || !o.pass1.recompute)
    unsupported("an --unpaired-options set that reuses the previous set's anchor list, hit list or pass 1 (recompute 0)");
  if (o.anchor_list.use_region_counts && !o.regions.recompute && k == 0)
    unsupported("an --unpaired-options set that uses region counts it does not compute");
  if (!o.anchor_list.collapse) unsupported("an --unpaired-options set without anchor collapsing");
  // (anchor_list.use_mp_region_counts and pass1.only_paired are not parsed for an unpaired set, gmapper.c:1651-1654,
  // :1686-1689: the reference leaves them as the allocator gave them, and they mean nothing for a read without a mate)
  if (o.hit_list.match_mode != 1 && o.hit_list.match_mode != 2) unsupported("an unpaired hit-list match mode other than 1 or 2");
  if (o.pass1.min_matches != o.hit_list.match_mode)
    unsupported("an --unpaired-options set whose pass-1 min_matches differs from its hit-list match mode");
  if (o.pass1.gapless != o.hit_list.gapless) unsupported("an --unpaired-options set that is gapless in one of hit list / pass 1 only");
  if (o.anchor_list.use_region_counts && o.hit_list.match_mode != 2)
    unsupported("region counts with a hit-list match mode other than 2");
  shrimp_map_params mp = map_params_from_globals();
  mp.match_mode = o.hit_list.match_mode;
  mp.use_regions = o.anchor_list.use_region_counts ? 1 : 0;
  mp.gapless = o.hit_list.gapless ? 1 : 0;
  mp.window_gen_threshold = o.hit_list.threshold;
  mp.sw_vect_threshold = o.pass1.threshold;
  mp.window_overlap = o.pass1.window_overlap;
  mp.num_tmp_outputs = o.pass1.num_outputs;
  mp.sw_full_threshold = o.pass2.threshold;
  mp.strata = o.pass2.strata ? 1 : 0;
  mp.num_outputs = o.pass2.num_outputs;
  return mp;
}

// true: the list is not the single default set gmapper.c builds from the globals (gmapper.c:2601-2632) but one the
// user gave with --unpaired-options; every set of it must be one stage_params() accepts
static bool custom_unpaired_options(const read_mapping_options_t *o, int n_options) {
  const bool rc = (match_mode == 2 && use_regions);
  if (n_options == 1 && o->regions.recompute == rc && o->anchor_list.recompute && o->anchor_list.collapse &&
      o->anchor_list.use_region_counts == rc && o->anchor_list.use_mp_region_counts == 0 && o->hit_list.recompute &&
      o->hit_list.gapless == gapless_sw && o->hit_list.match_mode == match_mode &&
      o->hit_list.threshold == window_gen_threshold && o->pass1.recompute && !o->pass1.only_paired &&
      o->pass1.gapless == gapless_sw && o->pass1.min_matches == match_mode && o->pass1.num_outputs == num_tmp_outputs &&
      o->pass1.threshold == sw_vect_threshold && o->pass1.window_overlap == window_overlap &&
      o->pass2.strata == strata_flag && o->pass2.num_outputs == num_outputs && o->pass2.threshold == sw_full_threshold &&
      o->pass2.stop_count == 0)
    return false;
  for (int k = 0; k < n_options; k++) {
    (void)stage_params(o[k], k);
    if (n_options > 1 && o[k].pass2.save_outputs) unsupported("save_outputs in a multi-stage --unpaired-options list");
  }
  return true;
}

static void check_paired_options(const readpair_mapping_options_t *o, int n_options) {
  if (n_options != 1) unsupported("a multi-stage --paired-options list");
  const pairing_options &p = o->pairing;
  if (p.pair_mode != pair_mode || p.min_insert_size != min_insert_size || p.max_insert_size != max_insert_size ||
      p.strata != strata_flag || p.save_outputs != compute_mapping_qualities || p.pass1_num_outputs != num_tmp_outputs ||
      p.pass2_num_outputs != num_outputs || p.pass1_threshold != sw_vect_threshold ||
      p.pass2_threshold != sw_full_threshold || p.stop_count != (half_paired ? 1 : 0))
    unsupported("a custom --paired-options set");
  if (half_paired && (n_unpaired_mapping_options[0] != 1 || n_unpaired_mapping_options[1] != 1))
    unsupported("a custom half-paired --unpaired-options set");
}

// ------------------------------------------------------------------------------------------------
// One entry of re_buffer[] as the loop of gmapper.c:411-553 leaves it -- read off the entry when the loop has
// already been there (`final`), predicted from the strings fasta_get_next_read_with_range stored otherwise.
// ------------------------------------------------------------------------------------------------
struct Prep {
  bool ok;            // reaches handle_read / handle_readpair
  int len;            // re->read_len (colours in colour space)
  int initbp;
  const char *seq;    // first base / colour (after the initial base in colour space)
  const char *qual;   // re->qual after trimming, or NULL
  int qual_len;
  const uint32_t *words;   // final entries: the packed read as given (re->read[re->input_strand])
};

static int8_t g_ls_code[256], g_cs_code[256];
static std::once_flag g_code_once;
static void init_codes() {
  // fasta_open's translate tables, fasta.c:150-200
  memset(g_ls_code, -1, sizeof(g_ls_code));
  memset(g_cs_code, -1, sizeof(g_cs_code));
  const char *ls = "ACGTUMRWSYKVHDB";
  for (int i = 0; ls[i]; i++) {
    g_ls_code[(int)ls[i]] = (int8_t)i;
    g_ls_code[(int)ls[i] + 32] = (int8_t)i;
  }
  for (const char *c = "NnXx."; *c; c++) g_ls_code[(int)*c] = BASE_N;
  for (int i = 0; i < 4; i++) g_cs_code['0' + i] = (int8_t)i;
  for (const char *c = "4NnXx."; *c; c++) g_cs_code[(int)*c] = BASE_N;
}

// mate: 0 unpaired / first of a pair, 1 second of a pair (which of --trim-first / --trim-second applies)
static Prep prep_predict(const read_entry *e, bool paired, int mate) {
  Prep p;
  memset(&p, 0, sizeof(p));
  const bool cs = shrimp_mode == MODE_COLOUR_SPACE;
  const char *seq = e->seq, *qual = Qflag ? e->qual : NULL;
  if (!seq || (Qflag && !qual)) return p;
  int L = (int)strlen(seq);
  int QL = qual ? (int)strlen(qual) : 0;
  if (trim && (!paired || (mate == 0 ? trim_first : trim_second))) {  // trim_read, gmapper.c:262-281
    int keep = L - trim_end - trim_front;
    if (keep < 0) keep = 0;
    seq += std::min(trim_front, L);
    if (qual) {
      qual += std::min(trim_front, QL);
      QL = std::min(keep, std::max(0, QL - trim_front));
    }
    L = keep;
  }
  if (!cs && trim_illumina && qual) {  // gmapper.c:440-453
    int trailing = 0;
    for (int j = 0; j < QL; j++) trailing = qual[j] == 'B' ? trailing + 1 : 0;
    if (trailing > 0) {
      L = std::max(0, L - trailing);
      QL = std::max(0, QL - trailing);
    }
  }
  int read_len = L;
  int avg_qv = 0;
  if (Qflag && !ignore_qvs && min_avg_qv >= 0)
    for (int j = 0; j < QL; j++) avg_qv += qual[j] - qual_delta;
  if (cs) {
    if (L < 1 || g_ls_code[(unsigned char)seq[0]] < 0 || g_ls_code[(unsigned char)seq[0]] > BASE_T) return p;
    p.initbp = g_ls_code[(unsigned char)seq[0]];
    read_len--;
    seq++;
  }
  if (read_len > 0) avg_qv /= read_len;
  if (read_len > longest_read_len || (Qflag && !ignore_qvs && avg_qv < min_avg_qv)) return p;  // gmapper.c:496-527
  if (read_len <= 0) return p;
  const int8_t *code = cs ? g_cs_code : g_ls_code;
  for (int j = 0; j < read_len; j++)
    if (code[(unsigned char)seq[j]] < 0) return p;   // the loop will stop the run when it gets there
  p.ok = true;
  p.len = read_len;
  p.seq = seq;
  p.qual = qual;
  p.qual_len = QL;
  return p;
}

static Prep prep_final(const read_entry *e) {
  Prep p;
  memset(&p, 0, sizeof(p));
  p.ok = true;
  p.len = e->read_len;
  p.initbp = e->initbp[e->input_strand];           // read_reverse swapped them (gmapper.c:174-187)
  p.words = e->read[e->input_strand];
  p.qual = Qflag ? e->qual : NULL;
  p.qual_len = p.qual ? (int)strlen(p.qual) : 0;
  return p;
}

// ------------------------------------------------------------------------------------------------
// The look-ahead batch of one thread
// ------------------------------------------------------------------------------------------------
struct Batch {
  read_entry *base = nullptr, *last = nullptr;
  int n_entries = 0;
  bool paired = false;
  std::vector<int32_t> row_of;   // entry -> row of the device batch, -1 = not mapped
  int n_rows = 0, stride = 0, xstride = 0, qstride = 0;
  std::vector<uint32_t> reads;
  std::vector<int32_t> rlen, xover;
  std::vector<int8_t> initbp;
  std::vector<uint8_t> quals;
  // results
  std::vector<shrimp_hit> hits;
  std::vector<uint8_t> edits;
  std::vector<int32_t> n_unp;        // unpaired records per row
  std::vector<int64_t> first_unp;
  std::vector<shrimp_pair> pairs;
  std::vector<int32_t> n_pairs;      // per pair
  std::vector<int64_t> first_pair;
  // the later option sets of a multi-stage --unpaired-options list (mapping.c:1790-1841): the rows whose earlier
  // sets did not meet their stop condition, mapped again with the next set's parameters
  struct Stage {
    std::vector<int32_t> sub_of_row;   // row of the batch -> row of this stage's call, -1 = not mapped in this stage
    std::vector<shrimp_hit> hits;
    std::vector<uint8_t> edits;
    std::vector<int32_t> n_unp;
    std::vector<int64_t> first_unp;
  };
  std::vector<Stage> later;
};
static thread_local Batch t_batch;
static thread_local long long t_drop_snapshot, t_own_drops;

static void fill_row(Batch &B, int row, const Prep &p) {
  const bool cs = shrimp_mode == MODE_COLOUR_SPACE;
  uint32_t *w = &B.reads[(size_t)row * B.stride];
  if (p.words) {
    memcpy(w, p.words, (size_t)BPTO32BW(p.len) * 4);
  } else {
    const int8_t *code = cs ? g_cs_code : g_ls_code;
    for (int j = 0; j < p.len; j++) w[j >> 3] |= (uint32_t)(code[(unsigned char)p.seq[j]] & 15) << (4 * (j & 7));
  }
  B.rlen[row] = p.len;
  B.initbp[row] = (int8_t)p.initbp;
  if (B.xstride) {   // read_entry::crossover_score, gmapper.c:532-543
    int32_t *x = &B.xover[(size_t)row * B.xstride];
    for (int j = 0; j < p.len; j++) {
      int v = (int)(score_alpha * log(pr_err_from_qv(p.qual[j] - qual_delta) / 3.0) / log(2.0));
      if (v > -1)
        v = -1;
      else if (v < 2 * crossover_score)
        v = 2 * crossover_score;
      x[j] = v;
    }
  }
  if (B.qstride) memcpy(&B.quals[(size_t)row * B.qstride], p.qual, (size_t)std::min(p.qual_len, B.qstride - 1));
}

// the option list of the handle_read call that triggered this batch (every call of a run passes the same list)
static thread_local const read_mapping_options_t *t_stages;   // NULL: the default set, parameters from the globals
static thread_local int t_n_stages;

// read_pass2's stop condition (mapping.c:1737-1749): `stop_count` alignments at or above the stop threshold
static bool stage_done(const read_mapping_options_t &o, const shrimp_hit *h, int n) {
  if (o.pass2.stop_count == 0) return true;
  int cnt = 0;
  for (int i = 0; i < n; i++)
    if (h[i].score_full >= (int)abs_or_pct(o.pass2.stop_threshold, h[i].score_max)) cnt++;
  return cnt >= o.pass2.stop_count;
}

static void run_later_stages(Batch &B, int max_len) {
  const bool cs = shrimp_mode == MODE_COLOUR_SPACE;
  shrimp_gpu_ctx *ctx = thread_ctx();
  const int n = B.n_rows;
  std::vector<char> go((size_t)n, 0);   // rows that move on to the next set
  for (int r = 0; r < n; r++) go[r] = !stage_done(t_stages[0], B.hits.data() + B.first_unp[r], B.n_unp[r]);
  for (int k = 1; k < t_n_stages; k++) {
    B.later.emplace_back();
    Batch::Stage &S = B.later.back();
    S.sub_of_row.assign((size_t)n, -1);
    int m = 0;
    for (int r = 0; r < n; r++)
      if (go[r]) S.sub_of_row[r] = m++;
    if (m == 0) break;
    std::vector<uint32_t> reads((size_t)m * B.stride);
    std::vector<int32_t> rlen((size_t)m), xover((size_t)m * B.xstride);
    std::vector<int8_t> initbp((size_t)m);
    std::vector<uint8_t> quals((size_t)m * B.qstride);
    for (int r = 0; r < n; r++) {
      const int sub = S.sub_of_row[r];
      if (sub < 0) continue;
      memcpy(&reads[(size_t)sub * B.stride], &B.reads[(size_t)r * B.stride], (size_t)B.stride * 4);
      rlen[sub] = B.rlen[r];
      initbp[sub] = B.initbp[r];
      if (B.xstride) memcpy(&xover[(size_t)sub * B.xstride], &B.xover[(size_t)r * B.xstride], (size_t)B.xstride * 4);
      if (B.qstride) memcpy(&quals[(size_t)sub * B.qstride], &B.quals[(size_t)r * B.qstride], (size_t)B.qstride);
    }
    shrimp_map_params mp = stage_params(t_stages[k], k);
    if (B.xstride) {
      mp.crossover_scores = xover.data();
      mp.crossover_stride = B.xstride;
      mp.read_quals = quals.data();
      mp.qual_stride = B.qstride;
    }
    shrimp_map_stats st;
    memset(&st, 0, sizeof(st));
    int64_t n_hits = 0, e_used = 0;
    S.n_unp.assign((size_t)m, 0);
    S.hits.resize((size_t)m * std::max(1, mp.num_outputs));
    size_t e_cap = std::max<size_t>(4096, (size_t)m * 2 * (size_t)max_len);
    for (;;) {
      S.edits.resize(e_cap);
      const int rc = shrimp_gpu_map_reads(ctx, &mp, m, reads.data(), B.stride, rlen.data(), cs ? initbp.data() : nullptr,
                                          S.hits.data(), (int64_t)S.hits.size(), S.n_unp.data(), S.edits.data(),
                                          (int64_t)e_cap, &n_hits, &e_used, nullptr, 0, nullptr, &st);
      if (rc == SHRIMP_E_NOMEM && (size_t)e_used > e_cap) {
        e_cap = (size_t)e_used + 4096;
        continue;
      }
      if (rc != SHRIMP_OK) die("shrimp_gpu_map_reads (a later option set)");
      break;
    }
    tstats.add(st);
    S.first_unp.assign((size_t)m + 1, 0);
    for (int i = 0; i < m; i++) S.first_unp[i + 1] = S.first_unp[i] + S.n_unp[i];
    for (int r = 0; r < n; r++) {
      const int sub = S.sub_of_row[r];
      go[r] = sub >= 0 && !stage_done(t_stages[k], S.hits.data() + S.first_unp[sub], S.n_unp[sub]);
    }
  }
}

static void run_batch(Batch &B, const std::vector<Prep> &preps, const shrimp_map_params *stage0 = nullptr) {
  const bool cs = shrimp_mode == MODE_COLOUR_SPACE;
  shrimp_gpu_ctx *ctx = thread_ctx();
  const double t0 = omp_get_wtime();
  int max_len = 1;
  B.n_rows = 0;
  B.row_of.assign(B.n_entries, -1);
  for (int i = 0; i < B.n_entries; i++)
    if (preps[i].ok) {
      B.row_of[i] = B.n_rows++;
      max_len = std::max(max_len, preps[i].len);
    }
  const bool use_q = cs && Qflag && !ignore_qvs;       // per-position crossover scores + post_sw on read qualities
  B.stride = BPTO32BW(max_len) + 1;
  B.xstride = use_q ? max_len : 0;
  B.qstride = use_q ? max_len + 1 : 0;
  B.reads.assign((size_t)B.n_rows * B.stride, 0);
  B.rlen.assign(B.n_rows, 0);
  B.initbp.assign(B.n_rows, 0);
  B.xover.assign((size_t)B.n_rows * B.xstride, 0);
  B.quals.assign((size_t)B.n_rows * B.qstride, 0);
  for (int i = 0; i < B.n_entries; i++)
    if (preps[i].ok) {
      if (use_q && preps[i].qual_len < preps[i].len) {
        fprintf(stderr, "gmapper-b200: read [%s]: quality string shorter than the read\n", B.base[i].name);
        exit(1);
      }
      fill_row(B, B.row_of[i], preps[i]);
    }
  shrimp_map_params mp = stage0 ? *stage0 : (!B.paired && t_stages) ? stage_params(t_stages[0], 0) : map_params_from_globals();
  B.later.clear();
  if (use_q) {
    mp.crossover_scores = B.xover.data();
    mp.crossover_stride = B.xstride;
    mp.read_quals = B.quals.data();
    mp.qual_stride = B.qstride;
  }
  shrimp_map_stats st;
  memset(&st, 0, sizeof(st));
  int64_t n_hits = 0, e_used = 0;
  const int n = B.n_rows;
  const double t1 = omp_get_wtime();
  tstats.t_prep += t1 - t0;
  tstats.reads += (uint64_t)n;
  B.first_unp.assign((size_t)n + 1, 0);
  B.n_unp.assign((size_t)std::max(n, 1), 0);
  size_t e_cap = std::max<size_t>(B.edits.size(), std::max<size_t>(4096, (size_t)n * 2 * (size_t)max_len));
  if (!B.paired) {
    B.hits.resize((size_t)std::max(n, 1) * std::max(1, mp.num_outputs));
    for (;;) {
      B.edits.resize(e_cap);
      const int rc = shrimp_gpu_map_reads(ctx, &mp, n, B.reads.data(), B.stride, B.rlen.data(),
                                          cs ? B.initbp.data() : nullptr, B.hits.data(), (int64_t)B.hits.size(),
                                          B.n_unp.data(), B.edits.data(), (int64_t)e_cap, &n_hits, &e_used, nullptr, 0,
                                          nullptr, &st);
      if (rc == SHRIMP_E_NOMEM && (size_t)e_used > e_cap) {
        e_cap = (size_t)e_used + 4096;
        continue;
      }
      if (rc != SHRIMP_OK) die("shrimp_gpu_map_reads");
      break;
    }
    B.pairs.clear();
    for (int r = 0; r < n; r++) B.first_unp[r + 1] = B.first_unp[r] + B.n_unp[r];
    if (t_stages && t_n_stages > 1) run_later_stages(B, max_len);
  } else {
    const int np = n / 2;
    shrimp_pair_params pp = {pair_mode, min_insert_size, max_insert_size, half_paired ? 1 : 0};
    B.hits.resize((size_t)std::max(np, 1) * num_outputs * 4);
    B.pairs.resize((size_t)std::max(np, 1) * num_outputs);
    B.n_pairs.assign((size_t)std::max(np, 1), 0);
    int64_t n_pairs_out = 0;
    for (;;) {
      B.edits.resize(e_cap);
      const int rc = shrimp_gpu_map_pairs(ctx, &mp, &pp, np, B.reads.data(), B.stride, B.rlen.data(),
                                          cs ? B.initbp.data() : nullptr, B.hits.data(), (int64_t)B.hits.size(), &n_hits,
                                          B.pairs.data(), (int64_t)B.pairs.size(), &n_pairs_out, B.n_pairs.data(),
                                          B.n_unp.data(), B.edits.data(), (int64_t)e_cap, &e_used, &st);
      if (rc == SHRIMP_E_NOMEM && (size_t)e_used > e_cap) {
        e_cap = (size_t)e_used + 4096;
        continue;
      }
      if (rc != SHRIMP_OK) die("shrimp_gpu_map_pairs");
      break;
    }
    B.first_pair.assign((size_t)np + 1, 0);
    for (int k = 0; k < np; k++) B.first_pair[k + 1] = B.first_pair[k] + B.n_pairs[k];
    B.first_unp[0] = 2 * n_pairs_out;   // the members of the pairs come first (shrimp_b200.h)
    for (int r = 0; r < n; r++) B.first_unp[r + 1] = B.first_unp[r] + B.n_unp[r];
  }
  tstats.add(st);
  const double dt = omp_get_wtime() - t1;
  tstats.t_device += dt;
  if (tstats.t_first_device == 0) tstats.t_first_device = dt;   // allocations of the context's buffers
}

// entries [re, re + ahead) of the chunk that are loaded; `step` entries per unit (2 in paired mode).  The bound on
// re_buffer[], whose ends the shim is not told, is shim_lookahead.h's.
static thread_local LookaheadBound<read_entry> t_bound;
static int lookahead_limit(const read_entry *re, int step) {
  const long long drops = (total_reads_dropped + total_pairs_dropped) - t_own_drops - t_drop_snapshot;
  long long limit = t_bound.limit(re, drops, step, (long long)chunk_size);
  if (const char *e = getenv("SHRIMP_B200_BATCH")) limit = std::min<long long>(limit, std::max(step, atoi(e)));
  if (limit < step) limit = step;
  int n = step;
  while (n + step <= limit) {
    bool loaded = true;
    for (int k = 0; k < step; k++) loaded = loaded && re[n + k].name != NULL && re[n + k].seq != NULL;
    if (!loaded) break;
    n += step;
  }
  return n;
}

static bool row_matches(const Batch &B, int row, const read_entry *e) {
  const Prep p = prep_final(e);
  if (B.rlen[row] != p.len) return false;
  if (shrimp_mode == MODE_COLOUR_SPACE && B.initbp[row] != p.initbp) return false;
  const uint32_t *w = &B.reads[(size_t)row * B.stride];
  const int full = p.len / 8, rest = p.len % 8;
  if (memcmp(w, p.words, (size_t)full * 4)) return false;
  if (rest && ((w[full] ^ p.words[full]) & ((1u << (4 * rest)) - 1))) return false;
  if (B.xstride) {
    if (!e->crossover_score || memcmp(&B.xover[(size_t)row * B.xstride], e->crossover_score, (size_t)p.len * 4))
      return false;
    if (p.qual_len < p.len || memcmp(&B.quals[(size_t)row * B.qstride], p.qual, (size_t)p.len)) return false;
  }
  return true;
}

// Makes sure t_batch holds the records of entry `re` (and of re + 1 in paired mode); returns its entry index.
static int ensure_batch(read_entry *re, bool paired) {
  std::call_once(g_code_once, init_codes);
  Batch &B = t_batch;
  const int step = paired ? 2 : 1;
  bool have = B.base && B.paired == paired && re > B.last && re >= B.base && re + step <= B.base + B.n_entries &&
              (!paired || ((re - B.base) & 1) == 0);
  if (have) {
    const int i = (int)(re - B.base);
    for (int k = 0; k < step && have; k++) have = B.row_of[i + k] >= 0 && row_matches(B, B.row_of[i + k], re + k);
    if (!have) {   // the prediction was off: map this entry on its own, from what the loop made of it
      B.base = re;
      B.n_entries = step;
      std::vector<Prep> preps;
      for (int k = 0; k < step; k++) preps.push_back(prep_final(re + k));
      run_batch(B, preps);
      tstats.mispredicted++;
    }
  } else {
    B.base = re;
    B.paired = paired;
    B.n_entries = lookahead_limit(re, step);
    std::vector<Prep> preps((size_t)B.n_entries);
    for (int k = 0; k < step; k++) preps[k] = prep_final(re + k);
    for (int i = step; i < B.n_entries; i += step) {
      bool ok = true;
      for (int k = 0; k < step; k++) {
        preps[i + k] = prep_predict(re + i + k, paired, k);
        ok = ok && preps[i + k].ok;
      }
      for (int k = 0; k < step; k++) preps[i + k].ok = ok;   // a pair is dropped as a whole (gmapper.c:517-525)
    }
    run_batch(B, preps);
    tstats.batches++;
  }
  B.last = re + step - 1;
  return (int)(re - B.base);
}

// ------------------------------------------------------------------------------------------------
// shrimp_hit -> struct read_hit + struct sw_full_results (what hit_run_full_sw / hit_run_post_sw leave behind,
// mapping.c:331-402, :1609-1625)
// ------------------------------------------------------------------------------------------------
static void build_hit(const std::vector<uint8_t> &edits, const shrimp_hit &h, read_entry *re, read_hit *rh) {
  const bool cs = shrimp_mode == MODE_COLOUR_SPACE;
  memset(rh, 0, sizeof(*rh));
  struct sw_full_results *s = (struct sw_full_results *)my_calloc(sizeof(*s), &mem_mapping, "sfrp [%s]", re->name);
  rh->sfrp = s;
  rh->cn = h.cn;
  rh->st = re->input_strand;      // hit_run_full_sw re-orients every hit to the input strand (mapping.c:349-351)
  rh->gen_st = h.gen_st;
  rh->g_off = h.g_off;
  rh->w_len = h.w_len;
  rh->score_window_gen = h.score_window_gen;
  rh->score_vector = h.score_vector;
  rh->score_full = h.score_full;
  rh->score_max = h.score_max;
  rh->pct_score_full = (1000 * 100 * h.score_full) / h.score_max;
  rh->pass2_key = h.pass2_key;
  rh->matches = h.matches;
  rh->pair_min = rh->pair_max = -1;
  rh->saved = 1;
  s->in_use = true;
  s->mqv = 255;
  s->read_start = h.read_start;
  s->rmapped = h.rmapped;
  s->genome_start = h.genome_start;
  s->gmapped = h.gmapped;
  s->matches = h.sfr_matches;
  s->mismatches = h.mismatches;
  s->insertions = h.insertions;
  s->deletions = h.deletions;
  s->crossovers = h.crossovers;
  s->score = h.sw_score;
  s->posterior = h.posterior;
  if (compute_mapping_qualities) {
    s->posterior_score = h.score_full;
    s->pct_posterior_score = (1000 * 100 * h.score_full) / h.score_max;
  }
  // dbalign / qralign (shim_align.h), on the genome strand and the read strand the alignment ran on
  const uint8_t *ed = &edits[(size_t)h.edit_off];
  const int n = h.edit_len;
  char *db = (char *)xmalloc((size_t)n + 1), *qr = (char *)xmalloc((size_t)n + 1);
  const uint32_t *gen = h.gen_st == 0 ? genome_contigs[h.cn] : genome_contigs_rc[h.cn];
  edit_to_strings(ed, n, gen, h.genome_start, re->read[rh->st], h.read_start, cs, cs ? re->initbp[rh->st] : 0,
                  re->read_len, db, qr);
  s->dbalign = db;
  s->qralign = qr;
  if (cs && compute_mapping_qualities && h.score_full > 0) {
    // sfrp->qual (get_base_qualities, sw-post.c:591-608): rmapped bytes right after the edit script
    s->qual = (char *)xmalloc((size_t)h.rmapped + 1);
    memcpy(s->qual, ed + n, (size_t)h.rmapped);
    s->qual[h.rmapped] = 0;
  }
}

static void free_hits(read_entry *re, read_hit *rh, int n) {
  for (int i = 0; i < n; i++) free_sfrp(&rh[i].sfrp, re, &mem_mapping);
}

// the unpaired records of row `row` -> read_output now (save_outputs false) or re->final_unpaired_hits
// (read_save_final_hits, mapping.c:1754-1770)
static void emit_records(const shrimp_hit *h, int n, const std::vector<uint8_t> &edits, read_entry *re, bool save_outputs) {
  if (n <= 0) return;
  read_hit *rh = (read_hit *)my_malloc((size_t)n * sizeof(read_hit), &mem_mapping, "final_unpaired_hits [%s]", re->name);
  const double t0 = omp_get_wtime();
  for (int i = 0; i < n; i++) build_hit(edits, h[i], re, &rh[i]);
  tstats.t_build += omp_get_wtime() - t0;
  tstats.records += (uint64_t)n;
  re->final_matches += n;
#pragma omp atomic
  total_reads_matched++;
#pragma omp atomic
  total_single_matches += re->final_matches;
  if (save_outputs) {
    re->final_unpaired_hits = rh;
    re->n_final_unpaired_hits = n;
  } else {
    std::vector<read_hit *> ptr((size_t)n);
    for (int i = 0; i < n; i++) ptr[i] = &rh[i];
    const double t1 = omp_get_wtime();
    read_output(re, ptr.data(), n);
    tstats.t_output += omp_get_wtime() - t1;
    free_hits(re, rh, n);
    my_free(rh, (size_t)n * sizeof(read_hit), &mem_mapping, "final_unpaired_hits [%s]", re->name);
  }
  re->mapped = true;
}

static void emit_unpaired(const Batch &B, int row, read_entry *re, bool save_outputs) {
  const int n = B.n_unp[row];
  if (n > 0) emit_records(B.hits.data() + B.first_unp[row], n, B.edits, re, save_outputs);
  // the later option sets this read went through, in their order (every set that finds alignments prints them,
  // mapping.c:1824-1833)
  for (const Batch::Stage &S : B.later) {
    const int sub = S.sub_of_row[row];
    if (sub < 0) break;
    if (S.n_unp[sub] > 0) emit_records(S.hits.data() + S.first_unp[sub], S.n_unp[sub], S.edits, re, save_outputs);
  }
}

}  // namespace shrimp_shim

using namespace shrimp_shim;

// ================================================================================================
// The reference's symbols
// ================================================================================================

void handle_read(struct read_entry *re, struct read_mapping_options_t *options, int n_options) {   // mapping.c:1773
  const llint before = gettimeinusecs();
  if (pair_mode != PAIR_NONE) unsupported("handle_read outside handle_readpair in paired mode");
  static thread_local const read_mapping_options_t *checked;
  static thread_local bool checked_custom;
  if (checked != options) {   // once per thread: the list is the same for every read of a run
    checked_custom = custom_unpaired_options(options, n_options);
    checked = options;
  }
  t_stages = checked_custom ? options : nullptr;
  t_n_stages = n_options;
  const int i = ensure_batch(re, false);
  Batch &B = t_batch;
  emit_unpaired(B, B.row_of[i], re, options[0].pass2.save_outputs);
  tpg.read_handle_usecs += gettimeinusecs() - before;
  t_drop_snapshot = (total_reads_dropped + total_pairs_dropped) - t_own_drops;

  // mapping.c:1849-1867
  if (aligned_reads_file != NULL && re->mapped) {
#pragma omp critical(aligned_reads_file)
    { fasta_write_read(aligned_reads_file, re); }
  }
  if ((unaligned_reads_file != NULL || sam_unaligned) && !re->mapped) {
#pragma omp critical(unaligned_reads_file)
    {
      if (unaligned_reads_file != NULL) fasta_write_read(unaligned_reads_file, re);
    }
    if (sam_unaligned) hit_output(re, NULL, NULL, false, NULL, 0);
  }
}

void handle_readpair(pair_entry *pe, struct readpair_mapping_options_t *options, int n_options) {   // mapping.c:2504
  const llint before = gettimeinusecs();
  read_entry *re1 = pe->re[0], *re2 = pe->re[1];
  check_paired_options(options, n_options);
  if (re2 != re1 + 1) unsupported("a pair whose mates are not neighbours in the chunk buffer");
  const int i = ensure_batch(re1, true);
  Batch &B = t_batch;
  const int row = B.row_of[i], k = row / 2;
  const int np = B.n_pairs[k];
  const shrimp_pair *P = B.pairs.data() + B.first_pair[k];

  if (np > 0) {
    // readpair_save_final_hits (mapping.c:2446-2500): the distinct read_hits of the pairs go to a pool per mate, in
    // the order the pairs name them; every pooled hit lists the pairs it is part of
    read_hit_pair *fp = (read_hit_pair *)my_calloc((size_t)np * sizeof(read_hit_pair), &mem_mapping,
                                                   "final_paired_hits [%s,%s]", re1->name, re2->name);
    std::vector<int> slot_of[2];
    std::vector<read_hit> pool[2];
    std::vector<std::vector<int>> idx[2];
    for (int p = 0; p < np; p++) {
      fp[p].score = P[p].score;
      fp[p].score_max = P[p].score_max;
      fp[p].key = P[p].key;
      fp[p].pct_score = (1000 * 100 * P[p].score) / P[p].score_max;
      fp[p].insert_size = P[p].insert_size;
      fp[p].improper_mapping = false;
      for (int nip = 0; nip < 2; nip++) {
        const shrimp_hit &h = B.hits[(size_t)P[p].hit_idx[nip]];
        int q = -1;
        for (size_t j = 0; j < slot_of[nip].size(); j++)
          if (slot_of[nip][j] == h.hit_slot) q = (int)j;
        if (q < 0) {
          q = (int)slot_of[nip].size();
          slot_of[nip].push_back(h.hit_slot);
          pool[nip].emplace_back();
          build_hit(B.edits, h, pe->re[nip], &pool[nip].back());
          idx[nip].emplace_back();
        }
        idx[nip][q].push_back(p);
        fp[p].rh[nip] = NULL;
        fp[p].rh_idx[nip] = q;
      }
    }
    for (int nip = 0; nip < 2; nip++) {
      const int ps = (int)pool[nip].size();
      read_hit *pp = (read_hit *)my_malloc((size_t)ps * sizeof(read_hit), &mem_mapping,
                                           "final_paired_hit_pool[%d] [%s,%s]", nip, re1->name, re2->name);
      for (int q = 0; q < ps; q++) {
        pp[q] = pool[nip][q];
        const int m = (int)idx[nip][q].size();
        pp[q].paired_hit_idx = (int *)my_malloc((size_t)m * sizeof(int), &mem_mapping, "paired_hits [%s]",
                                                pe->re[nip]->name);
        for (int j = 0; j < m; j++) pp[q].paired_hit_idx[j] = idx[nip][q][j];
        pp[q].n_paired_hit_idx = m;
      }
      pe->final_paired_hit_pool[nip] = pp;
      pe->final_paired_hit_pool_size[nip] = ps;
    }
    pe->final_paired_hits = fp;
    pe->n_final_paired_hits = np;
    pe->mapped = true;
    re1->final_matches += np;
#pragma omp atomic
    total_pairs_matched++;
#pragma omp atomic
    total_paired_matches += re1->final_matches;
    if (!options[0].pairing.save_outputs) {
      // readpair_output_no_mqv takes pointers (mapping.c:2575-2581)
      for (int p = 0; p < np; p++)
        for (int nip = 0; nip < 2; nip++) fp[p].rh[nip] = &pe->final_paired_hit_pool[nip][fp[p].rh_idx[nip]];
      readpair_output_no_mqv(pe, fp, np);
      for (int nip = 0; nip < 2; nip++) {
        for (int q = 0; q < pe->final_paired_hit_pool_size[nip]; q++) {
          read_hit *rh = &pe->final_paired_hit_pool[nip][q];
          free_sfrp(&rh->sfrp, pe->re[nip], &mem_mapping);
          my_free(rh->paired_hit_idx, rh->n_paired_hit_idx * sizeof(int), &mem_mapping, "paired_hits [%s]",
                  pe->re[nip]->name);
        }
        my_free(pe->final_paired_hit_pool[nip], pe->final_paired_hit_pool_size[nip] * sizeof(read_hit), &mem_mapping,
                "final_paired_hit_pool[%d] [%s,%s]", nip, re1->name, re2->name);
        pe->final_paired_hit_pool[nip] = NULL;
        pe->final_paired_hit_pool_size[nip] = 0;
      }
      my_free(fp, (size_t)np * sizeof(read_hit_pair), &mem_mapping, "final_paired_hits [%s,%s]", re1->name, re2->name);
      pe->final_paired_hits = NULL;
      pe->n_final_paired_hits = 0;
    }
  }
  tpg.read_handle_usecs += gettimeinusecs() - before;

  if (half_paired) {   // mapping.c:2607-2611: every pair falls through (stop_threshold 101 %, gmapper.c:2686-2687)
    const llint b2 = gettimeinusecs();
    emit_unpaired(B, row, re1, unpaired_mapping_options[0][0].pass2.save_outputs);
    emit_unpaired(B, row + 1, re2, unpaired_mapping_options[1][0].pass2.save_outputs);
    tpg.read_handle_usecs += gettimeinusecs() - b2;
  }
  t_drop_snapshot = (total_reads_dropped + total_pairs_dropped) - t_own_drops;

  // OUTPUT, mapping.c:2613-2636
  const double t_out = omp_get_wtime();
  readpair_output(pe);
  tstats.t_output += omp_get_wtime() - t_out;
  if (aligned_reads_file != NULL && (pe->mapped || re1->mapped || re2->mapped)) {
#pragma omp critical(aligned_reads_file)
    {
      fasta_write_read(aligned_reads_file, re1);
      fasta_write_read(aligned_reads_file, re2);
    }
  }
  if ((unaligned_reads_file != NULL || sam_unaligned) && !(pe->mapped || re1->mapped || re2->mapped)) {
#pragma omp critical(unaligned_reads_file)
    {
      if (unaligned_reads_file != NULL) {
        fasta_write_read(unaligned_reads_file, re1);
        fasta_write_read(unaligned_reads_file, re2);
      }
    }
    if (sam_unaligned) {
      hit_output(re1, NULL, NULL, true, NULL, 0);
      hit_output(re2, NULL, NULL, false, NULL, 0);
    }
  }
}

// The genome span between the 5' ends of two mapped mates (mapping.c:405-456); readpair_output uses it for improper
// pairs (output.c:1233).
int get_insert_size(struct read_hit *rh, struct read_hit *rh_mp) {
  if (rh_mp == NULL || rh == NULL || rh->cn != rh_mp->cn) return 0;
  struct Ends {
    int start, end;
    bool rev;
  };
  auto ends = [](const read_hit *h) {
    Ends e;
    e.rev = h->gen_st == 1;
    const struct sw_full_results *s = h->sfrp;
    if (!e.rev)
      e.start = s->genome_start + 1;
    else
      e.start = ((int)genome_len[h->cn] - s->genome_start) - (s->rmapped - 1 - s->deletions + s->insertions);
    e.end = e.start + s->gmapped - 1;
    return e;
  };
  const Ends a = ends(rh), b = ends(rh_mp);
  const int fivep = a.rev ? a.end : a.start - 1;
  const int fivep_mp = b.rev ? b.end : b.start - 1;
  return fivep_mp - fivep;
}
