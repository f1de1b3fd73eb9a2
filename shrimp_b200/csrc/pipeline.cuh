// Chunk pipeline state shared by scan.cu / pass1.cu / sw_full.cu / pipeline.cu.
#pragma once
#include "genome.cuh"

namespace shrimp {

// One candidate window of a read strand: a read_hit (gmapper-definitions.h:125-153) reduced to
// the fields the device stages produce/consume.  Coordinates are positive-strand (g_off_pos_strand).
struct DevHit {
  uint32_t g_off;       // window start inside contig cn
  int32_t cn;
  int32_t w_len;
  int32_t matches;      // summed anchor weights
  int32_t wg;           // score_window_gen
  int32_t score_max;    // min(read_len, w_len) * match
  int32_t ax, ay, alen, awidth;   // joined anchor, relative to the window
  int32_t score_vector; // true sw_vector score (-1 = not computed), replaced by the replayed value in pass 1
  int32_t pct_vector;
};

// Mapping parameters in device form (doubles pre-divided on the host exactly as abs_or_pct does).
struct MapParamsDev {
  int colour_space;
  int match_mode, gapless, hash_filter_calls, use_region_counts;
  int region_bits, region_overlap;
  uint32_t list_cutoff;
  int num_tmp_outputs;
  int min_matches;
  int match, b_gap_open, b_gap_ext;     // CLI sign
  double window_len, window_len_frac;    // window_len < 0: absolute
  double wgen_thr, wgen_frac;            // window_gen_threshold and threshold/100.0
  double vect_thr, vect_frac;
  double full_thr, full_frac;
  double overlap_thr, overlap_frac;      // window_overlap
  int Gflag, Tflag;
  int anchor_width;
  // pair modes that reverse a mate before mapping (read_reverse gmapper.c:174-187, pair_reverse
  // gmapper-defaults.h:184-191): [0] even reads (first mates), [1] odd reads.  A reversed read is mapped with its
  // strands swapped (row 2r = the reverse complement, row 2r + 1 = the read as given) and its input strand is 1.
  int rev_mate[2];
};

// abs_or_pct (util.h:48-53) with the division already done on the host: x<0 ? -x : base*(x/100.0)
__host__ __device__ __forceinline__ double abs_or_pct_d(double thr, double frac, double base) {
  return thr < 0 ? -thr : base * frac;
}

}  // namespace shrimp
