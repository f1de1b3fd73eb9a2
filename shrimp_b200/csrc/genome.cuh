// Device-resident genome + seed index (owned by shrimp_gpu_ctx::genome).
#pragma once
#include "common.cuh"

namespace shrimp {

#define SHRIMP_MAX_SEEDS 16
#define SHRIMP_MAX_RUNS 8

// Seed table handed to kernels by value (the reference keeps `seed[]` as a global, gmapper.h:158).
struct SeedTable {
  int n_seeds;
  int max_span, min_span;
  int hflag;
  unsigned long long mask[SHRIMP_MAX_SEEDS];
  int span[SHRIMP_MAX_SEEDS];
  int weight[SHRIMP_MAX_SEEDS];
  // fast projection (scan.cu): the seed as runs of consecutive care positions, in k-mer order.  Run q takes
  // run_len bases starting at base run_src of the k-mer and drops them at base run_dst of the bucket id.
  // n_runs = 0: no fast form (span > 32, more than SHRIMP_MAX_RUNS runs, or -H) -> generic kmer_to_mapidx.
  unsigned char n_runs[SHRIMP_MAX_SEEDS];
  unsigned char run_src[SHRIMP_MAX_SEEDS][SHRIMP_MAX_RUNS];
  unsigned char run_len[SHRIMP_MAX_SEEDS][SHRIMP_MAX_RUNS];
  unsigned char run_dst[SHRIMP_MAX_SEEDS][SHRIMP_MAX_RUNS];
};

// HBM layout: every orientation of the genome is ONE packed 4-bit array in global coordinates
// (contig cn occupies nibbles [contig_off[cn], contig_off[cn] + genome_len[cn])), replacing the
// reference's per-contig heap arrays (gmapper.h:264-272).  The projection is CSR per seed:
// offs[sn][m] .. offs[sn][m+1] index pos[sn] (ascending global start positions of bucket m),
// replacing uint32_t ***genomemap / **genomemap_len (gmapper.h:262-263).
struct DeviceGenome {
  int num_contigs = 0;
  int colour_space = 0;
  uint64_t total_len = 0;
  size_t words = 0;                 // packed words per orientation (incl. padding)
  std::vector<uint32_t> h_off;      // num_contigs + 1
  std::vector<uint32_t> h_len;
  DevBuf d_off, d_len;              // uint32 [num_contigs+1], [num_contigs]
  DevBuf d_ls, d_ls_rc, d_cs, d_cs_rc;
  // index
  SeedTable seeds{};
  bool have_index = false;
  uint32_t nbuckets[SHRIMP_MAX_SEEDS] = {0};
  uint64_t total[SHRIMP_MAX_SEEDS] = {0};
  DevBuf d_offs[SHRIMP_MAX_SEEDS];  // uint32 [nbuckets+1]
  DevBuf d_pos[SHRIMP_MAX_SEEDS];   // uint32 [total], then (sparse projections) the bucket heads, see build_bucket_heads
  uint32_t head_off[SHRIMP_MAX_SEEDS] = {0};   // first word of the bucket heads in d_pos[sn], 0 = none
  std::vector<std::string> contig_names;   // set by shrimp_gpu_projection_load
};

struct IndexView {
  const uint32_t *offs[SHRIMP_MAX_SEEDS];
  const uint32_t *pos[SHRIMP_MAX_SEEDS];
  uint32_t head_off[SHRIMP_MAX_SEEDS];   // bucket heads behind the position lists (0 = none)
};

// Bucket heads of a SPARSE projection (a few entries per bucket: C1, C2, C4, C5).  A k-mer lookup through the CSR costs
// two dependent DRAM accesses of 64 bytes each for 8 bytes of bucket bounds and ~24 bytes of list (profiles/
// r02_summary.md section 2: 7 x the algorithmic bytes).  The head of bucket m is one 64-byte record behind the
// position lists, in the same allocation so that a list is addressed by one 32-bit word offset wherever it lies:
//   word 0 = list length; length <= 15: words 1..length = the list itself; longer: word 1 = its CSR start.
// One DRAM access per k-mer; the list entries are then read from the line that is already on chip.
#define SHRIMP_HEAD_WORDS 16
int build_bucket_heads(shrimp_gpu_ctx *ctx, DeviceGenome *g);

struct GenomeView {
  const uint32_t *ls, *ls_rc, *cs, *cs_rc;
  const uint32_t *contig_off;  // [num_contigs + 1]
  const uint32_t *contig_len;
  int num_contigs;
};

inline DeviceGenome *genome_of(shrimp_gpu_ctx *ctx) { return (DeviceGenome *)ctx->genome; }
int seed_table_init(SeedTable &S, int n_seeds, const uint64_t *masks, const int32_t *spans, const int32_t *weights,
                    int hflag, const char *who);

// KMER_TO_MAPIDX (gmapper.h:370) for the k-mer of seed sn that starts at base `start` of the
// packed sequence `seq`: kmer_to_mapidx_orig (gmapper.h:349-368) concatenates the low two bits
// of the bases under the mask, mask bit 0 (= last base of the k-mer) first and therefore most
// significant; kmer_to_mapidx_hash (gmapper.h:323-336, -H) hashes the masked 4-bit window words.
__device__ __forceinline__ uint32_t hash32(uint32_t a) {  // gmapper.h:309-319
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) ^ (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a + 0xd3a2646cu) ^ (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) ^ (a >> 16);
  return a;
}

__device__ __forceinline__ uint32_t kmer_to_mapidx(const SeedTable &S, int sn, const uint32_t *seq, uint64_t start) {
  const int span = S.span[sn];
  const unsigned long long mask = S.mask[sn];
  if (!S.hflag) {
    uint32_t m = 0;
    for (int i = 0; i < span; i++) {
      if ((mask >> i) & 1ull) m = (m << 2) | (extract4(seq, start + (uint64_t)(span - 1 - i)) & 3u);
    }
    return m;
  }
  uint32_t m = 0;
  const int nw = (S.max_span + 7) / 8;
  for (int w = 0; w < nw; w++) {
    uint32_t word = 0;
    for (int n = 0; n < 8; n++) {
      int i = 8 * w + n;
      if (i < span && ((mask >> i) & 1ull)) word |= extract4(seq, start + (uint64_t)(span - 1 - i)) << (4 * n);
    }
    m = hash32(word ^ m);
  }
  return m & ((1u << 24) - 1u);
}

}  // namespace shrimp
