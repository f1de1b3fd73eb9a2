// post_sw on the device: colour-space mapping qualities.
//
// Replaces common/sw-post.c post_sw :640-758 (load_local_vectors :472-545, forward_backward :270-372,
// post_traceback :182-207, fix_base_calls :548-581, get_base_qualities :584-601, get_posterior :604-626) as
// hit_run_post_sw calls it (mapping.c:1609-1625) for every alignment with a positive full-SW score.
//
// The model is a 16-node chain (node = letter pair, left/right) over the aligned read columns.  One HALF-WARP per
// alignment, lane = node: the forward and backward recurrences take their four predecessor / successor values by
// width-16 shuffles, sums run in the reference's index order, the scale (minimum over the nodes) is a half-warp
// reduction.  The forward and the backward recurrence advance in the same loop (two independent chains, one shared
// logarithm per lane); forwards[][] and backwards[][] of all columns wait for the posterior pass in a global scratch
// row per resident half-warp (written and read back by the same lane, 128 contiguous bytes per column: it lives in
// L2), so that shared memory holds only the columns and the scales and the register file bounds the occupancy.  The kernel is persistent: a
// CTA takes groups of alignments grid-stride.  The two alignments of a warp run in lockstep (trip counts = the
// longer one).  Every double is the reference's, bit for bit: the -log() emission
// terms come from the host (libm) as constants / a 256-entry table over the quality characters, and exp() / log()
// inside the recurrences are glibc_math.cuh -- the libm algorithms operation by operation -- because equally likely
// alternatives (two colour errors of the same quality) tie exactly and the last bit of a sum decides the base call.
// Compiled with -fmad=false: only the fused multiply-adds the transcription names are fused.
#include "stages.cuh"
#include "glibc_math.cuh"
#include <vector>

namespace shrimp {

#define PS_CTAS_PER_SM 7
#define PS_LEFT(i) (((i) >> 2) & 3)
#define PS_RIGHT(i) ((i) & 3)

__device__ __forceinline__ double half_min(double v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    // the values are never NaN: a compare-and-select (the reference's own `<` scan, sw-post.c:343-347) instead of
    // fmin's NaN handling.  Which zero a tie of +0 and -0 yields does not matter: the scale is only subtracted and
    // summed, and x - (+-0) = x, exp(+-0) = 1
    const double w = __shfl_xor_sync(0xffffffffu, v, o, 16);
    v = w < v ? w : v;
  }
  return v;
}
__device__ __forceinline__ int warp_max_int(int v) {
  // redux.sync leaves its result in a uniform register: the loops bounded by it are known to be warp-uniform and
  // the shuffles inside need no reconvergence code
  return __reduce_max_sync(0xffffffffu, v);
}
__device__ __forceinline__ int ps_cstols(int first_letter, int colour) {   // cstols, util.h:157-180
  if (first_letter == 15 || colour < 0 || colour > 3) return 15;
  return (first_letter % 2 == 0) ? (4 + first_letter + colour) % 4 : (4 + first_letter - colour) % 4;
}
__device__ __forceinline__ int ps_qv_from_pr_err(double pr_err, double log10v, const glibc_math::Tables &GT) {   // util.h:268-276
  if (pr_err > .99999999) return 0;
  else if (pr_err < 1E-25) return 250;
  else return (int)(-10.0 * glibc_math::log_glibc(pr_err, GT) / log10v);
}

struct PsCol {   // one aligned read column (struct column, sw-post.c:61-78, without the recurrences' arrays)
  int8_t let;    // genome letter 0-3, -1 = other (N), -2 = no letter emission (read base against a gap)
  int8_t col;    // colour emitted
  uint8_t q;     // quality character of the colour (table index)
  uint8_t kind;  // 0: constant crossover rate (FASTA), 1: quality table, 2: colour N / run of N: rate .75
  int8_t call;   // the full SW's base call (layer letter), 15 = N
  int8_t maxp;   // max_posterior
  uint8_t qual;  // 33 + base quality
  uint8_t pad;
};

__host__ __device__ inline size_t ps_smem_doubles_per_half(int max_cols) {
  // forwscale[max_cols], backscale[max_cols], columns[max_cols], the segmented XOR scan of the read (one byte per position)
  return 2 * (size_t)max_cols + ((size_t)max_cols * sizeof(PsCol) + 7) / 8 + ((size_t)max_cols + 7) / 8;
}

__global__ void __launch_bounds__(128, PS_CTAS_PER_SM) post_sw_kernel(const PostParams P, int halves_per_cta, int max_cols,
                                                                      int n_groups, double *fw_scratch) {
  extern __shared__ double ps_smem[];
  const int lane = threadIdx.x & 31, hl = lane & 15;
  const int hidx = threadIdx.x >> 4;   // half-warp of the CTA
  double *fscale = ps_smem + (size_t)hidx * ps_smem_doubles_per_half(max_cols);
  double *bsc = fscale + max_cols;
  PsCol *cols = (PsCol *)(bsc + max_cols);
  uint8_t *pxs = (uint8_t *)(cols + max_cols);   // per read position
  // forwards[max_cols][16] and backwards[max_cols][16] of this half-warp
  double *fw = fw_scratch + ((size_t)blockIdx.x * halves_per_cta + hidx) * (size_t)max_cols * 32;
  double *bw = fw + (size_t)max_cols * 16;
  const glibc_math::Tables GT = {P.gm_tab, P.gm_tab + 8, P.gm_tab + 8 + 256, P.gm_tab + 8 + 256 + 18};
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
  int slot = grp * halves_per_cta + hidx;
  const bool live = hidx < halves_per_cta && slot < P.n_tasks;
  if (!live) slot = P.n_tasks - 1;
  const FullTask T = P.tasks[slot];
  FullResult R = P.results[slot];
  const bool run = live && T.run && R.score > 0;   // mapping.c:1648: score_full > 0
  const uint32_t *genome = T.gen_st ? P.genome_rc : P.genome_fwd;
  const uint32_t *read = P.reads + (size_t)T.ridx * P.stride;
  const uint8_t *ops = P.ops + (size_t)slot * (size_t)P.ops_stride;
  const uint8_t *rq = P.read_quals ? P.read_quals + (size_t)(T.ridx >> 1) * (size_t)P.qual_stride + P.qual_vector_offset
                                   : nullptr;
  const int init_bp = T.initbp;
#define PS_EXP(x) glibc_math::exp_glibc((x), GT)
#define PS_LOG(x) glibc_math::log_glibc((x), GT)
  // ---- load_local_vectors (sw-post.c:472-545) on all 16 lanes ------------------------------------------------
  // The full SW's base call at read position j on layer kk is the layer's start letter (kk + initbp) % 4 XORed with
  // the colours since the last N (cstols, util.h:157-180, is XOR on the 2-bit codes; a colour N gives letter N and
  // restarts the layer, sw-full-cs.c:1181-1196): a segmented XOR scan over the read.  Column index, read position and
  // genome position of every edit operation are prefix sums over the script.
  const int rlen_t = run ? T.rlen : 0, n_ops = run ? R.ops_len : 0;
  const int w_rlen = warp_max_int(rlen_t), w_ops = warp_max_int(n_ops);
  const int hshift = (lane >> 4) << 4;
  int start_run = 0, min_qv = 10000;
  {
    int carry = 0, any_n = 0;
    for (int p0 = 0; p0 < w_rlen; p0 += 16) {
      const int p = p0 + hl;
      const bool in = p < rlen_t;
      const int c = in ? (int)extract4(read, (uint64_t)p) : 15;
      int v = c != 15 ? c : 0, f = c == 15 ? 1 : 0;
#pragma unroll
      for (int d = 1; d < 16; d <<= 1) {
        const int v2 = __shfl_up_sync(0xffffffffu, v, d, 16), f2 = __shfl_up_sync(0xffffffffu, f, d, 16);
        if (hl >= d) {
          if (!f) v ^= v2;
          f |= f2;
        }
      }
      if (!f) v ^= carry;
      if (in) pxs[p] = (uint8_t)v;
      carry = __shfl_sync(0xffffffffu, v, 15, 16);
      // the colours before the alignment (:484-496): their XOR, or N as soon as one of them is N
      const bool pre = in && p < R.read_start;
      any_n |= (int)((__ballot_sync(0xffffffffu, pre && c == 15) >> hshift) & 0xffffu);
      int x = (pre && c != 15) ? c : 0, mq = (pre && rq) ? (int)rq[p] : 10000;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        x ^= __shfl_xor_sync(0xffffffffu, x, o, 16);
        mq = min(mq, __shfl_xor_sync(0xffffffffu, mq, o, 16));
      }
      start_run ^= x;
      min_qv = min(min_qv, mq);
    }
    if (any_n) {
      start_run = 15;
      min_qv = 0;
    }
  }
  __syncwarp();
  int len = 0;
  {
    int colbase = 0, genbase = 0;
    const uint64_t g0 = (uint64_t)T.goff_global + (uint64_t)(R.genome_start - (int)T.goff_contig);
    for (int q0 = 0; q0 < w_ops; q0 += 16) {
      const int q = q0 + hl;
      const bool in = q < n_ops;
      const int op = in ? (int)ops[R.ops_start + q] : 0, type = op & 3, kk = (op >> 4) & 3;
      const int isread = (in && type != 1) ? 1 : 0, isgen = (in && type != 2) ? 1 : 0;
      int ir = isread, ig = isgen;
#pragma unroll
      for (int d = 1; d < 16; d <<= 1) {
        const int r2 = __shfl_up_sync(0xffffffffu, ir, d, 16), g2 = __shfl_up_sync(0xffffffffu, ig, d, 16);
        if (hl >= d) {
          ir += r2;
          ig += g2;
        }
      }
      const int col = colbase + ir - isread;
      if (isread && col < max_cols) {
        const int j = R.read_start + col;
        const int c = (int)extract4(read, (uint64_t)j);
        PsCol pc;
        int gl = 15;   // genome letter of a match column
        if (type == 3) {
          const int g = (int)extract4(genome, g0 + (uint64_t)(genbase + ig - isgen));
          pc.let = (int8_t)(g <= 3 ? g : -1);
          gl = g <= 3 ? g : 15;
        } else {
          pc.let = -2;
        }
        if ((col == 0 && start_run == 15) || c == 15) {
          pc.col = 0;
          pc.kind = 2;
          pc.q = 0;
        } else {
          pc.col = (int8_t)(c ^ (col == 0 ? start_run : 0));
          if (rq) {
            pc.kind = 1;
            pc.q = (uint8_t)(col == 0 ? min(min_qv, (int)rq[j]) : (int)rq[j]);
          } else {
            pc.kind = 0;
            pc.q = 0;
          }
        }
        // base_call = the letter qralign shows (sw-post.c:519): the layer's letter -- or, for an unknown colour, the
        // genome's letter that pretty_print puts in its place on a match column (sw-full-cs.c:1049-1056)
        pc.call = (int8_t)(c == 15 ? gl : (((kk + init_bp) & 3) ^ (int)pxs[j]));
        pc.maxp = 0;
        pc.qual = 33;
        pc.pad = 0;
        cols[col] = pc;
      }
      colbase += __shfl_sync(0xffffffffu, ir, 15, 16);
      genbase += __shfl_sync(0xffffffffu, ig, 15, 16);
    }
    len = colbase;
  }
  if (len > max_cols) len = 0;   // cannot happen: columns <= read length
  if (hl == 0 && len > 0 && P.columns) atomicAdd(P.columns, (unsigned long long)len);
  const int wlen = warp_max_int(len);
  __syncwarp();
  // -log() emission terms of node hl at column c (nodePrior, sw-post.c:112-140): (0 - a) - b
  auto node_prior = [&](const PsCol &pc, int node) -> double {
    double val = 0;
    if (pc.let != -2) val = val - (PS_RIGHT(node) == pc.let ? P.la1 : P.la2);
    const bool same = (PS_LEFT(node) ^ PS_RIGHT(node)) == pc.col;
    const double l1 = pc.kind == 0 ? P.lc1 : pc.kind == 2 ? P.ln1 : P.lc1_tab[pc.q];
    const double l2 = pc.kind == 0 ? P.lc2 : pc.kind == 2 ? P.ln2 : P.lc2_tab[pc.q];
    val = val - (same ? l1 : l2);
    return val;
  };
  // ---- do_forwards (sw-post.c:318-361) and do_backwards (:270-316) in ONE loop --------------------------------
  // Iteration t advances the forward recurrence to column t and the backward recurrence to column len - 1 - t: two
  // independent dependency chains per warp, and -- since the logarithm of a forward sum depends on the node's left
  // letter only and that of a backward sum on its right letter only (four distinct values each) -- ONE logarithm
  // per lane serves both: lanes 0-7 of the half-warp take the forward sums, lanes 8-15 the backward sums, and every
  // node fetches its two results by shuffle.  exp(-forwards[i-1][k]) and exp(-(nodePrior(i+1, k) + backwards[i+1][k]))
  // depend on k only: one exponential each per lane, the sums pick their four terms by shuffle in the reference's
  // order of k.  All shuffles are full-mask and unconditional: both halves of the warp execute them whatever their
  // own state.
  double f = 0, run_scale = 0, b = 0, bscale = 0;
  const bool fwd_lane = hl < 8;
  const int q4 = hl & 3;
  PsCol blank;
  blank.let = -2; blank.col = 0; blank.kind = 0; blank.q = 0; blank.call = 15; blank.maxp = 0; blank.qual = 33; blank.pad = 0;
  for (int t = 0; t < wlen; t++) {
    const bool on = t < len;
    const int c = len - 1 - t;   // backward column
    const PsCol pc = on ? cols[t] : blank;
    double nf, nb;
    if (t == 0) {
      nf = PS_LEFT(hl) == init_bp ? node_prior(pc, hl) : HUGE_VAL;
      nb = 0.0;   // last column: backwards = 0, backscale = 0
    } else {
      const PsCol nxt = on ? cols[c + 1] : blank;
      const double val = node_prior(pc, hl);
      const double ef = PS_EXP(-1 * (f));
      const double eb = PS_EXP(-1 * (node_prior(nxt, hl) + b));
      double s = 0;
#pragma unroll
      for (int m = 0; m < 4; m++) {
        const int src = fwd_lane ? 4 * m + q4 : 4 * q4 + m;
        const double xf = __shfl_sync(0xffffffffu, ef, src, 16), xb = __shfl_sync(0xffffffffu, eb, src, 16);
        s += fwd_lane ? xf : xb;
      }
      const double lg = PS_LOG(s);
      const double lf = __shfl_sync(0xffffffffu, lg, PS_LEFT(hl), 16), lb = __shfl_sync(0xffffffffu, lg, 8 + PS_RIGHT(hl), 16);
      nf = val - lf;
      nb = on ? -lb : 0.0;
    }
    // forwscale / backscale: minimum over the nodes (column 0: over the nodes that start from the initial base; the
    // others are +infinity and never the minimum)
    const double scf = half_min(nf), scb = half_min(nb);
    nf -= scf;
    nb -= scb;
    if (on) {
      f = nf;
      run_scale = t == 0 ? scf : scf + run_scale;
      fw[(size_t)t * 16 + hl] = f;
      b = nb;
      bscale = t == 0 ? scb : scb + bscale;
      bw[(size_t)c * 16 + hl] = b;
      if (hl == 0) {
        fscale[t] = run_scale;
        bsc[c] = bscale;
      }
    }
  }
  double total_score = 0;
  {
    const double e = PS_EXP(-1 * (f));
    double val = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) val += __shfl_sync(0xffffffffu, e, j, 16);
    total_score = -PS_LOG(val) + run_scale;
  }
  __syncwarp();
  // ---- posteriors of the four letters per column, post_traceback (:182-207), get_base_qualities (:584-601) -------
  // no dependency between the columns: the loads of the next ones overlap the exponential of this one
  for (int i = 0; i < wlen; i++) {
    const bool on = i < len;
    const double fwv = on ? fw[(size_t)i * 16 + hl] : 0.0, bwv = on ? bw[(size_t)i * 16 + hl] : 0.0;
    const double fs = on ? fscale[i] : 0.0, bs = on ? bsc[i] : 0.0;
    const int bc = on ? (int)cols[i].call : 15;
    const double e = PS_EXP(-1 * (fwv + bwv + fs + bs - total_score));
    double p = 0;
#pragma unroll
    for (int m = 0; m < 4; m++) p += __shfl_sync(0xffffffffu, e, 4 * m + (hl & 3), 16);
    const double p0 = __shfl_sync(0xffffffffu, p, 0, 16), p1 = __shfl_sync(0xffffffffu, p, 1, 16);
    const double p2 = __shfl_sync(0xffffffffu, p, 2, 16), p3 = __shfl_sync(0xffffffffu, p, 3, 16);
    if (on && hl == 0) {
      int maxval = 0;
      double pm = p0;
      if (p1 > pm) { maxval = 1; pm = p1; }
      if (p2 > pm) { maxval = 2; pm = p2; }
      if (p3 > pm) { maxval = 3; pm = p3; }
      // the base quality needs a log(): park 1 - posterior[base call] in the column's forwscale slot (dead from
      // here on: the shuffles above come after every lane's read) and take the logarithms of all columns together
      // afterwards, a lane per column
      fscale[i] = bc != 15 ? 1 - (bc == 0 ? p0 : bc == 1 ? p1 : bc == 2 ? p2 : p3) : 2.0;
      cols[i].maxp = (int8_t)maxval;
    }
  }
  __syncwarp();
  {   // get_base_qualities (sw-post.c:584-601)
    const double log10v = PS_LOG(10.0);
    for (int c = hl; c < len; c += 16) {
      const double pe = fscale[c];
      int tmp = pe == 2.0 ? 0 : ps_qv_from_pr_err(pe, log10v, GT);
      if (tmp > 40) tmp = 40;
      cols[c].qual = (uint8_t)(33 + tmp);
    }
  }
  __syncwarp();
  // ---- fix_base_calls (:548-581) on all lanes, get_posterior (:604-626) by one lane over the gap columns only ----
  {
    uint8_t *wops = P.ops + (size_t)slot * (size_t)P.ops_stride;
    uint8_t *qout = P.quals_out + (size_t)slot * (size_t)P.max_rlen;
    const uint64_t g0 = (uint64_t)T.goff_global + (uint64_t)(R.genome_start - (int)T.goff_contig);
    const bool doit = run && len > 0;
    const int n_ops3 = doit ? n_ops : 0;
    int matches = 0, mismatches = 0, crossovers = 0;
    int colbase = 0, genbase = 0, last_type = 0;
    double res = (doit && hl == 0) ? PS_EXP(-total_score) : 0.0;
    for (int q0 = 0; q0 < w_ops; q0 += 16) {
      const int q = q0 + hl;
      const bool in = q < n_ops3;
      const int op = in ? (int)wops[R.ops_start + q] : 0, type = in ? (op & 3) : 0;
      const int isread = (in && type != 1) ? 1 : 0, isgen = (in && type != 2) ? 1 : 0;
      int ir = isread, ig = isgen;
#pragma unroll
      for (int d = 1; d < 16; d <<= 1) {
        const int r2 = __shfl_up_sync(0xffffffffu, ir, d, 16), g2 = __shfl_up_sync(0xffffffffu, ig, d, 16);
        if (hl >= d) {
          ir += r2;
          ig += g2;
        }
      }
      // type of the previous column (for the gap-open test)
      int ptype = __shfl_up_sync(0xffffffffu, type, 1, 16);
      if (hl == 0) ptype = last_type;
      const int col = colbase + ir - isread;
      if (isread) {
        const PsCol pc = cols[col];
        const int crt = pc.maxp;
        const int prev_base = col == 0 ? init_bp : (int)cols[col - 1].maxp;
        const bool lower = (prev_base ^ crt) != pc.col;
        if (lower) crossovers++;
        if (type == 3) {
          const int g = (int)extract4(genome, g0 + (uint64_t)(genbase + ig - isgen));
          if (g == crt) matches++;
          else mismatches++;
        }
        // bit 3: bits 4-5 are the base call itself (not a layer); bit 2: lower case (crossover before this base)
        wops[R.ops_start + q] = (uint8_t)(type | (lower ? 4 : 0) | 8 | (crt << 4));
        qout[col] = pc.qual;
      }
      // gap columns, in script order: a genome base against a gap multiplies by the deletion terms, a read base
      // against a gap by the insertion terms; the open term when the previous column was not the same kind of gap
      const bool gap = in && type != 3;
      const bool opens = gap && (q == 0 || ptype != type);
      uint32_t gm = (__ballot_sync(0xffffffffu, gap) >> hshift) & 0xffffu;
      const uint32_t dm = (__ballot_sync(0xffffffffu, gap && type == 1) >> hshift) & 0xffffu;
      const uint32_t om = (__ballot_sync(0xffffffffu, opens) >> hshift) & 0xffffu;
      if (hl == 0)
        while (gm) {
          const int bpos = __ffs(gm) - 1;
          gm &= gm - 1;
          const bool del = (dm >> bpos) & 1u;
          res *= del ? P.pr_del_extend : P.pr_ins_extend;
          if ((om >> bpos) & 1u) res *= del ? P.pr_del_open : P.pr_ins_open;
        }
      colbase += __shfl_sync(0xffffffffu, ir, 15, 16);
      genbase += __shfl_sync(0xffffffffu, ig, 15, 16);
      const int lt_ = __shfl_sync(0xffffffffu, type, 15, 16);
      last_type = lt_;
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      matches += __shfl_xor_sync(0xffffffffu, matches, o, 16);
      mismatches += __shfl_xor_sync(0xffffffffu, mismatches, o, 16);
      crossovers += __shfl_xor_sync(0xffffffffu, crossovers, o, 16);
    }
    if (doit && hl == 0) {
      R.matches = matches;
      R.mismatches = mismatches;
      R.crossovers = crossovers;
      R.posterior = res;
      {   // mapping.c:1619-1621 (-fmad=false: the host's operations in the host's order)
        const double ps = P.score_alpha * PS_LOG(res) / P.log2v + (double)R.rmapped * P.score_2ab;
        int v = (int)rint(ps);
        R.post_score = v < 0 ? 0 : v;
        R.pad_ = 0;
      }
      P.results[slot] = R;
    }
  }
  __syncwarp();   // the next group reuses the columns
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Four lanes per alignment, four nodes per lane (round 2).  The 16-node chain factorises: the forward sum of a node
// depends on its LEFT letter only -- F_q = sum over m of exp(-f[(m, q)]), f'[(l, r)] = prior(l, r) - log F_l -- and
// the backward sum on its RIGHT letter only -- B_q = sum over m of exp(-(prior'(q, m) + b[(q, m)])),
// b'[(l, r)] = -log B_r (so backwards[][] has four distinct values per column).  Lane q of a quad owns the four
// nodes (m, q) of the forward chain and the four nodes (q, m) of the backward chain: both sums are LOCAL, in the
// reference's order of m, and only the four logarithms travel (an all-gather over four lanes).  Per column a quad
// evaluates 32 exponentials and 8 logarithms -- what the recurrences need, none twice -- against 32 + 16 on a
// half-warp, with 20 shuffles instead of 80, and a warp advances eight alignments per iteration instead of two.
// Every double is computed by the same operations in the same order as in the half-warp kernel above (and in
// sw-post.c); tests/test_gpu_post_sw.py runs both.
#define PQ_W 4
// nodePrior (sw-post.c:112-140) of the nodes of one column: val = 0; val -= (right == letter ? la1 : la2) when the
// column emits a letter; val -= ((left ^ right) == colour ? l1 : l2).  Four distinct values per column: the
// selections are made once per column, a node picks its value by two compares.
struct PsEmit {
  double v[2][2];   // [right == letter ? 0 : 1][colour matches ? 0 : 1]
  int let, col;
};
__device__ __forceinline__ PsEmit ps_emit(const PostParams &P, const PsCol &pc) {
  PsEmit E;
  const double l1 = pc.kind == 0 ? P.lc1 : pc.kind == 2 ? P.ln1 : P.lc1_tab[pc.q];
  const double l2 = pc.kind == 0 ? P.lc2 : pc.kind == 2 ? P.ln2 : P.lc2_tab[pc.q];
  const bool has = pc.let != -2;
  const double a1 = has ? 0.0 - P.la1 : 0.0, a2 = has ? 0.0 - P.la2 : 0.0;
  E.v[0][0] = a1 - l1;
  E.v[0][1] = a1 - l2;
  E.v[1][0] = a2 - l1;
  E.v[1][1] = a2 - l2;
  E.let = has ? pc.let : 4;   // 4: no letter of the node equals it; a1 == a2 == 0 then
  E.col = pc.col;
  return E;
}
__device__ __forceinline__ double ps_prior(const PsEmit &E, int l, int r) {
  const bool lm = r == E.let, cm = (l ^ r) == E.col;
  const double m0 = cm ? E.v[0][0] : E.v[0][1], m1 = cm ? E.v[1][0] : E.v[1][1];
  return lm ? m0 : m1;
}

// MINB (resident CTAs per SM the register allocation aims at): 6 = 80 registers measured fastest (12.0 ms per C2 step;
// 4 = 128 registers: 12.5, 8 = 64 registers with spills: 13.3)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) post_sw_quad_kernel(const PostParams P, int quads_per_cta, int max_cols,
                                                                 int n_groups, double *fw_scratch) {
  extern __shared__ double ps_smem[];
  const int lane = threadIdx.x & 31, ql = lane & 3;
  const int qidx = threadIdx.x >> 2;   // quad of the CTA
  double *fscale = ps_smem + (size_t)qidx * ps_smem_doubles_per_half(max_cols);
  double *bsc = fscale + max_cols;
  PsCol *cols = (PsCol *)(bsc + max_cols);
  uint8_t *pxs = (uint8_t *)(cols + max_cols);
  // forwards[max_cols][16] and the four distinct backwards[max_cols][.] of this quad
  double *fw = fw_scratch + ((size_t)blockIdx.x * quads_per_cta + qidx) * (size_t)max_cols * 20;
  double *bwr = fw + (size_t)max_cols * 16;
  const glibc_math::Tables GT = {P.gm_tab, P.gm_tab + 8, P.gm_tab + 8 + 256, P.gm_tab + 8 + 256 + 18};
  const int hshift = lane & ~3;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
  int slot = grp * quads_per_cta + qidx;
  const bool live = slot < P.n_tasks;
  if (!live) slot = P.n_tasks - 1;
  const FullTask T = P.tasks[slot];
  FullResult R = P.results[slot];
  const bool run = live && T.run && R.score > 0;
  const uint32_t *genome = T.gen_st ? P.genome_rc : P.genome_fwd;
  const uint32_t *read = P.reads + (size_t)T.ridx * P.stride;
  const uint8_t *ops = P.ops + (size_t)slot * (size_t)P.ops_stride;
  const uint8_t *rq = P.read_quals ? P.read_quals + (size_t)(T.ridx >> 1) * (size_t)P.qual_stride + P.qual_vector_offset
                                   : nullptr;
  const int init_bp = T.initbp;
  // ---- load_local_vectors (sw-post.c:472-545), as in the half-warp kernel with scans over four lanes --------------
  const int rlen_t = run ? T.rlen : 0, n_ops = run ? R.ops_len : 0;
  const int w_rlen = warp_max_int(rlen_t), w_ops = warp_max_int(n_ops);
  int start_run = 0, min_qv = 10000;
  {
    int carry = 0, any_n = 0;
    for (int p0 = 0; p0 < w_rlen; p0 += PQ_W) {
      const int p = p0 + ql;
      const bool in = p < rlen_t;
      const int c = in ? (int)extract4(read, (uint64_t)p) : 15;
      int v = c != 15 ? c : 0, f = c == 15 ? 1 : 0;
#pragma unroll
      for (int d = 1; d < PQ_W; d <<= 1) {
        const int v2 = __shfl_up_sync(0xffffffffu, v, d, PQ_W), f2 = __shfl_up_sync(0xffffffffu, f, d, PQ_W);
        if (ql >= d) {
          if (!f) v ^= v2;
          f |= f2;
        }
      }
      if (!f) v ^= carry;
      if (in) pxs[p] = (uint8_t)v;
      carry = __shfl_sync(0xffffffffu, v, PQ_W - 1, PQ_W);
      const bool pre = in && p < R.read_start;
      any_n |= (int)((__ballot_sync(0xffffffffu, pre && c == 15) >> hshift) & 0xfu);
      int x = (pre && c != 15) ? c : 0, mq = (pre && rq) ? (int)rq[p] : 10000;
#pragma unroll
      for (int o = PQ_W / 2; o > 0; o >>= 1) {
        x ^= __shfl_xor_sync(0xffffffffu, x, o, PQ_W);
        mq = min(mq, __shfl_xor_sync(0xffffffffu, mq, o, PQ_W));
      }
      start_run ^= x;
      min_qv = min(min_qv, mq);
    }
    if (any_n) {
      start_run = 15;
      min_qv = 0;
    }
  }
  __syncwarp();
  int len = 0;
  {
    int colbase = 0, genbase = 0;
    const uint64_t g0 = (uint64_t)T.goff_global + (uint64_t)(R.genome_start - (int)T.goff_contig);
    for (int q0 = 0; q0 < w_ops; q0 += PQ_W) {
      const int q = q0 + ql;
      const bool in = q < n_ops;
      const int op = in ? (int)ops[R.ops_start + q] : 0, type = op & 3, kk = (op >> 4) & 3;
      const int isread = (in && type != 1) ? 1 : 0, isgen = (in && type != 2) ? 1 : 0;
      int ir = isread, ig = isgen;
#pragma unroll
      for (int d = 1; d < PQ_W; d <<= 1) {
        const int r2 = __shfl_up_sync(0xffffffffu, ir, d, PQ_W), g2 = __shfl_up_sync(0xffffffffu, ig, d, PQ_W);
        if (ql >= d) {
          ir += r2;
          ig += g2;
        }
      }
      const int col = colbase + ir - isread;
      if (isread && col < max_cols) {
        const int j = R.read_start + col;
        const int c = (int)extract4(read, (uint64_t)j);
        PsCol pc;
        int gl = 15;
        if (type == 3) {
          const int g = (int)extract4(genome, g0 + (uint64_t)(genbase + ig - isgen));
          pc.let = (int8_t)(g <= 3 ? g : -1);
          gl = g <= 3 ? g : 15;
        } else {
          pc.let = -2;
        }
        if ((col == 0 && start_run == 15) || c == 15) {
          pc.col = 0;
          pc.kind = 2;
          pc.q = 0;
        } else {
          pc.col = (int8_t)(c ^ (col == 0 ? start_run : 0));
          if (rq) {
            pc.kind = 1;
            pc.q = (uint8_t)(col == 0 ? min(min_qv, (int)rq[j]) : (int)rq[j]);
          } else {
            pc.kind = 0;
            pc.q = 0;
          }
        }
        pc.call = (int8_t)(c == 15 ? gl : (((kk + init_bp) & 3) ^ (int)pxs[j]));
        pc.maxp = 0;
        pc.qual = 33;
        pc.pad = 0;
        cols[col] = pc;
      }
      colbase += __shfl_sync(0xffffffffu, ir, PQ_W - 1, PQ_W);
      genbase += __shfl_sync(0xffffffffu, ig, PQ_W - 1, PQ_W);
    }
    len = colbase;
  }
  if (len > max_cols) len = 0;
  if (ql == 0 && len > 0 && P.columns) atomicAdd(P.columns, (unsigned long long)len);
  const int wlen = warp_max_int(len);
  __syncwarp();
  // ---- do_forwards (sw-post.c:318-361) and do_backwards (:270-316) in one loop -----------------------------------
  // f[l] = forwards[t][(l, ql)]; b[r] = backwards[c][(ql, r)] (the same four values on every lane of the quad)
  double f[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}, run_scale = 0, bscale = 0;
  PsCol blank;
  blank.let = -2; blank.col = 0; blank.kind = 0; blank.q = 0; blank.call = 15; blank.maxp = 0; blank.qual = 33; blank.pad = 0;
  for (int t = 0; t < wlen; t++) {
    const bool on = t < len;
    const int c = len - 1 - t;   // backward column
    const PsEmit E = ps_emit(P, on ? cols[t] : blank);
    double nf[4], nb[4];
    if (t == 0) {
#pragma unroll
      for (int l = 0; l < 4; l++) {
        nf[l] = l == init_bp ? ps_prior(E, l, ql) : HUGE_VAL;
        nb[l] = 0.0;
      }
    } else {
      const PsEmit En = ps_emit(P, on ? cols[c + 1] : blank);
      double sf = 0, sb = 0;
#pragma unroll
      for (int m = 0; m < 4; m++) {
        sf += PS_EXP(-1 * (f[m]));
        sb += PS_EXP(-1 * (ps_prior(En, ql, m) + b[m]));
      }
      const double lgf = PS_LOG(sf), lgb = PS_LOG(sb);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const double lf = __shfl_sync(0xffffffffu, lgf, k, PQ_W), lb = __shfl_sync(0xffffffffu, lgb, k, PQ_W);
        nf[k] = ps_prior(E, k, ql) - lf;
        nb[k] = on ? -lb : 0.0;
      }
    }
    // forwscale: minimum over the 16 nodes; backscale: over the four distinct values
    double scf = nf[0], scb = nb[0];
#pragma unroll
    for (int k = 1; k < 4; k++) {
      scf = nf[k] < scf ? nf[k] : scf;
      scb = nb[k] < scb ? nb[k] : scb;
    }
#pragma unroll
    for (int o = PQ_W / 2; o > 0; o >>= 1) {
      const double w = __shfl_xor_sync(0xffffffffu, scf, o, PQ_W);
      scf = w < scf ? w : scf;
    }
    if (on) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        f[k] = nf[k] - scf;
        b[k] = nb[k] - scb;
        fw[(size_t)t * 16 + k * 4 + ql] = f[k];
      }
      run_scale = t == 0 ? scf : scf + run_scale;
      bscale = t == 0 ? scb : scb + bscale;
      bwr[(size_t)c * 4 + ql] = ql == 0 ? b[0] : ql == 1 ? b[1] : ql == 2 ? b[2] : b[3];
      if (ql == 0) {
        fscale[t] = run_scale;
        bsc[c] = bscale;
      }
    }
  }
  double total_score = 0;
  {
    double e[4];
#pragma unroll
    for (int l = 0; l < 4; l++) e[l] = PS_EXP(-1 * (f[l]));
    double val = 0;
#pragma unroll
    for (int l = 0; l < 4; l++)
#pragma unroll
      for (int r = 0; r < 4; r++) val += __shfl_sync(0xffffffffu, e[l], r, PQ_W);   // node order j = 4 l + r
    total_score = -PS_LOG(val) + run_scale;
  }
  __syncwarp();
  // ---- posteriors of the four letters per column, post_traceback (:182-207), get_base_qualities (:584-601) -------
  for (int i = 0; i < wlen; i++) {
    const bool on = i < len;
    const double bwv = on ? bwr[(size_t)i * 4 + ql] : 0.0;
    const double fs = on ? fscale[i] : 0.0, bs = on ? bsc[i] : 0.0;
    const int bc = on ? (int)cols[i].call : 15;
    double p = 0;   // posterior of letter ql: the nodes (m, ql) in the order of m
#pragma unroll
    for (int m = 0; m < 4; m++) {
      const double fwv = on ? fw[(size_t)i * 16 + m * 4 + ql] : 0.0;
      p += PS_EXP(-1 * (fwv + bwv + fs + bs - total_score));
    }
    const double p0 = __shfl_sync(0xffffffffu, p, 0, PQ_W), p1 = __shfl_sync(0xffffffffu, p, 1, PQ_W);
    const double p2 = __shfl_sync(0xffffffffu, p, 2, PQ_W), p3 = __shfl_sync(0xffffffffu, p, 3, PQ_W);
    if (on && ql == 0) {
      int maxval = 0;
      double pm = p0;
      if (p1 > pm) { maxval = 1; pm = p1; }
      if (p2 > pm) { maxval = 2; pm = p2; }
      if (p3 > pm) { maxval = 3; pm = p3; }
      fscale[i] = bc != 15 ? 1 - (bc == 0 ? p0 : bc == 1 ? p1 : bc == 2 ? p2 : p3) : 2.0;
      cols[i].maxp = (int8_t)maxval;
    }
  }
  __syncwarp();
  {   // get_base_qualities (sw-post.c:584-601)
    const double log10v = PS_LOG(10.0);
    for (int c = ql; c < len; c += PQ_W) {
      const double pe = fscale[c];
      int tmp = pe == 2.0 ? 0 : ps_qv_from_pr_err(pe, log10v, GT);
      if (tmp > 40) tmp = 40;
      cols[c].qual = (uint8_t)(33 + tmp);
    }
  }
  __syncwarp();
  // ---- fix_base_calls (:548-581) on all lanes, get_posterior (:604-626) by one lane over the gap columns only ----
  {
    uint8_t *wops = P.ops + (size_t)slot * (size_t)P.ops_stride;
    uint8_t *qout = P.quals_out + (size_t)slot * (size_t)P.max_rlen;
    const uint64_t g0 = (uint64_t)T.goff_global + (uint64_t)(R.genome_start - (int)T.goff_contig);
    const bool doit = run && len > 0;
    const int n_ops3 = doit ? n_ops : 0;
    int matches = 0, mismatches = 0, crossovers = 0;
    int colbase = 0, genbase = 0, last_type = 0;
    double res = (doit && ql == 0) ? PS_EXP(-total_score) : 0.0;
    for (int q0 = 0; q0 < w_ops; q0 += PQ_W) {
      const int q = q0 + ql;
      const bool in = q < n_ops3;
      const int op = in ? (int)wops[R.ops_start + q] : 0, type = in ? (op & 3) : 0;
      const int isread = (in && type != 1) ? 1 : 0, isgen = (in && type != 2) ? 1 : 0;
      int ir = isread, ig = isgen;
#pragma unroll
      for (int d = 1; d < PQ_W; d <<= 1) {
        const int r2 = __shfl_up_sync(0xffffffffu, ir, d, PQ_W), g2 = __shfl_up_sync(0xffffffffu, ig, d, PQ_W);
        if (ql >= d) {
          ir += r2;
          ig += g2;
        }
      }
      int ptype = __shfl_up_sync(0xffffffffu, type, 1, PQ_W);
      if (ql == 0) ptype = last_type;
      const int col = colbase + ir - isread;
      if (isread) {
        const PsCol pc = cols[col];
        const int crt = pc.maxp;
        const int prev_base = col == 0 ? init_bp : (int)cols[col - 1].maxp;
        const bool lower = (prev_base ^ crt) != pc.col;
        if (lower) crossovers++;
        if (type == 3) {
          const int g = (int)extract4(genome, g0 + (uint64_t)(genbase + ig - isgen));
          if (g == crt) matches++;
          else mismatches++;
        }
        wops[R.ops_start + q] = (uint8_t)(type | (lower ? 4 : 0) | 8 | (crt << 4));
        qout[col] = pc.qual;
      }
      const bool gap = in && type != 3;
      const bool opens = gap && (q == 0 || ptype != type);
      uint32_t gm = (__ballot_sync(0xffffffffu, gap) >> hshift) & 0xfu;
      const uint32_t dm = (__ballot_sync(0xffffffffu, gap && type == 1) >> hshift) & 0xfu;
      const uint32_t om = (__ballot_sync(0xffffffffu, opens) >> hshift) & 0xfu;
      if (ql == 0)
        while (gm) {
          const int bpos = __ffs(gm) - 1;
          gm &= gm - 1;
          const bool del = (dm >> bpos) & 1u;
          res *= del ? P.pr_del_extend : P.pr_ins_extend;
          if ((om >> bpos) & 1u) res *= del ? P.pr_del_open : P.pr_ins_open;
        }
      colbase += __shfl_sync(0xffffffffu, ir, PQ_W - 1, PQ_W);
      genbase += __shfl_sync(0xffffffffu, ig, PQ_W - 1, PQ_W);
      last_type = __shfl_sync(0xffffffffu, type, PQ_W - 1, PQ_W);
    }
#pragma unroll
    for (int o = PQ_W / 2; o > 0; o >>= 1) {
      matches += __shfl_xor_sync(0xffffffffu, matches, o, PQ_W);
      mismatches += __shfl_xor_sync(0xffffffffu, mismatches, o, PQ_W);
      crossovers += __shfl_xor_sync(0xffffffffu, crossovers, o, PQ_W);
    }
    if (doit && ql == 0) {
      R.matches = matches;
      R.mismatches = mismatches;
      R.crossovers = crossovers;
      R.posterior = res;
      {   // mapping.c:1619-1621 (-fmad=false: the host's operations in the host's order)
        const double ps = P.score_alpha * PS_LOG(res) / P.log2v + (double)R.rmapped * P.score_2ab;
        int v = (int)rint(ps);
        R.post_score = v < 0 ? 0 : v;
        R.pad_ = 0;
      }
      P.results[slot] = R;
    }
  }
  __syncwarp();   // the next group reuses the columns
  }
}

static int launch_post_sw_quad(shrimp_gpu_ctx *ctx, const PostParams &P, DevBuf &scratch) {
  const int max_cols = std::max(1, P.max_rlen);
  const size_t per_quad = ps_smem_doubles_per_half(max_cols) * sizeof(double);
  int warps = 4;   // eight alignments per warp
  while (warps > 1 && per_quad * 8 * warps > 64 * 1024) warps >>= 1;
  if (per_quad * 8 * warps > 200 * 1024) {
    set_error("post_sw: reads of %d bases need more shared memory than a CTA has", P.max_rlen);
    return SHRIMP_E_RANGE;
  }
  const int quads = 8 * warps;
  const size_t smem = per_quad * quads;
  auto kern = post_sw_quad_kernel<6>;
  SH_OPT_IN_SMEM(kern, ctx->device);
  int per_sm = 0;
  SH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, quads * 4, smem));
  if (per_sm < 1) per_sm = 1;
  const int n_groups = (P.n_tasks + quads - 1) / quads;
  const int grid = std::min(n_groups, ctx->sm_count * per_sm);
  SH_TRY(scratch.ensure((size_t)grid * quads * (size_t)max_cols * 20 * sizeof(double)));
  kern<<<grid, quads * 4, smem, ctx->stream>>>(P, quads, max_cols, n_groups, scratch.as<double>());
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_POST);
  return SHRIMP_OK;
}

int launch_post_sw(shrimp_gpu_ctx *ctx, const PostParams &P, DevBuf &scratch) {
  if (P.n_tasks <= 0) return SHRIMP_OK;
  if (!getenv("SHRIMP_POST_SW_HALF")) return launch_post_sw_quad(ctx, P, scratch);   // the half-warp kernel: A/B and tests
  const int max_cols = std::max(1, P.max_rlen);
  const size_t per_half = ps_smem_doubles_per_half(max_cols) * sizeof(double);
  const int halves = 8;
  if (per_half * halves > 200 * 1024) {
    set_error("post_sw: reads of %d bases need more shared memory than a CTA has", P.max_rlen);
    return SHRIMP_E_RANGE;
  }
  const size_t smem = per_half * halves;
  auto kern = post_sw_kernel;
  SH_OPT_IN_SMEM(kern, ctx->device);
  int per_sm = 0;
  SH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, halves * 16, smem));
  if (per_sm < 1) per_sm = 1;
  const int n_groups = (P.n_tasks + halves - 1) / halves;
  const int grid = std::min(n_groups, ctx->sm_count * per_sm);
  SH_TRY(scratch.ensure((size_t)grid * halves * (size_t)max_cols * 32 * sizeof(double)));
  kern<<<grid, halves * 16, smem, ctx->stream>>>(P, halves, max_cols, n_groups, scratch.as<double>());
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_POST);
  return SHRIMP_OK;
}

__global__ void glibc_explog_kernel(const double *x, int n, const unsigned long long *gm_tab, double *e, double *l) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const glibc_math::Tables GT = {gm_tab, gm_tab + 8, gm_tab + 8 + 256, gm_tab + 8 + 256 + 18};
  e[i] = glibc_math::exp_glibc(x[i], GT);
  l[i] = glibc_math::log_glibc(x[i], GT);
}

}  // namespace shrimp

// Diagnostic entry: exp() and log() of n doubles as the device computes them in post_sw (glibc_math.cuh), for the
// test that compares them bit for bit with the host's libm.
extern "C" int shrimp_gpu_glibc_explog(shrimp_gpu_ctx *ctx, const double *x, int n, double *exp_out, double *log_out) {
  using namespace shrimp;
  if (!ctx || !x || !exp_out || !log_out || n < 0) {
    set_error("shrimp_gpu_glibc_explog: invalid argument");
    return SHRIMP_E_ARG;
  }
  if (n == 0) return SHRIMP_OK;
  std::vector<unsigned long long> gm(8 + 256 + 18 + 256);
  memcpy(gm.data(), GLIBC_EXP_CONST, 8 * 8);
  memcpy(gm.data() + 8, GLIBC_EXP_TAB, 256 * 8);
  memcpy(gm.data() + 8 + 256, GLIBC_LOG_CONST, 18 * 8);
  memcpy(gm.data() + 8 + 256 + 18, GLIBC_LOG_TAB, 256 * 8);
  DevBuf dt, dx, de, dl;
  int rc = SHRIMP_OK;
  if ((rc = dt.ensure(gm.size() * 8)) == SHRIMP_OK && (rc = dx.ensure((size_t)n * 8)) == SHRIMP_OK &&
      (rc = de.ensure((size_t)n * 8)) == SHRIMP_OK && (rc = dl.ensure((size_t)n * 8)) == SHRIMP_OK) {
    cudaMemcpyAsync(dt.p, gm.data(), gm.size() * 8, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dx.p, x, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
    glibc_explog_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(dx.as<double>(), n, dt.as<unsigned long long>(),
                                                                   de.as<double>(), dl.as<double>());
    ctx->launches++;
    cudaMemcpyAsync(exp_out, de.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(log_out, dl.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      set_error("shrimp_gpu_glibc_explog: %s", cudaGetErrorString(e));
      rc = SHRIMP_E_CUDA;
    }
  }
  dt.release(); dx.release(); de.release(); dl.release();
  return rc;
}
