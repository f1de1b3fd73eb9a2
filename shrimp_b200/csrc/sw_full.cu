// Full Smith-Waterman with traceback, letter space.
//
// Replaces common/sw-full-ls.c: full_sw :154-403 (banded 3-state affine DP with back-pointers,
// global-in-read by default, optional local mode with early exit and threshold-band retry),
// do_backtrace :413-516 and the coordinates of sw_full_ls :637-683; band geometry from
// common/anchors.c (anchor_join :9-54, anchor_widen :57-63, anchor_get_x_range :66-95).
//
// Round-1 mapping of the work: one thread per alignment (<= 30 per read and thousands of reads per
// chunk give enough parallelism; each alignment is only ~1-2 k band cells).  All per-alignment state
// lives in global scratch laid out [cell][task] so that a warp's accesses coalesce:
//   row[3][glen+1][NT]  rolling DP row (north, west, northwest scores per column)
//   bp[rlen*glen][NT]   one byte of back-pointers per cell (2 bits per state)
//   ops[NT][rlen+glen]  edit script per task, filled from the end like the reference's backtrace buffer
// The reference keeps a full (dblen+1)x(qrlen+1) matrix of 16-byte cells per thread.
#include "band.cuh"

namespace shrimp {





// back-pointer byte: bits 0-1 northwest state (0 none, 1 from north, 2 from northwest, 3 from west),
// bits 2-3 north state (0 none, 1 N<-N, 2 N<-NW), bits 4-5 west state (0 none, 1 W<-NW, 2 W<-W)
enum { ST_NW = 0, ST_N = 1, ST_W = 2 };

template <bool LOCAL>
__device__ int full_sw_ls_dev(const FullParams &P, const FullTask &T, int t, const uint32_t *genome,
                              const uint32_t *read, bool use_anchor, int &ret_i, int &ret_j, int &end_n, int &end_w,
                              int &end_nw, unsigned long long &cells) {
  const int lena = T.glen, lenb = T.rlen;
  const int NT = P.NT;
  const int ao = P.a_open, ae = P.a_ext, bo = P.b_open, be = P.b_ext;
  const bool revcmpl = T.gen_st && P.Tflag;
  Rect rect;
  if (use_anchor && P.anchor_width >= 0) {
    rect.x = T.ax;
    rect.y = T.ay;
    rect.length = T.alen;
    rect.width = T.awidth;
    rect.x -= P.anchor_width / 2;  // anchor_widen
    rect.y += P.anchor_width / 2;
    rect.width += P.anchor_width;
  } else {  // threshold band, sw-full-ls.c:178-191
    long long y0 = (lenb * P.match - T.thresh) / P.match;
    rect = rect_join2(0, y0, 1, 1, lena - 1, lenb - 1 - y0, 1, 1);
  }
  int32_t *rowN = P.row + t, *rowW = rowN + (size_t)(P.max_glen + 1) * NT, *rowNW = rowW + (size_t)(P.max_glen + 1) * NT;
  uint8_t *bp = P.bp + t;
  for (int c = 0; c <= lena; c++) {  // row -1: local-style init (:194-196)
    rowNW[(size_t)c * NT] = 0;
    rowN[(size_t)c * NT] = -bo;
    rowW[(size_t)c * NT] = -ao;
  }
  int score = 0, max_i = 0, max_j = 0;
  const int init_nw = LOCAL ? 0 : NEG_HALF, init_n = LOCAL ? -bo : NEG_HALF, init_w = LOCAL ? -ao : NEG_HALF;
  bool done = false;
  for (int i = 0; i < lenb && !done; i++) {
    int x_min, x_max;
    rect_x_range(rect, lena, i, x_min, x_max);
    const uint32_t q = extract4(read, (uint64_t)i);
    cells += (unsigned long long)(x_max - x_min + 1);
    // left edge cell (i, x_min-1), storage column x_min
    size_t c0 = (size_t)x_min * NT;
    int d_nw = rowNW[c0], d_n = rowN[c0], d_w = rowW[c0];  // cell (i-1, x_min-1)
    rowNW[c0] = init_nw;
    rowN[c0] = init_n;
    rowW[c0] = init_w;
    if (x_min >= 1) bp[((size_t)i * lena + (x_min - 1)) * NT] = 0;
    int l_nw = init_nw, l_w = init_w;  // cell (i, j-1)
    for (int j = x_min; j <= x_max; j++) {
      const size_t c = (size_t)(j + 1) * NT;
      const int u_nw = rowNW[c], u_n = rowN[c], u_w = rowW[c];  // cell (i-1, j)
      const uint32_t d = extract4(genome, (uint64_t)T.goff_global + (uint64_t)j);
      const int ms = (d == q) ? P.match : P.mismatch;
      int tmp, v_nw, v_n, v_w;
      uint32_t b_nw, b_n, b_w;
      // northwest (:261-296)
      if (!revcmpl) {
        tmp = d_nw + ms; b_nw = 2;
        if (d_n + ms > tmp) { tmp = d_n + ms; b_nw = 1; }
        if (d_w + ms > tmp) { tmp = d_w + ms; b_nw = 3; }
      } else {
        tmp = d_w + ms; b_nw = 3;
        if (d_n + ms > tmp) { tmp = d_n + ms; b_nw = 1; }
        if (d_nw + ms > tmp) { tmp = d_nw + ms; b_nw = 2; }
      }
      if (LOCAL && tmp <= 0) { tmp = 0; b_nw = 0; }
      v_nw = tmp;
      // north (:299-324)
      if (!revcmpl) {
        tmp = u_nw - bo - be; b_n = 2;
        if (u_n - be > tmp) { tmp = u_n - be; b_n = 1; }
      } else {
        tmp = u_n - be; b_n = 1;
        if (u_nw - bo - be > tmp) { tmp = u_nw - bo - be; b_n = 2; }
      }
      if (LOCAL && tmp <= 0) { tmp = 0; b_n = 0; }
      v_n = tmp;
      // west (:327-352)
      if (!revcmpl) {
        tmp = l_nw - ao - ae; b_w = 1;
        if (l_w - ae > tmp) { tmp = l_w - ae; b_w = 2; }
      } else {
        tmp = l_w - ae; b_w = 2;
        if (l_nw - ao - ae > tmp) { tmp = l_nw - ao - ae; b_w = 1; }
      }
      if (LOCAL && tmp <= 0) { tmp = 0; b_w = 0; }
      v_w = tmp;
      bp[((size_t)i * lena + j) * NT] = (uint8_t)(b_nw | (b_n << 2) | (b_w << 4));
      d_nw = u_nw; d_n = u_n; d_w = u_w;
      rowNW[c] = v_nw; rowN[c] = v_n; rowW[c] = v_w;
      l_nw = v_nw; l_w = v_w;
      if (LOCAL || i == lenb - 1) {  // (:357-368)
        int best = v_n > v_nw ? v_n : v_nw;
        best = best > v_w ? best : v_w;
        if (best > score) { score = best; max_i = i; max_j = j; end_n = v_n; end_w = v_w; end_nw = v_nw; }
      }
      if (LOCAL && score == T.maxscore) { done = true; break; }
    }
    if (done) break;
    if (i + 1 < lenb) {  // cells right of the band that the next row will read (:376-383)
      int nmin, nmax;
      rect_x_range(rect, lena, i + 1, nmin, nmax);
      for (int j = x_max + 1; j <= nmax; j++) {
        const size_t c = (size_t)(j + 1) * NT;
        rowNW[c] = init_nw;
        rowN[c] = init_n;
        rowW[c] = init_w;
        bp[((size_t)i * lena + j) * NT] = 0;
      }
    }
  }
  ret_i = max_i;
  ret_j = max_j;
  return score;
}

__global__ void __launch_bounds__(128) sw_full_ls_kernel(const FullParams P) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;  // scratch column of this launch
  if (slot >= P.n_tasks) return;
  const int t = P.perm ? P.perm[slot] : slot;              // task id
  const FullTask T = P.tasks[t];
  FullResult R;
  memset(&R, 0, sizeof(R));
  if (!T.run) {
    P.results[t] = R;
    return;
  }
  const uint32_t *genome = T.gen_st ? P.genome_rc : P.genome_fwd;
  const uint32_t *read = P.reads + (size_t)T.ridx * P.stride;
  unsigned long long cells = 0;
  int ei = 0, ej = 0, score, e_n = 0, e_w = 0, e_nw = 0;
  if (P.local) {
    score = full_sw_ls_dev<true>(P, T, slot, genome, read, true, ei, ej, e_n, e_w, e_nw, cells);
    if (score != T.maxscore) {  // :395-398
      e_n = e_w = e_nw = 0;
      score = full_sw_ls_dev<true>(P, T, slot, genome, read, false, ei, ej, e_n, e_w, e_nw, cells);
    }
  } else {
    score = full_sw_ls_dev<false>(P, T, slot, genome, read, true, ei, ej, e_n, e_w, e_nw, cells);
  }
  R.score = score;
  // ---- do_backtrace (:413-516) ----
  const int lena = T.glen, NT = P.NT;
  const uint8_t *bp = P.bp + slot;
  uint8_t *ops = P.ops + (size_t)t * (size_t)(P.max_glen + P.max_rlen);
  int i = ei, j = ej;
  // state of the end cell: northwest unless west is strictly greater, unless north is strictly
  // greater than that (:419-427)
  int st = ST_NW;
  {
    int fromscore = e_nw;
    if (e_w > fromscore) { st = ST_W; fromscore = e_w; }
    if (e_n > fromscore) st = ST_N;
  }
  int k = (T.glen + T.rlen) - 1;
  int read_start = 0, genome_start = 0;
  // `from` of the reference = back-pointer of the current state in the current cell
  auto back_of = [&](int ci, int cj, int state) -> int {
    if (ci < 0 || cj < 0) return 0;
    const uint8_t b = bp[((size_t)ci * lena + cj) * NT];
    return state == ST_NW ? (b & 3) : state == ST_N ? ((b >> 2) & 3) : ((b >> 4) & 3);
  };
  int from = back_of(i, j, st);
  if (score > 0 && from != 0) {
    while (i >= 0 && j >= 0) {
      int next_state;
      if (st == ST_N) {  // FROM_NORTH_*: deletion (read base against a gap)
        ops[k] = 2;
        R.deletions++;
        read_start = i--;
        next_state = (from == 1) ? ST_N : ST_NW;
      } else if (st == ST_W) {  // FROM_WEST_*: insertion (genome base against a gap)
        ops[k] = 1;
        R.insertions++;
        genome_start = j--;
        next_state = (from == 2) ? ST_W : ST_NW;
      } else {
        ops[k] = 3;
        if (extract4(genome, (uint64_t)T.goff_global + (uint64_t)j) == extract4(read, (uint64_t)i)) R.matches++;
        else R.mismatches++;
        read_start = i--;
        genome_start = j--;
        next_state = (from == 1) ? ST_N : (from == 2) ? ST_NW : ST_W;
      }
      st = next_state;
      from = back_of(i, j, st);
      k--;
      if (from == 0) break;
    }
  }
  R.read_start = read_start;
  R.gmapped = ej - genome_start + 1;
  R.genome_start = genome_start + (int)T.goff_contig;
  R.rmapped = ei - read_start + 1;
  R.ops_start = k + 1;
  R.ops_len = (T.glen + T.rlen) - (k + 1);
  P.results[t] = R;
  if (cells) atomicAdd(P.cells, cells);
}

int launch_sw_full_ls(shrimp_gpu_ctx *ctx, const FullParams &P) {
  if (P.n_tasks <= 0) return SHRIMP_OK;
  sw_full_ls_kernel<<<(P.n_tasks + 127) / 128, 128, 0, ctx->stream>>>(P);
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_FULL);
  return SHRIMP_OK;
}

}  // namespace shrimp
