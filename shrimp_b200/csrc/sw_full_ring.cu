// Full Smith-Waterman with traceback -- shared-memory band-ring kernels (letter and colour space).
//
// Same recurrences, tie-breaks and outputs as sw_full.cu / sw_full_cs.cu (which restate
// common/sw-full-ls.c:154-516 and common/sw-full-cs.c:249-937 and stay as the fall-back for very wide
// bands); what changes is where the DP state lives and how the work is laid out:
//   * one thread per alignment, but the rolling DP row is a ring in SHARED memory indexed by the
//     offset inside the row's band, s = j - x_min(i) + 1 (s = 0 is the edge cell (i, x_min-1)), laid
//     out [state][s][thread] so that a warp's accesses are conflict-free.  The band of
//     anchor_get_x_range (anchors.c:66-95) only moves right, so the ring is updated in place:
//     row i reads the previous row at s + (x_min(i) - x_min(i-1)) and writes s.
//   * cells of the previous row that lie right of its band are never stored: reading them yields the
//     initial cell, which is what the reference writes there (sw-full-ls.c:376-383).
//   * back-pointers go to global memory as [row][s][task]: the threads of a warp walk their bands in
//     lockstep, so a warp's store is one contiguous segment (1 B per cell in letter space; colour
//     space packs its 12 five-bit pointers into one 64-bit word per cell).
//   * tasks are bucketed by ring width W (32/64/128/256) on the device (pipeline.cu), each bucket one
//     launch with W * states * 4 B of shared memory per thread.
//   * colour space: the 12+8 cross-layer candidates per layer are reduced to a per-layer first-max
//     plus a running top-2 over layers, which keeps the reference's candidate order (same layer first,
//     then layers 0..3, strict > to replace) with a third of the compare-selects.
#include "band.cuh"

namespace shrimp {

enum { D_N_N = 1, D_N_NW = 2, D_W_NW = 3, D_W_W = 4, D_NW_N = 5, D_NW_NW = 6, D_NW_W = 7 };  // sw-full-cs.c:43-49
#define CSC(layer, dir) ((uint32_t)(((dir) << 2) | (layer)))
enum { ST_NW = 0, ST_N = 1, ST_W = 2 };

__device__ __forceinline__ int cstols_r(int first_letter, int colour) {  // util.h:157-180
  if (first_letter == 15 || colour < 0 || colour > 3) return 15;
  return (first_letter % 2 == 0) ? (4 + first_letter + colour) % 4 : (4 + first_letter - colour) % 4;
}

// ------------------------------------------------------------------------------------------------
// letter space
// ------------------------------------------------------------------------------------------------
template <bool LOCAL, int BLOCK, bool REV>
__device__ int ring_ls_dp(const FullParams &P, const FullTask &T, int slot, int32_t *sm, const uint32_t *genome,
                          const uint32_t *read, const Rect &rect, int &ret_i, int &ret_j, int &end_n, int &end_w,
                          int &end_nw, unsigned long long &cells) {
  const int W = P.W;
#define SMR(st, s) sm[((s) * 3 + (st)) * BLOCK]   // [band offset][state][thread]: one pointer, constant state offsets
  const int lena = T.glen, lenb = T.rlen;
  const size_t NT = (size_t)P.NT;
  const int ao = P.a_open, ae = P.a_ext, bo = P.b_open, be = P.b_ext;
  constexpr bool revcmpl = REV;   // T.gen_st && P.Tflag of every task of the launch (the tasks are grouped by it)
  uint8_t *bp = P.bp + slot;
  int score = 0, max_i = 0, max_j = 0;
  const int init_nw = LOCAL ? 0 : NEG_HALF, init_n = LOCAL ? -bo : NEG_HALF, init_w = LOCAL ? -ao : NEG_HALF;
  int pxmin = 0, pxmax = -1;
  bool done = false;
  for (int i = 0; i < lenb && !done; i++) {
    int x_min, x_max;
    rect_x_range(rect, lena, i, x_min, x_max);
    const uint32_t q = extract4(read, (uint64_t)i);
    cells += (unsigned long long)(x_max - x_min + 1);
    // cell (i-1, x_min-1)
    int d_nw, d_n, d_w;
    if (i == 0) { d_nw = 0; d_n = -bo; d_w = -ao; }
    else if (x_min - 1 > pxmax) { d_nw = init_nw; d_n = init_n; d_w = init_w; }
    else { const int sp = x_min - pxmin; d_nw = SMR(0, sp); d_n = SMR(1, sp); d_w = SMR(2, sp); }
    SMR(0, 0) = init_nw; SMR(1, 0) = init_n; SMR(2, 0) = init_w;   // edge cell (i, x_min-1)
    bp[((size_t)i * W) * NT] = 0;
    int l_nw = init_nw, l_w = init_w;
    const int delta = x_min - pxmin;
    uint32_t gpos = T.goff_global + (uint32_t)x_min;
    uint32_t gword = genome[gpos >> 3];
#pragma unroll 4
    for (int j = x_min; j <= x_max; j++) {
      const int s = j - x_min + 1;
      int u_nw, u_n, u_w;  // cell (i-1, j)
      if (i == 0) { u_nw = 0; u_n = -bo; u_w = -ao; }
      else if (j > pxmax) { u_nw = init_nw; u_n = init_n; u_w = init_w; }
      else { const int sp = s + delta; u_nw = SMR(0, sp); u_n = SMR(1, sp); u_w = SMR(2, sp); }
      if ((gpos & 7u) == 0u) gword = genome[gpos >> 3];
      const uint32_t d = (gword >> (4u * (gpos & 7u))) & 15u;
      gpos++;
      const int ms = (d == q) ? P.match : P.mismatch;
      int tmp, v_nw, v_n, v_w;
      uint32_t b_nw, b_n, b_w;
      if (!revcmpl) {  // northwest (sw-full-ls.c:261-296)
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
      if (!revcmpl) {  // north (:299-324)
        tmp = u_nw - bo - be; b_n = 2;
        if (u_n - be > tmp) { tmp = u_n - be; b_n = 1; }
      } else {
        tmp = u_n - be; b_n = 1;
        if (u_nw - bo - be > tmp) { tmp = u_nw - bo - be; b_n = 2; }
      }
      if (LOCAL && tmp <= 0) { tmp = 0; b_n = 0; }
      v_n = tmp;
      if (!revcmpl) {  // west (:327-352)
        tmp = l_nw - ao - ae; b_w = 1;
        if (l_w - ae > tmp) { tmp = l_w - ae; b_w = 2; }
      } else {
        tmp = l_w - ae; b_w = 2;
        if (l_nw - ao - ae > tmp) { tmp = l_nw - ao - ae; b_w = 1; }
      }
      if (LOCAL && tmp <= 0) { tmp = 0; b_w = 0; }
      v_w = tmp;
      bp[((size_t)i * W + s) * NT] = (uint8_t)(b_nw | (b_n << 2) | (b_w << 4));
      d_nw = u_nw; d_n = u_n; d_w = u_w;
      SMR(0, s) = v_nw; SMR(1, s) = v_n; SMR(2, s) = v_w;
      l_nw = v_nw; l_w = v_w;
      if (LOCAL || i == lenb - 1) {  // (:357-368)
        int best = v_n > v_nw ? v_n : v_nw;
        best = best > v_w ? best : v_w;
        if (best > score) { score = best; max_i = i; max_j = j; end_n = v_n; end_w = v_w; end_nw = v_nw; }
      }
      if (LOCAL && score == T.maxscore) { done = true; break; }
    }
    pxmin = x_min;
    pxmax = x_max;
  }
#undef SMR
  ret_i = max_i;
  ret_j = max_j;
  return score;
}

template <int BLOCK, bool REV>
__global__ void __launch_bounds__(BLOCK) sw_full_ls_ring_kernel(const FullParams P) {
  extern __shared__ int32_t ring_smem[];
  const int slot = blockIdx.x * BLOCK + threadIdx.x;
  if (slot >= P.n_tasks) return;
  const int t = P.perm ? P.perm[slot] : slot;
  const FullTask T = P.tasks[t];
  FullResult R;
  memset(&R, 0, sizeof(R));
  if (!T.run) {
    P.results[t] = R;
    return;
  }
  int32_t *sm = ring_smem + threadIdx.x;
  const uint32_t *genome = T.gen_st ? P.genome_rc : P.genome_fwd;
  const uint32_t *read = P.reads + (size_t)T.ridx * P.stride;
  unsigned long long cells = 0;
  int ei = 0, ej = 0, score, e_n = 0, e_w = 0, e_nw = 0;
  Rect rect = task_rect(T, P.anchor_width, P.match, true);
  if (P.local) {
    score = ring_ls_dp<true, BLOCK, REV>(P, T, slot, sm, genome, read, rect, ei, ej, e_n, e_w, e_nw, cells);
    if (score != T.maxscore) {  // sw-full-ls.c:395-398: redo with the threshold band
      e_n = e_w = e_nw = 0;
      rect = task_rect(T, P.anchor_width, P.match, false);
      score = ring_ls_dp<true, BLOCK, REV>(P, T, slot, sm, genome, read, rect, ei, ej, e_n, e_w, e_nw, cells);
    }
  } else {
    score = ring_ls_dp<false, BLOCK, REV>(P, T, slot, sm, genome, read, rect, ei, ej, e_n, e_w, e_nw, cells);
  }
  R.score = score;
  // ---- do_backtrace (sw-full-ls.c:413-516) ----
  const int W = P.W;
  const size_t NT = (size_t)P.NT;
  const uint8_t *bp = P.bp + slot;
  uint8_t *ops = P.ops + (size_t)t * (size_t)(P.max_glen + P.max_rlen);
  int i = ei, j = ej;
  int st = ST_NW;
  {
    int fromscore = e_nw;
    if (e_w > fromscore) { st = ST_W; fromscore = e_w; }
    if (e_n > fromscore) st = ST_N;
  }
  int k = (T.glen + T.rlen) - 1;
  int read_start = 0, genome_start = 0;
  auto back_of = [&](int ci, int cj, int state) -> int {
    if (ci < 0 || cj < 0) return 0;
    int xmn, xmx;
    rect_x_range(rect, T.glen, ci, xmn, xmx);
    const int s = cj - xmn + 1;
    if (s <= 0 || cj > xmx) return 0;  // edge cell / right of the band: back-pointer 0 in the reference
    const uint8_t b = bp[((size_t)ci * W + s) * NT];
    return state == ST_NW ? (b & 3) : state == ST_N ? ((b >> 2) & 3) : ((b >> 4) & 3);
  };
  int from = back_of(i, j, st);
  if (score > 0 && from != 0) {
    while (i >= 0 && j >= 0) {
      int next_state;
      if (st == ST_N) {
        ops[k] = 2;
        R.deletions++;
        read_start = i--;
        next_state = (from == 1) ? ST_N : ST_NW;
      } else if (st == ST_W) {
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

// ------------------------------------------------------------------------------------------------
// colour space
// ------------------------------------------------------------------------------------------------
// running top-2 over layers (first index wins ties): (v1, c1) best, (v2, c2) best of the rest
#define TOP2_INIT(v, c) v1 = (v); c1 = (c); v2 = INT_MIN; c2 = 0u;
#define TOP2_PUSH(v, c)                    \
  if ((v) > v1) { v2 = v1; c2 = c1; v1 = (v); c1 = (c); } \
  else if ((v) > v2) { v2 = (v); c2 = (c); }

template <bool LOCAL, int BLOCK>
__device__ int ring_cs_dp(const FullParams &P, const FullTask &T, int slot, int32_t *sm, const uint32_t *genome,
                          const uint32_t *read, const Rect &rect, int &ret_i, int &ret_j, int &ret_k, int end_sc[3],
                          unsigned long long &cells) {
  const int W = P.W;
#define SMR(st, s) sm[((st) * W + (s)) * BLOCK]
  const int lena = T.glen, lenb = T.rlen;
  const size_t NT = (size_t)P.NT;
  const int ao = P.a_open, ae = P.a_ext, bo = P.b_open, be = P.b_ext;
  const bool revcmpl = T.gen_st && P.Tflag;
  unsigned long long *bp = P.bp64 + slot;
  int score = 0, max_i = 0, max_j = 0, max_k = 0;
  int letter[4];
#pragma unroll
  for (int k = 0; k < 4; k++) letter[k] = (k + T.initbp) % 4;
  int pxmin = 0, pxmax = -1;
  const int xp = P.xover;  // FASTA reads: global crossover penalty
  for (int i = 0; i < lenb; i++) {
    int x_min, x_max;
    rect_x_range(rect, lena, i, x_min, x_max);
    const bool nt = i < lenb - P.indel_taboo_len;
    const int colour = (int)extract4(read, (uint64_t)i);
    int qk[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (colour == 15) {
        qk[k] = 15;
        letter[k] = (k + T.initbp) % 4;
      } else {
        qk[k] = cstols_r(letter[k], colour);
        letter[k] = qk[k];
      }
    }
    cells += (unsigned long long)(x_max - x_min + 1);
    int d[12], l[12];
    int ini[12];  // initial cell of this mode (edge cells, cells right of the band)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int add = k == 0 ? 0 : xp;
      ini[3 * k + 0] = LOCAL ? -bo + add : NEG_HALF;
      ini[3 * k + 1] = LOCAL ? -ao + add : NEG_HALF;
      ini[3 * k + 2] = LOCAL ? add : NEG_HALF;
    }
    if (i == 0) {  // row -1: local-style init with the global crossover penalty (sw-full-cs.c:268-270)
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int add = k == 0 ? 0 : P.xover;
        d[3 * k + 0] = -bo + add; d[3 * k + 1] = -ao + add; d[3 * k + 2] = add;
      }
    } else if (x_min - 1 > pxmax) {
#pragma unroll
      for (int s = 0; s < 12; s++) d[s] = ini[s];
    } else {
      const int sp = x_min - pxmin;
#pragma unroll
      for (int s = 0; s < 12; s++) d[s] = SMR(s, sp);
    }
#pragma unroll
    for (int s = 0; s < 12; s++) {
      l[s] = ini[s];
      SMR(s, 0) = ini[s];
    }
    bp[((size_t)i * W) * NT] = 0ull;
    const int delta = x_min - pxmin;
    for (int j = x_min; j <= x_max; j++) {
      const int sidx = j - x_min + 1;
      int u[12], v[12];
      if (i == 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int add = k == 0 ? 0 : P.xover;
          u[3 * k + 0] = -bo + add; u[3 * k + 1] = -ao + add; u[3 * k + 2] = add;
        }
      } else if (j > pxmax) {
#pragma unroll
        for (int s = 0; s < 12; s++) u[s] = ini[s];
      } else {
        const int sp = sidx + delta;
#pragma unroll
        for (int s = 0; s < 12; s++) u[s] = SMR(s, sp);
      }
      const int dbj = (int)extract4(genome, (uint64_t)T.goff_global + (uint64_t)j);
      // ---- per-layer first-max of the three diagonal sources and of the two north sources --------
      int mv[4], nv[4];
      uint32_t mc[4], nc[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int dn = d[3 * q + 0], dw = d[3 * q + 1], dnw = d[3 * q + 2];
        if (!revcmpl) {
          mv[q] = dnw; mc[q] = CSC(q, D_NW_NW);
          if (nt && dn > mv[q]) { mv[q] = dn; mc[q] = CSC(q, D_NW_N); }
          if (dw > mv[q]) { mv[q] = dw; mc[q] = CSC(q, D_NW_W); }
        } else {
          mv[q] = dw; mc[q] = CSC(q, D_NW_W);
          if (nt && dn > mv[q]) { mv[q] = dn; mc[q] = CSC(q, D_NW_N); }
          if (dnw > mv[q]) { mv[q] = dnw; mc[q] = CSC(q, D_NW_NW); }
        }
        const int A = u[3 * q + 2] - bo - be, B = u[3 * q + 0] - be;
        if (!revcmpl) {
          if (nt) {
            nv[q] = A; nc[q] = CSC(q, D_N_NW);
            if (B > nv[q]) { nv[q] = B; nc[q] = CSC(q, D_N_N); }
          } else {
            nv[q] = B; nc[q] = CSC(q, D_N_N);
          }
        } else {
          nv[q] = B; nc[q] = CSC(q, D_N_N);
          if (nt && A > nv[q]) { nv[q] = A; nc[q] = CSC(q, D_N_NW); }
        }
      }
      int v1, v2, w1, w2;
      uint32_t c1, c2, e1, e2;
      TOP2_INIT(mv[0], mc[0]);
      TOP2_PUSH(mv[1], mc[1]);
      TOP2_PUSH(mv[2], mc[2]);
      TOP2_PUSH(mv[3], mc[3]);
      {
        const int a1 = v1, a2 = v2;
        const uint32_t b1 = c1, b2 = c2;
        TOP2_INIT(nv[0], nc[0]);
        TOP2_PUSH(nv[1], nc[1]);
        TOP2_PUSH(nv[2], nc[2]);
        TOP2_PUSH(nv[3], nc[3]);
        w1 = v1; w2 = v2; e1 = c1; e2 = c2;
        v1 = a1; v2 = a2; c1 = b1; c2 = b2;
      }
      unsigned long long bits = 0ull;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int resetval = k != 0 ? xp : 0;
        int ms, tmp;
        uint32_t t2;
        if (dbj == 15 || qk[k] == 15) ms = 0;  // N scores 0 (sw-full-cs.c:358-361)
        else ms = (dbj == qk[k]) ? P.match : P.mismatch;
        // northwest (:362-437): same layer first, then the best other layer + crossover
        {
          const bool own = (int)(c1 & 3u) == k;
          const int ov = own ? v2 : v1;
          const uint32_t oc = own ? c2 : c1;
          tmp = mv[k] + ms; t2 = mc[k];
          if (ov + ms + xp > tmp) { tmp = ov + ms + xp; t2 = oc; }
          if (LOCAL && tmp <= resetval) { tmp = resetval; t2 = 0; }
          v[3 * k + 2] = tmp;
          bits |= (unsigned long long)t2 << (5 * (3 * k + 2));
        }
        // north (:447-501)
        {
          const bool own = (int)(e1 & 3u) == k;
          const int ov = own ? w2 : w1;
          const uint32_t oc = own ? e2 : e1;
          tmp = nv[k]; t2 = nc[k];
          if (ov + xp > tmp) { tmp = ov + xp; t2 = oc; }
          if (LOCAL && tmp <= resetval) { tmp = resetval; t2 = 0; }
          v[3 * k + 0] = tmp;
          bits |= (unsigned long long)t2 << (5 * (3 * k + 0));
        }
        // west (:511-545), same layer only
        {
          const int lw = l[3 * k + 1], lnw = l[3 * k + 2];
          if (!revcmpl) {
            tmp = lnw - ao - ae; t2 = CSC(k, D_W_NW);
            if (!nt || lw - ae > tmp) { tmp = lw - ae; t2 = CSC(k, D_W_W); }
          } else {
            tmp = lw - ae; t2 = CSC(k, D_W_W);
            if (nt && lnw - ao - ae > tmp) { tmp = lnw - ao - ae; t2 = CSC(k, D_W_NW); }
          }
          if (LOCAL && tmp <= resetval) { tmp = resetval; t2 = 0; }
          v[3 * k + 1] = tmp;
          bits |= (unsigned long long)t2 << (5 * (3 * k + 1));
        }
        if (LOCAL || i == lenb - 1) {  // max score (:552-580)
          const int vn = v[3 * k + 0], vw = v[3 * k + 1], vnw = v[3 * k + 2];
          bool upd = false;
          if (!revcmpl) {
            if (vnw > score) { score = vnw; upd = true; }
            if (vn > score) { score = vn; upd = true; }
            if (vw > score) { score = vw; upd = true; }
          } else {
            if (vw > score) { score = vw; upd = true; }
            if (vn > score) { score = vn; upd = true; }
            if (vnw > score) { score = vnw; upd = true; }
          }
          if (upd) {
            max_i = i; max_j = j; max_k = k;
            end_sc[0] = vn; end_sc[1] = vw; end_sc[2] = vnw;
          }
        }
      }
      bp[((size_t)i * W + sidx) * NT] = bits;
#pragma unroll
      for (int s = 0; s < 12; s++) {
        d[s] = u[s];
        l[s] = v[s];
        SMR(s, sidx) = v[s];
      }
    }
    pxmin = x_min;
    pxmax = x_max;
  }
#undef SMR
  ret_i = max_i;
  ret_j = max_j;
  ret_k = max_k;
  return score;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) sw_full_cs_ring_kernel(const FullParams P) {
  extern __shared__ int32_t ring_smem[];
  const int slot = blockIdx.x * BLOCK + threadIdx.x;
  if (slot >= P.n_tasks) return;
  const int t = P.perm ? P.perm[slot] : slot;
  const FullTask T = P.tasks[t];
  FullResult R;
  memset(&R, 0, sizeof(R));
  if (!T.run) {
    P.results[t] = R;
    return;
  }
  int32_t *sm = ring_smem + threadIdx.x;
  const uint32_t *genome = T.gen_st ? P.genome_rc : P.genome_fwd;
  const uint32_t *read = P.reads + (size_t)T.ridx * P.stride;
  unsigned long long cells = 0;
  int ei = 0, ej = 0, ek = 0, esc[3] = {0, 0, 0};
  const Rect rect = task_rect(T, P.anchor_width, P.match, true);
  const int score = P.local ? ring_cs_dp<true, BLOCK>(P, T, slot, sm, genome, read, rect, ei, ej, ek, esc, cells)
                            : ring_cs_dp<false, BLOCK>(P, T, slot, sm, genome, read, rect, ei, ej, ek, esc, cells);
  if (cells) atomicAdd(P.cells, cells);
  if (!(score >= 0 && score >= T.thresh)) {  // sw_full_cs :1216-1226
    P.results[t] = R;
    return;
  }
  R.score = score;
  const int W = P.W;
  const size_t NT = (size_t)P.NT;
  const unsigned long long *bp = P.bp64 + slot;
  uint8_t *ops = P.ops + (size_t)t * (size_t)(P.max_glen + P.max_rlen);
  auto layer_letter = [&](int k, int i) -> int {
    int letter = (k + T.initbp) % 4, out = 15;
    for (int q = 0; q <= i; q++) {
      const int colour = (int)extract4(read, (uint64_t)q);
      if (colour == 15) {
        out = 15;
        letter = (k + T.initbp) % 4;
      } else {
        out = cstols_r(letter, colour);
        letter = out;
      }
    }
    return out;
  };
  auto back_of = [&](int ci, int cj, int k, int state) -> int {  // state: 0 north, 1 west, 2 northwest
    if (ci < 0 || cj < 0) return 0;
    int xmn, xmx;
    rect_x_range(rect, T.glen, ci, xmn, xmx);
    const int s = cj - xmn + 1;
    if (s <= 0 || cj > xmx) return 0;
    return (int)((bp[((size_t)ci * W + s) * NT] >> (5 * (3 * k + state))) & 31ull);
  };
  int i = ei, j = ej, k = ek;
  int state = 2, fromscore = esc[2];  // do_backtrace :643-652
  if (esc[1] > fromscore) { state = 1; fromscore = esc[1]; }
  if (esc[0] > fromscore) state = 0;
  int from = back_of(i, j, k, state);
  int off = (T.glen + T.rlen) - 1;
  int read_start = 0, genome_start = 0;
  if (from != 0) {
    while (i >= 0 && j >= 0) {
      const int dir = from >> 2, lay = from & 3;
      uint8_t op;
      if (dir == D_N_N || dir == D_N_NW) {
        R.deletions++;
        read_start = i--;
        op = (uint8_t)(2 | (k << 4));
      } else if (dir == D_W_W || dir == D_W_NW) {
        R.insertions++;
        genome_start = j--;
        op = 1;
      } else {
        const int dbj = (int)extract4(genome, (uint64_t)T.goff_global + (uint64_t)j);
        const int q = layer_letter(k, i);
        if (dbj == q || dbj == 15 || q == 15) R.matches++;
        else R.mismatches++;
        read_start = i--;
        genome_start = j--;
        op = (uint8_t)(3 | (k << 4));
      }
      if (k != lay) {
        op |= 4;
        R.crossovers++;
        k = lay;
      }
      ops[off] = op;
      const int nstate = (dir == D_N_N || dir == D_NW_N) ? 0 : (dir == D_W_W || dir == D_NW_W) ? 1 : 2;
      from = back_of(i, j, k, nstate);
      off--;
      if (from == 0) break;
    }
  }
  off++;
  if (k != 0 && off < T.glen + T.rlen) {  // :931-934
    ops[off] |= 4;
    R.crossovers++;
  }
  R.read_start = read_start;
  R.gmapped = ej - genome_start + 1;
  R.genome_start = genome_start + (int)T.goff_contig;
  R.rmapped = ei - read_start + 1;
  R.ops_start = off;
  R.ops_len = (T.glen + T.rlen) - off;
  P.results[t] = R;
}

// ------------------------------------------------------------------------------------------------
// colour space, four lanes per alignment (one per layer)
// ------------------------------------------------------------------------------------------------
// The thread-per-alignment colour kernel above needs 12 ring rows per thread, which leaves 4 warps per SM and
// every one of its ~200 instructions per cell exposed to ALU latency.  Here lane k of a quad owns layer k:
// 3 ring rows per lane (4x the resident warps for the same shared memory), a third of the arithmetic per
// lane, and the cross-layer candidates (northwest and north moves may cross over) exchanged inside the quad
// with shuffles of (value << 7 | layer priority | back-pointer code).  The "minus infinity" of the global-mode edge cells is -2^22
// here instead of -INT_MAX/2 so that the packed form fits 32 bits: every value derived from it carries
// exactly one such term, so all comparisons -- and with them every back-pointer -- come out as in the
// reference, and such cells can never hold the winning score.
#define NEG_Q (-(1 << 22))
#define QUAD_THREADS 128

__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// The eight quads of a warp step through rows and band offsets in LOCKSTEP (trip counts are the warp maxima, a
// quad outside its own band is predicated off), so the cross-layer exchange can use full-mask shuffles -- a
// shuffle under a computed sub-warp mask costs a WARPSYNC/collective sequence of ~10 instructions each.
template <bool LOCAL, bool REV, bool TABOO>
__device__ int quad_cs_dp(const FullParams &P, const FullTask &T, bool run, int slot, int k, int qbase, int32_t *sm,
                          const uint32_t *genome, const uint32_t *read, const Rect &rect, int &ret_i, int &ret_j,
                          int end_sc[3], unsigned long long &cells) {
  const int W = P.W;
  // ring layout [band offset][state][thread]: the three states of a cell sit at compile-time offsets of one pointer
  constexpr int sstride = QUAD_THREADS, slotstride = 3 * QUAD_THREADS;
  const int lena = T.glen, lenb = run ? T.rlen : 0;
  const size_t bstride = (size_t)P.NT * 4;  // u16 elements between two band offsets
  const int ao = P.a_open, ae = P.a_ext, bo = P.b_open, be = P.b_ext;
  constexpr bool revcmpl = REV;   // T.gen_st && P.Tflag of every task of the launch (the tasks are grouped by it)
  unsigned short *bp_task = (unsigned short *)P.bp64 + (size_t)slot * 4 + k;
  int score = 0, max_i = 0, max_j = 0;
  int letter = (k + T.initbp) % 4;
  int pxmin = 0, pxmax = -1;
  // crossover penalty: global, or per read position when the read came with qualities (sw-full-cs.c:312)
  const int16_t *xrow = P.xover_pos ? P.xover_pos + (size_t)(T.ridx >> 1) * (size_t)P.xover_stride : nullptr;
  int add_prev = k == 0 ? 0 : P.xover;   // what row -1 was initialised with: the global penalty (:268-270)
  const int match = P.match, mismatch = P.mismatch;
  const int kprio = (3 - k) << 5;   // tie order of the cross-layer maximum: lowest layer first
  const int nrows_w = warp_max_i(lenb);
  for (int i = 0; i < nrows_w; i++) {
    const bool row_on = i < lenb;
    int x_min = 0, x_max = -1;
    if (row_on) rect_x_range(rect, lena, i, x_min, x_max);
    const int width = row_on ? x_max - x_min + 1 : 0;
    const int wmax = warp_max_i(width);
    const bool nt = TABOO ? i < lenb - P.indel_taboo_len : true;   // rows past lenb are switched off anyway
    int qk = 15;
    if (row_on) {
      const int colour = (int)extract4(read, (uint64_t)i);
      if (colour == 15) {
        letter = (k + T.initbp) % 4;
      } else {
        qk = cstols_r(letter, colour);
        letter = qk;
      }
      if (k == 0) cells += (unsigned long long)width;
    }
    const int m_eq = qk == 15 ? 0 : match, m_ne = qk == 15 ? 0 : mismatch;
    const int xp = (xrow && row_on) ? (int)xrow[i] : P.xover;
    const int add = k == 0 ? 0 : xp;
    const int ini_n = LOCAL ? -bo + add : NEG_Q, ini_w = LOCAL ? -ao + add : NEG_Q, ini_nw = LOCAL ? add : NEG_Q;
    const int resetval = k != 0 ? xp : 0;
    // what a cell of the previous row reads as when it lies right of that row's band: the initial cell written
    // with that row's penalty (:604-612), or row -1 itself (local-style init with the global crossover penalty)
    const int r_n = (i == 0 || LOCAL) ? -bo + add_prev : NEG_Q, r_w = (i == 0 || LOCAL) ? -ao + add_prev : NEG_Q,
              r_nw = (i == 0 || LOCAL) ? add_prev : NEG_Q;
    int d_n = r_n, d_w = r_w, d_nw = r_nw;
    const int delta = x_min - pxmin;
    if (row_on && i > 0 && x_min - 1 <= pxmax) {
      const int32_t *p = sm + delta * slotstride;
      d_n = p[0]; d_w = p[sstride]; d_nw = p[2 * sstride];
    }
    unsigned short *bp = bp_task + (size_t)i * W * bstride;
    if (row_on) {
      sm[0] = ini_n; sm[sstride] = ini_w; sm[2 * sstride] = ini_nw;  // edge cell (i, x_min-1)
      *bp = 0;
    }
    int l_w = ini_w, l_nw = ini_nw;
    int32_t *cur = sm;                                 // ring slot of offset s
    const int32_t *prev = sm + delta * slotstride;   // ring slot of the previous row's cell in the same column
    uint32_t gpos = T.goff_global + (uint32_t)x_min;
    uint32_t gword = genome[gpos >> 3];
#pragma unroll 4
    for (int s = 1; s <= wmax; s++) {
      const bool on = s <= width;
      cur += slotstride;
      prev += slotstride;
      bp += bstride;
      const int j = x_min + s - 1;
      int u_n = r_n, u_w = r_w, u_nw = r_nw;
      if (on && j <= pxmax) { u_n = prev[0]; u_w = prev[sstride]; u_nw = prev[2 * sstride]; }
      if (on && (gpos & 7u) == 0u) gword = genome[gpos >> 3];
      const int dbj = (int)((gword >> (4u * (gpos & 7u))) & 15u);
      gpos++;
      const int ms = dbj == qk ? m_eq : (dbj == 15 ? 0 : m_ne);   // N on either side scores 0
      // own layer: first-max of the diagonal sources and of the north sources; md / nd hold the back-pointer code
      // of the choice (direction << 2 | own layer, CSC)
      int mv, md, nv, nd;
      if (!revcmpl) {
        mv = d_nw; md = CSC(k, D_NW_NW);
        if (nt && d_n > mv) { mv = d_n; md = CSC(k, D_NW_N); }
        if (d_w > mv) { mv = d_w; md = CSC(k, D_NW_W); }
      } else {
        mv = d_w; md = CSC(k, D_NW_W);
        if (nt && d_n > mv) { mv = d_n; md = CSC(k, D_NW_N); }
        if (d_nw > mv) { mv = d_nw; md = CSC(k, D_NW_NW); }
      }
      {
        const int A = u_nw - bo - be, B = u_n - be;
        if (!revcmpl) {
          if (nt) {
            nv = A; nd = CSC(k, D_N_NW);
            if (B > nv) { nv = B; nd = CSC(k, D_N_N); }
          } else {
            nv = B; nd = CSC(k, D_N_N);
          }
        } else {
          nv = B; nd = CSC(k, D_N_N);
          if (nt && A > nv) { nv = A; nd = CSC(k, D_N_NW); }
        }
      }
      // best other layer: the reference tries layers in ascending order and replaces on strict > (sw-full-cs.c:
      // 375-437, :460-501), i.e. maximum value, ties to the lowest layer -- a plain integer max over the keys
      // (value << 7 | (3 - layer) << 5 | back-pointer code) of the three other lanes: the partner lane's key and
      // the maximum of the other pair, two xor-shuffles per key
      const int pm = (mv << 7) | kprio | md, pn = (nv << 7) | kprio | nd;
      const int xm = __shfl_xor_sync(0xffffffffu, pm, 1), xn = __shfl_xor_sync(0xffffffffu, pn, 1);
      const int ym = __shfl_xor_sync(0xffffffffu, max(pm, xm), 2), yn = __shfl_xor_sync(0xffffffffu, max(pn, xn), 2);
      const int km = max(xm, ym), kn = max(xn, yn);
      const int ov = km >> 7, ow = kn >> 7;
      const int oc = km & 31, oe = kn & 31;
      int tmp, v_n, v_w, v_nw;
      uint32_t t2, bits;
      tmp = mv + ms; t2 = (uint32_t)md;
      if (ov + ms + xp > tmp) { tmp = ov + ms + xp; t2 = (uint32_t)oc; }
      if (LOCAL && tmp <= resetval) { tmp = resetval; t2 = 0; }
      v_nw = tmp; bits = t2 << 10;
      tmp = nv; t2 = (uint32_t)nd;
      if (ow + xp > tmp) { tmp = ow + xp; t2 = (uint32_t)oe; }
      if (LOCAL && tmp <= resetval) { tmp = resetval; t2 = 0; }
      v_n = tmp; bits |= t2;
      if (!revcmpl) {  // west (:511-545), same layer only
        tmp = l_nw - ao - ae; t2 = CSC(k, D_W_NW);
        if (!nt || l_w - ae > tmp) { tmp = l_w - ae; t2 = CSC(k, D_W_W); }
      } else {
        tmp = l_w - ae; t2 = CSC(k, D_W_W);
        if (nt && l_nw - ao - ae > tmp) { tmp = l_nw - ao - ae; t2 = CSC(k, D_W_NW); }
      }
      if (LOCAL && tmp <= resetval) { tmp = resetval; t2 = 0; }
      v_w = tmp; bits |= t2 << 5;
      d_n = u_n; d_w = u_w; d_nw = u_nw;
      l_w = v_w; l_nw = v_nw;
      if (on) {
        *bp = (unsigned short)bits;
        cur[0] = v_n; cur[sstride] = v_w; cur[2 * sstride] = v_nw;
        if (LOCAL || i == lenb - 1) {  // per-lane first maximum in traversal order; the quad combines at the end
          int best = v_nw > v_n ? v_nw : v_n;
          best = best > v_w ? best : v_w;
          if (best > score) {
            score = best; max_i = i; max_j = j;
            end_sc[0] = v_n; end_sc[1] = v_w; end_sc[2] = v_nw;
          }
        }
      }
    }
    if (row_on) {
      pxmin = x_min;
      pxmax = x_max;
      add_prev = add;
    }
  }
  ret_i = max_i;
  ret_j = max_j;
  return score;
}

template <bool REV, bool TABOO>
__global__ void __launch_bounds__(QUAD_THREADS) sw_full_cs_quad_kernel(const FullParams P) {
  extern __shared__ int32_t ring_smem[];
  const int k = threadIdx.x & 3, qbase = (threadIdx.x & 31) & ~3;
  int slot = blockIdx.x * (QUAD_THREADS / 4) + (threadIdx.x >> 2);
  const bool live = slot < P.n_tasks;
  if (!live) slot = P.n_tasks - 1;   // dead quads shadow the last task and stay in the warp's lockstep, switched off
  const int t = P.perm ? P.perm[slot] : slot;
  const FullTask T = P.tasks[t];
  FullResult R;
  memset(&R, 0, sizeof(R));
  const bool run = live && T.run;
  int32_t *sm = ring_smem + threadIdx.x;
  const uint32_t *genome = T.gen_st ? P.genome_rc : P.genome_fwd;
  const uint32_t *read = P.reads + (size_t)T.ridx * P.stride;
  unsigned long long cells = 0;
  int ei = 0, ej = 0, esc[3] = {0, 0, 0};
  const Rect rect = task_rect(T, P.anchor_width, P.match, true);
  int score = P.local ? quad_cs_dp<true, REV, TABOO>(P, T, run, slot, k, qbase, sm, genome, read, rect, ei, ej, esc, cells)
                      : quad_cs_dp<false, REV, TABOO>(P, T, run, slot, k, qbase, sm, genome, read, rect, ei, ej, esc, cells);
  // the reference scans cells in (i, j, layer) order and keeps the first maximum (sw-full-cs.c:552-580):
  // every lane gathers the four per-layer candidates and picks the winner
  int ek = 0;
  int w_score = 0, w_i = 0, w_j = 0, w_e0 = 0, w_e1 = 0, w_e2 = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int os = __shfl_sync(0xffffffffu, score, qbase + q);
    const int oi = __shfl_sync(0xffffffffu, ei, qbase + q);
    const int oj = __shfl_sync(0xffffffffu, ej, qbase + q);
    const int o0 = __shfl_sync(0xffffffffu, esc[0], qbase + q);
    const int o1 = __shfl_sync(0xffffffffu, esc[1], qbase + q);
    const int o2 = __shfl_sync(0xffffffffu, esc[2], qbase + q);
    const bool better = os > w_score || (os == w_score && os > 0 && (oi < w_i || (oi == w_i && oj < w_j)));
    if (better) { w_score = os; w_i = oi; w_j = oj; ek = q; w_e0 = o0; w_e1 = o1; w_e2 = o2; }
  }
  if (k != 0 || !live) return;
  if (!T.run) {
    P.results[t] = R;
    return;
  }
  score = w_score; ei = w_i; ej = w_j; esc[0] = w_e0; esc[1] = w_e1; esc[2] = w_e2;
  if (cells) atomicAdd(P.cells, cells);
  if (!(score >= 0 && score >= T.thresh)) {  // sw_full_cs :1216-1226
    P.results[t] = R;
    return;
  }
  R.score = score;
  const int W = P.W;
  const size_t NT = (size_t)P.NT;
  const unsigned short *bp = (const unsigned short *)P.bp64 + (size_t)slot * 4;
  uint8_t *ops = P.ops + (size_t)t * (size_t)(P.max_glen + P.max_rlen);
  // Letter of layer kk at read row i: the layer's start letter XORed with the colours since the last N (cstols,
  // util.h:157-180, is XOR on the 2-bit codes; a colour N restarts the layer, sw-full-cs.c:1181-1196).  The
  // traceback asks for rows in descending order, so the XOR is carried from row to row (one colour per step) and
  // only recomputed from the start of the read at the first use and after crossing an N.
  int lx_row = -1, lx_val = 0;
  auto layer_letter = [&](int kk, int i) -> int {
    if ((int)extract4(read, (uint64_t)i) == 15) return 15;
    while (lx_row > i) {   // lx_val is the XOR at lx_row (colour there is not N)
      lx_val ^= (int)extract4(read, (uint64_t)lx_row);
      lx_row--;
      if ((int)extract4(read, (uint64_t)lx_row) == 15) lx_row = -1;
    }
    if (lx_row != i) {
      lx_val = 0;
      for (int q = 0; q <= i; q++) {
        const int colour = (int)extract4(read, (uint64_t)q);
        lx_val = colour == 15 ? 0 : lx_val ^ colour;
      }
      lx_row = i;
    }
    return ((kk + T.initbp) & 3) ^ lx_val;
  };
  auto back_of = [&](int ci, int cj, int kk, int state) -> int {  // state: 0 north, 1 west, 2 northwest
    if (ci < 0 || cj < 0) return 0;
    int xmn, xmx;
    rect_x_range(rect, T.glen, ci, xmn, xmx);
    const int s = cj - xmn + 1;
    if (s <= 0 || cj > xmx) return 0;
    return (int)((bp[((size_t)ci * W + s) * NT * 4 + kk] >> (5 * state)) & 31u);
  };
  int i = ei, j = ej, kk = ek;
  int state = 2, fromscore = esc[2];  // do_backtrace :643-652
  if (esc[1] > fromscore) { state = 1; fromscore = esc[1]; }
  if (esc[0] > fromscore) state = 0;
  int from = back_of(i, j, kk, state);
  int off = (T.glen + T.rlen) - 1;
  int read_start = 0, genome_start = 0;
  if (from != 0) {
    while (i >= 0 && j >= 0) {
      const int dir = from >> 2, lay = from & 3;
      uint8_t op;
      if (dir == D_N_N || dir == D_N_NW) {
        R.deletions++;
        read_start = i--;
        op = (uint8_t)(2 | (kk << 4));
      } else if (dir == D_W_W || dir == D_W_NW) {
        R.insertions++;
        genome_start = j--;
        op = 1;
      } else {
        const int dbj = (int)extract4(genome, (uint64_t)T.goff_global + (uint64_t)j);
        const int q = layer_letter(kk, i);
        if (dbj == q || dbj == 15 || q == 15) R.matches++;
        else R.mismatches++;
        read_start = i--;
        genome_start = j--;
        op = (uint8_t)(3 | (kk << 4));
      }
      if (kk != lay) {
        op |= 4;
        R.crossovers++;
        kk = lay;
      }
      ops[off] = op;
      const int nstate = (dir == D_N_N || dir == D_NW_N) ? 0 : (dir == D_W_W || dir == D_NW_W) ? 1 : 2;
      from = back_of(i, j, kk, nstate);
      off--;
      if (from == 0) break;
    }
  }
  off++;
  if (kk != 0 && off < T.glen + T.rlen) {  // :931-934
    ops[off] |= 4;
    R.crossovers++;
  }
  R.read_start = read_start;
  R.gmapped = ej - genome_start + 1;
  R.genome_start = genome_start + (int)T.goff_contig;
  R.rmapped = ei - read_start + 1;
  R.ops_start = off;
  R.ops_len = (T.glen + T.rlen) - off;
  P.results[t] = R;
}

// Shared memory per thread of one ring launch, and the block size chosen for it.
// (colour space runs four lanes per alignment, each with the 3 rows of its layer: 3 rows per THREAD either way)
size_t ring_smem_per_thread(bool cs, int W) { return (size_t)3 * (size_t)W * 4; }
int ring_block_threads(bool cs, int W) {
  if (cs) return QUAD_THREADS;
  const size_t per = ring_smem_per_thread(cs, W);
  return per * 64 <= 48 * 1024 ? 64 : 32;
}
bool ring_fits(bool cs, int W) { return ring_smem_per_thread(cs, W) * (cs ? QUAD_THREADS : 32) <= 200 * 1024; }

int launch_sw_full_ring(shrimp_gpu_ctx *ctx, const FullParams &P, bool cs) {
  if (P.n_tasks <= 0) return SHRIMP_OK;
  const int block = ring_block_threads(cs, P.W);
  const size_t smem = ring_smem_per_thread(cs, P.W) * block;
  const int per_block = cs ? block / 4 : block;
  const int grid = (P.n_tasks + per_block - 1) / per_block;
#define RING_LAUNCH(K)                                                                                   \
  do {                                                                                                   \
    SH_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));    \
    K<<<grid, block, smem, ctx->stream>>>(P);                                                            \
  } while (0)
  if (cs) {
    const bool taboo = P.indel_taboo_len != 0;
    if (P.rev) {
      if (taboo) RING_LAUNCH((sw_full_cs_quad_kernel<true, true>));
      else RING_LAUNCH((sw_full_cs_quad_kernel<true, false>));
    } else {
      if (taboo) RING_LAUNCH((sw_full_cs_quad_kernel<false, true>));
      else RING_LAUNCH((sw_full_cs_quad_kernel<false, false>));
    }
  } else {
    if (P.rev) {
      if (block == 64) RING_LAUNCH((sw_full_ls_ring_kernel<64, true>));
      else RING_LAUNCH((sw_full_ls_ring_kernel<32, true>));
    } else {
      if (block == 64) RING_LAUNCH((sw_full_ls_ring_kernel<64, false>));
      else RING_LAUNCH((sw_full_ls_ring_kernel<32, false>));
    }
  }
#undef RING_LAUNCH
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_FULL);
  return SHRIMP_OK;
}

}  // namespace shrimp
