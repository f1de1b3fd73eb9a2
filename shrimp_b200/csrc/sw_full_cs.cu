// Full Smith-Waterman with traceback, colour space.
//
// Replaces common/sw-full-cs.c: full_sw :249-623 -- the letter genome against the four letter-space
// translations ("layers") of a colour read, 3 states per layer, crossover transitions between
// layers on northwest and north moves (never on a genome-consuming gap, :535-538), per-colour
// crossover penalty, indel taboo zone -- do_backtrace :633-937 and the coordinates of sw_full_cs
// :1146-1236.  Same work mapping as sw_full.cu (thread per alignment, scratch [cell][task]):
//   row_cs[12][glen+1][NT]   rolling DP row: state s = 3*layer + {0 north, 1 west, 2 northwest}
//   bp_cs[rlen*glen*12][NT]  one back-pointer byte per state: (direction << 2) | source layer
// Edit script bytes: bits 0-1 = 1 insertion (genome base vs '-') / 2 deletion (read base vs '-') /
// 3 match-mismatch, bit 2 = crossover on this column, bits 4-5 = layer whose translation is printed.
#include "band.cuh"

namespace shrimp {

enum { D_N_N = 1, D_N_NW = 2, D_W_NW = 3, D_W_W = 4, D_NW_N = 5, D_NW_NW = 6, D_NW_W = 7 };  // sw-full-cs.c:43-49
#define CSF(layer, dir) ((uint8_t)(((dir) << 2) | (layer)))

// cstols (util.h:157-180)
__device__ __forceinline__ int cstols_dev(int first_letter, int colour) {
  if (first_letter == 15 || colour < 0 || colour > 3) return 15;
  return (first_letter % 2 == 0) ? (4 + first_letter + colour) % 4 : (4 + first_letter - colour) % 4;
}

template <bool LOCAL>
__device__ int full_sw_cs_dev(const FullParams &P, const FullTask &T, int t, const uint32_t *genome,
                              const uint32_t *read, int &ret_i, int &ret_j, int &ret_k, int end_sc[3],
                              unsigned long long &cells) {
  const int lena = T.glen, lenb = T.rlen, NT = P.NT;
  const int ao = P.a_open, ae = P.a_ext, bo = P.b_open, be = P.b_ext;
  const bool revcmpl = T.gen_st && P.Tflag;
  const Rect rect = task_rect(T, P.anchor_width, P.match, true);
  const size_t plane = (size_t)(P.max_glen + 1) * NT;
  int32_t *row = P.row_cs + t;  // state s at row[s*plane + c*NT]
  uint8_t *bp = P.bp_cs + t;
  for (int c = 0; c <= lena; c++) {  // row -1: local-style init with the global crossover penalty (:268-270)
    for (int k = 0; k < 4; k++) {
      const int add = k == 0 ? 0 : P.xover;
      row[(3 * k + 0) * plane + (size_t)c * NT] = -bo + add;
      row[(3 * k + 1) * plane + (size_t)c * NT] = -ao + add;
      row[(3 * k + 2) * plane + (size_t)c * NT] = add;
    }
  }
  int score = 0, max_i = 0, max_j = 0, max_k = 0;
  int letter[4];  // running letter of each layer (the four translations of the read, :1181-1196)
  for (int k = 0; k < 4; k++) letter[k] = (k + T.initbp) % 4;
  for (int i = 0; i < lenb; i++) {
    int x_min, x_max;
    rect_x_range(rect, lena, i, x_min, x_max);
    // global crossover penalty, or the read position's when the read came with qualities (sw-full-cs.c:312)
    const int xp = P.xover_pos ? (int)P.xover_pos[(size_t)(T.ridx >> 1) * (size_t)P.xover_stride + i] : P.xover;
    const bool nt = i < lenb - P.indel_taboo_len;
    const int colour = (int)extract4(read, (uint64_t)i);
    int qk[4];
    for (int k = 0; k < 4; k++) {
      if (colour == 15) {
        qk[k] = 15;
        letter[k] = (k + T.initbp) % 4;
      } else {
        qk[k] = cstols_dev(letter[k], colour);
        letter[k] = qk[k];
      }
    }
    cells += (unsigned long long)(x_max - x_min + 1);
    // left edge cell (i, x_min-1) at storage column x_min
    int d[12], l[12];
    {
      const size_t c0 = (size_t)x_min * NT;
      for (int s = 0; s < 12; s++) d[s] = row[s * plane + c0];
      for (int k = 0; k < 4; k++) {
        const int add = k == 0 ? 0 : xp;
        l[3 * k + 0] = LOCAL ? -bo + add : NEG_HALF;
        l[3 * k + 1] = LOCAL ? -ao + add : NEG_HALF;
        l[3 * k + 2] = LOCAL ? add : NEG_HALF;
      }
      for (int s = 0; s < 12; s++) row[s * plane + c0] = l[s];
      if (x_min >= 1)
        for (int s = 0; s < 12; s++) bp[(((size_t)i * lena + (x_min - 1)) * 12 + s) * NT] = 0;
    }
    for (int j = x_min; j <= x_max; j++) {
      const size_t c = (size_t)(j + 1) * NT;
      int u[12], v[12];
      uint8_t b[12];
      for (int s = 0; s < 12; s++) u[s] = row[s * plane + c];  // cell (i-1, j)
      const int dbj = (int)extract4(genome, (uint64_t)T.goff_global + (uint64_t)j);
      for (int k = 0; k < 4; k++) {
        const int resetval = k != 0 ? xp : 0;
        int ms, tmp;
        uint8_t t2;
        if (dbj == 15 || qk[k] == 15) ms = 0;  // N scores 0 (:358-361)
        else ms = (dbj == qk[k]) ? P.match : P.mismatch;
        const int dn = d[3 * k + 0], dw = d[3 * k + 1], dnw = d[3 * k + 2];
        // northwest (:362-437)
        if (!revcmpl) {
          tmp = dnw + ms; t2 = CSF(k, D_NW_NW);
          if (nt && dn + ms > tmp) { tmp = dn + ms; t2 = CSF(k, D_NW_N); }
          if (dw + ms > tmp) { tmp = dw + ms; t2 = CSF(k, D_NW_W); }
        } else {
          tmp = dw + ms; t2 = CSF(k, D_NW_W);
          if (nt && dn + ms > tmp) { tmp = dn + ms; t2 = CSF(k, D_NW_N); }
          if (dnw + ms > tmp) { tmp = dnw + ms; t2 = CSF(k, D_NW_NW); }
        }
        for (int q = 0; q < 4; q++) {
          if (q == k) continue;
          const int qn = d[3 * q + 0] + ms + xp, qw = d[3 * q + 1] + ms + xp, qnw = d[3 * q + 2] + ms + xp;
          if (!revcmpl) {
            if (qnw > tmp) { tmp = qnw; t2 = CSF(q, D_NW_NW); }
            if (nt && qn > tmp) { tmp = qn; t2 = CSF(q, D_NW_N); }
            if (qw > tmp) { tmp = qw; t2 = CSF(q, D_NW_W); }
          } else {
            if (qw > tmp) { tmp = qw; t2 = CSF(q, D_NW_W); }
            if (nt && qn > tmp) { tmp = qn; t2 = CSF(q, D_NW_N); }
            if (qnw > tmp) { tmp = qnw; t2 = CSF(q, D_NW_NW); }
          }
        }
        if (LOCAL && tmp <= resetval) { tmp = resetval; t2 = 0; }
        v[3 * k + 2] = tmp; b[3 * k + 2] = t2;
        // north (:447-501)
        const int un = u[3 * k + 0], unw = u[3 * k + 2];
        if (!revcmpl) {
          tmp = unw - bo - be; t2 = CSF(k, D_N_NW);
          if (!nt || un - be > tmp) { tmp = un - be; t2 = CSF(k, D_N_N); }
        } else {
          tmp = un - be; t2 = CSF(k, D_N_N);
          if (nt && unw - bo - be > tmp) { tmp = unw - bo - be; t2 = CSF(k, D_N_NW); }
        }
        for (int q = 0; q < 4; q++) {
          if (q == k) continue;
          const int qnw = u[3 * q + 2] - bo - be + xp, qn = u[3 * q + 0] - be + xp;
          if (!revcmpl) {
            if (nt && qnw > tmp) { tmp = qnw; t2 = CSF(q, D_N_NW); }
            if (qn > tmp) { tmp = qn; t2 = CSF(q, D_N_N); }
          } else {
            if (qn > tmp) { tmp = qn; t2 = CSF(q, D_N_N); }
            if (nt && qnw > tmp) { tmp = qnw; t2 = CSF(q, D_N_NW); }
          }
        }
        if (LOCAL && tmp <= resetval) { tmp = resetval; t2 = 0; }
        v[3 * k + 0] = tmp; b[3 * k + 0] = t2;
        // west (:511-545), same layer only
        const int lw = l[3 * k + 1], lnw = l[3 * k + 2];
        if (!revcmpl) {
          tmp = lnw - ao - ae; t2 = CSF(k, D_W_NW);
          if (!nt || lw - ae > tmp) { tmp = lw - ae; t2 = CSF(k, D_W_W); }
        } else {
          tmp = lw - ae; t2 = CSF(k, D_W_W);
          if (nt && lnw - ao - ae > tmp) { tmp = lnw - ao - ae; t2 = CSF(k, D_W_NW); }
        }
        if (LOCAL && tmp <= resetval) { tmp = resetval; t2 = 0; }
        v[3 * k + 1] = tmp; b[3 * k + 1] = t2;
        // max score (:552-580)
        if (LOCAL || i == lenb - 1) {
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
      for (int s = 0; s < 12; s++) {
        d[s] = u[s];
        l[s] = v[s];
        row[s * plane + c] = v[s];
        bp[(((size_t)i * lena + j) * 12 + s) * NT] = b[s];
      }
    }
    if (i + 1 < lenb) {  // cells right of the band read by the next row (:604-612), penalty of colour i
      int nmin, nmax;
      rect_x_range(rect, lena, i + 1, nmin, nmax);
      for (int j = x_max + 1; j <= nmax; j++) {
        const size_t c = (size_t)(j + 1) * NT;
        for (int k = 0; k < 4; k++) {
          const int add = k == 0 ? 0 : xp;
          row[(3 * k + 0) * plane + c] = LOCAL ? -bo + add : NEG_HALF;
          row[(3 * k + 1) * plane + c] = LOCAL ? -ao + add : NEG_HALF;
          row[(3 * k + 2) * plane + c] = LOCAL ? add : NEG_HALF;
        }
        for (int s = 0; s < 12; s++) bp[(((size_t)i * lena + j) * 12 + s) * NT] = 0;
      }
    }
  }
  ret_i = max_i;
  ret_j = max_j;
  ret_k = max_k;
  return score;
}

__global__ void __launch_bounds__(128) sw_full_cs_kernel(const FullParams P) {
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
  int ei = 0, ej = 0, ek = 0, esc[3] = {0, 0, 0};
  const int score = P.local ? full_sw_cs_dev<true>(P, T, slot, genome, read, ei, ej, ek, esc, cells)
                            : full_sw_cs_dev<false>(P, T, slot, genome, read, ei, ej, ek, esc, cells);
  if (cells) atomicAdd(P.cells, cells);
  if (!(score >= 0 && score >= T.thresh)) {  // sw_full_cs :1216-1226: below threshold -> score 0, no traceback
    P.results[t] = R;
    return;
  }
  R.score = score;
  // letters of the four layers are needed by the match count: recompute the translation on demand
  const int lena = T.glen, NT = P.NT;
  const uint8_t *bp = P.bp_cs + slot;
  uint8_t *ops = P.ops + (size_t)t * (size_t)(P.max_glen + P.max_rlen);
  // qr[k][i] for all i: walk the read once per layer into the ops scratch tail? -- reads are short;
  // recompute by scanning from the last N (or the start) up to i.
  auto layer_letter = [&](int k, int i) -> int {
    int letter = (k + T.initbp) % 4, out = 15;
    for (int q = 0; q <= i; q++) {
      const int colour = (int)extract4(read, (uint64_t)q);
      if (colour == 15) {
        out = 15;
        letter = (k + T.initbp) % 4;
      } else {
        out = cstols_dev(letter, colour);
        letter = out;
      }
    }
    return out;
  };
  auto back_of = [&](int ci, int cj, int k, int state) -> int {  // state: 0 north, 1 west, 2 northwest
    if (ci < 0 || cj < 0) return 0;
    return bp[(((size_t)ci * lena + cj) * 12 + 3 * k + state) * NT];
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

int launch_sw_full_cs(shrimp_gpu_ctx *ctx, const FullParams &P) {
  if (P.n_tasks <= 0) return SHRIMP_OK;
  sw_full_cs_kernel<<<(P.n_tasks + 127) / 128, 128, 0, ctx->stream>>>(P);
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_FULL);
  return SHRIMP_OK;
}

}  // namespace shrimp
