// Chunk pipeline: shrimp_gpu_map_reads = handle_read (gmapper/mapping.c:1773-1868) for a whole
// chunk of reads.  Device stages: read reverse-complement -> seed scan (scan.cu) -> sw_vector over
// all candidate windows (sw_vector.cu) -> pass-1 replay + top-k (pass1.cu) -> full SW with
// traceback (sw_full.cu).  Host stage (this file, plain C++ on the few <=30 hits per read that
// survive): hit_run_post_sw's double arithmetic, the pass-2 threshold, duplicate removal with
// qsort and the final ranking -- the parts SURVEY section 8 a21 keeps on the host because their
// tie order is glibc qsort's.
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <omp.h>
#include <atomic>
#include "glibc_tables.inc"
#include <cub/cub.cuh>
#include "chunk.cuh"

namespace shrimp {

// ---- entry points of the other translation units -----------------------------------------------
size_t scan_smem_bytes(int cap, int max_rl, int k_cap, int bm_log2, int warps, int stash);
size_t scan_cta_smem_bytes(int cap, int max_rl, int k_cap, int bm_log2, int n_part, int win, bool global_arrays);
int launch_scan_cta(shrimp_gpu_ctx *ctx, ScanParams &P, int n_ctas, int threads);
int launch_scan_replay(shrimp_gpu_ctx *ctx, ScanParams &P, uint32_t n_rec, int ks_cap);
int launch_post_sw(shrimp_gpu_ctx *ctx, const PostParams &P, DevBuf &scratch);

int launch_scan(shrimp_gpu_ctx *ctx, ScanParams &P, int warps_per_cta, int n_ctas);
int launch_build_vec_tasks(shrimp_gpu_ctx *ctx, const TaskBuildParams &P);
int launch_window_slots(shrimp_gpu_ctx *ctx, const uint32_t *genome, const uint32_t *goff, const int32_t *glen,
                        const uint32_t *out, uint32_t n_tasks, uint32_t *slot);
int launch_sw_gapless(shrimp_gpu_ctx *ctx, const GaplessParams &P);
int launch_pass1_replay(shrimp_gpu_ctx *ctx, const Pass1Params &P);
int launch_select_unpaired(shrimp_gpu_ctx *ctx, const Pass1Params &P);
int launch_sw_full_ls(shrimp_gpu_ctx *ctx, const FullParams &P);
int launch_sw_full_cs(shrimp_gpu_ctx *ctx, const FullParams &P);
int launch_sw_full_ring(shrimp_gpu_ctx *ctx, const FullParams &P, bool cs);
bool ring_fits(bool cs, int W);

__constant__ uint8_t c_cmpl_r[16] = {3, 2, 1, 0, 0, 10, 9, 7, 8, 6, 5, 14, 13, 12, 11, 15};

__device__ __forceinline__ void put4_dev(uint32_t *a, int i, uint32_t v) {
  uint32_t w = a[i >> 3];
  w &= ~(0xfu << (4 * (i & 7)));
  w |= (v & 0xfu) << (4 * (i & 7));
  a[i >> 3] = w;
}

// reads_all row 2r = read r as given, row 2r+1 = its reverse complement
// (reverse_complement_read_ls util.c:541-598; reverse_complement_read_cs util.c:601-618)
__global__ void revcomp_reads_kernel(const uint32_t *in, uint32_t *out, int stride, int n_reads, const int32_t *read_len,
                                     const int8_t *initbp, int colour_space, int rev_even, int rev_odd) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const uint32_t *src = in + (size_t)r * stride;
  uint32_t *f = out + (size_t)(2 * r) * stride, *rc = f + stride;
  if ((r & 1) ? rev_odd : rev_even) {  // read_reverse: the two strands trade places
    rc = f;
    f = rc + stride;
  }
  for (int w = 0; w < stride; w++) {
    f[w] = src[w];
    rc[w] = 0;
  }
  const int rl = read_len[r];
  if (rl <= 0) return;
  if (!colour_space) {
    for (int i = 0; i < rl; i++) put4_dev(rc, rl - 1 - i, c_cmpl_r[extract4(src, (uint64_t)i)]);
  } else {
    int base = initbp[r];
    for (int i = 0; i < rl; i++) {
      const int c = (int)extract4(src, (uint64_t)i);
      // cstols, util.h:157-180
      base = (base == 15 || c > 3) ? 15 : ((base % 2 == 0) ? (4 + base + c) % 4 : (4 + base - c) % 4);
      if (i >= 1) put4_dev(rc, rl - i, (uint32_t)c);
    }
    const int ci = c_cmpl_r[initbp[r] & 15];
    put4_dev(rc, 0, (base > 3 || ci > 3) ? 15u : (uint32_t)(base ^ ci));  // lstocs
  }
}

// one thread per (read, selected slot): hit_run_full_sw's orientation logic (mapping.c:353-361,
// reverse_hit :254-263, anchor_reverse anchors.h:30-34)
__global__ void build_full_tasks_kernel(const FullBuildParams P) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int NT = P.M.num_tmp_outputs;
  if (idx >= P.n_reads * NT) return;
  const int r = idx / NT, k = idx % NT;
  if (k >= P.n_sel[r]) return;
  FullTask T;
  SelInfo I;
  make_full_task(P, r, P.sel[idx], T, I);
  const int out = P.task_off[r] + k;
  P.tasks[out] = T;
  P.info[out] = I;
}

// Class of every full-SW task (sw_full_ring.cu): the widest row of its band + the edge cell, rounded up to
// 32/64/128/256 (class RING_CLASSES = wider than any ring that fits shared memory, served by the global-scratch
// kernels); then whether the alignment runs with the reverse-complement tie order (gen_st && Tflag: a
// compile-time variant of the kernels), then the band width in sixteenths of the ring -- the alignments of a
// warp run in lockstep over the widest band among them, so neighbours in the task order should look alike.
// key = class * 32 + revcmpl * 16 + width bucket.
#define FULL_KEYS ((RING_CLASSES + 1) * 32)
__global__ void classify_full_tasks_kernel(const FullTask *tasks, int n, int anchor_width, int match, int local,
                                           int Tflag, int max_class, uint32_t *key, uint32_t *key_count) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const FullTask T = tasks[t];
  int bw = 2;
  if (T.run) {
    for (int pass = 0; pass < (local ? 2 : 1); pass++) {  // local mode may redo with the threshold band
      const Rect rect = task_rect(T, anchor_width, match, pass == 0);
      for (int i = 0; i < T.rlen; i++) {
        int x_min, x_max;
        rect_x_range(rect, T.glen, i, x_min, x_max);
        bw = max(bw, x_max - x_min + 2);
      }
    }
  }
  int c = bw <= 32 ? 0 : bw <= 64 ? 1 : bw <= 128 ? 2 : bw <= 256 ? 3 : RING_CLASSES;
  if (c > max_class) c = RING_CLASSES;
  const int W = 32 << c;
  const int bucket = c < RING_CLASSES ? min(15, (bw * 16 - 1) / W) : 0;
  const uint32_t k = (uint32_t)(c * 32 + ((T.gen_st && Tflag) ? 16 : 0) + bucket);
  key[t] = k;
  // one atomic per warp and key: the counters are few and every task hits one of them
  const uint32_t peers = __match_any_sync(__activemask(), k);
  if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&key_count[k], (uint32_t)__popc(peers));
}
// exclusive prefix of the key counts -> first slot of every key in perm (one thread: FULL_KEYS values)
__global__ void full_key_offsets_kernel(const uint32_t *key_count, uint32_t *key_off) {
  uint32_t run = 0;
  for (int k = 0; k < FULL_KEYS; k++) {
    key_off[k] = run;
    run += key_count[k];
  }
}
__global__ void group_full_tasks_kernel(const uint32_t *key, int n, uint32_t *key_off, int32_t *perm) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t k = key[t];
  const uint32_t peers = __match_any_sync(__activemask(), k);
  const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(&key_off[k], (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  perm[base + (uint32_t)__popc(peers & ((1u << lane) - 1u))] = t;
}

// ---- packed fetch: only the alignments that exist travel to the host ---------------------------------------------
// A full-SW task that stayed below its threshold has score 0 and no edit script (sw-full-cs.c:1216-1226;
// mapping.c:390-398 in letter space): read_pass2 drops it whatever else happens.  The tasks that have an alignment
// are compacted on the device, in task order -- SelInfo, FullResult, and one byte record each (edit script, then the
// base qualities of post_sw) at 4-byte granularity -- so that the D2H copy carries what the host stage looks at.
__global__ void pack_flag_kernel(const FullResult *res, const SelInfo *info, const int32_t *read_len, int n, int post_sw,
                                 int cs, uint32_t *keep, uint32_t *units, unsigned long long *vcells) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long cells = 0;
  if (i < n) {
    const FullResult r = res[i];
    const bool k = r.score > 0 || r.ops_len > 0;
    keep[i] = k ? 1u : 0u;
    units[i] = k ? (uint32_t)(r.ops_len + (post_sw ? r.rmapped : 0) + 3) >> 2 : 0u;
    if (!cs) cells = (unsigned long long)info[i].w_len * (unsigned long long)read_len[info[i].read_idx];
  }
  if (!cs) {   // the sw_vector re-run of hit_run_full_sw (mapping.c:386) is counted for every task
    for (int o = 16; o > 0; o >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, o);
    if ((threadIdx.x & 31) == 0 && cells) atomicAdd(vcells, cells);
  }
}

__global__ void pack_scatter_kernel(const FullResult *res, const SelInfo *info, const uint8_t *ops, size_t ops_stride,
                                    const uint8_t *fqual, int max_rl, const uint32_t *keep_pos, const uint32_t *unit_off,
                                    int n, int post_sw, SelInfo *info2, FullResult *res2, uint8_t *pool) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || keep_pos[i + 1] == keep_pos[i]) return;
  const uint32_t p = keep_pos[i];
  FullResult r = res[i];
  const uint8_t *src = ops + ops_stride * (size_t)i + r.ops_start;
  uint8_t *dst = pool + 4 * (size_t)unit_off[i];
  for (int b = 0; b < r.ops_len; b++) dst[b] = src[b];
  if (post_sw) {
    const uint8_t *q = fqual + (size_t)max_rl * (size_t)i;
    for (int b = 0; b < r.rmapped; b++) dst[r.ops_len + b] = q[b];
  }
  r.ops_start = (int32_t)unit_off[i];   // packed: the record's offset in the pool, in units of 4 bytes
  info2[p] = info[i];
  res2[p] = r;
}

__global__ void pack_counts_kernel(const int32_t *task_off, const uint32_t *keep_pos, int n_reads, int32_t *n_kept) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_reads) n_kept[r] = (int32_t)(keep_pos[task_off[r + 1]] - keep_pos[task_off[r]]);
}

void free_pipeline(shrimp_gpu_ctx *ctx) {
  Pipeline *p = (Pipeline *)ctx->pipeline;
  if (!p) return;
  DevBuf *bufs[] = {&p->d_in, &p->d_reads, &p->d_read_len, &p->d_initbp, &p->d_hits, &p->d_rs_range, &p->d_counters,
                    &p->d_overflow, &p->d_overflow2, &p->d_scan_slab, &p->d_tie_ent, &p->d_tie_order, &p->d_tie_rec, &p->d_prof, &p->d_mp_tab, &p->d_mp_epoch, &p->d_xover, &p->d_quals, &p->d_fqual, &p->d_pstab, &p->d_gmtab, &p->d_scratch, &p->d_task[0], &p->d_task[1], &p->d_vtrue[0], &p->d_vtrue[1],
                    &p->d_slot, &p->d_writer, &p->d_sel, &p->d_nsel, &p->d_ftasks, &p->d_finfo, &p->d_fresults,
                    &p->d_frow, &p->d_fbp[0], &p->d_fbp[1], &p->d_fbp[2], &p->d_fbp[3], &p->d_fbp[4], &p->d_fbp[5], &p->d_fops,
                    &p->d_taskoff, &p->d_scan_tmp, &p->d_perm, &p->d_pair_min, &p->d_pair_max, &p->d_saved, &p->d_pairsel,
                    &p->d_npairsel, &p->d_taskof, &p->d_pairoff, &p->d_pk, &p->d_pk_info, &p->d_pk_res, &p->d_pk_pool};
  for (DevBuf *b : bufs) b->release();
  HostBuf *hb[] = {&p->h_info, &p->h_results, &p->h_ops, &p->h_nsel, &p->h_hits, &p->h_range, &p->h_xover, &p->h_fqual,
                   &p->h_pairsel, &p->h_npairsel, &p->h_saved};
  for (HostBuf *b : hb) b->release();
  delete p;
  ctx->pipeline = nullptr;
}

// restores ctx->stream on every exit path of run_full_sw
struct ctx_stream_guard {
  shrimp_gpu_ctx *c;
  cudaStream_t s;
  ~ctx_stream_guard() { c->stream = s; }
};

// Full SW over the n tasks of FP.tasks: classify by ring width, one ring launch per class and
// sub-batch (scratch bounded to ~2 GB of back-pointers), global-scratch kernels for the rest.
int run_full_sw(shrimp_gpu_ctx *ctx, DevBuf &d_perm, DevBuf &d_row, DevBuf *d_bp /*[RING_CLASSES+2], the last for short runs*/,
                       FullParams FP, int n, bool cs, uint32_t *d_cls_count) {
  if (n <= 0) return SHRIMP_OK;
  cudaStream_t st = ctx->stream;
  // d_perm: perm[n] grouped by key | key[n] | key counts | key offsets
  SH_TRY(d_perm.ensure(((size_t)2 * n + 2 * FULL_KEYS) * 4));
  int32_t *perm = d_perm.as<int32_t>();
  uint32_t *d_key = (uint32_t *)perm + n, *d_key_count = d_key + n, *d_key_off = d_key_count + FULL_KEYS;
  (void)d_cls_count;
  SH_CUDA(cudaMemsetAsync(d_key_count, 0, FULL_KEYS * 4, st));
  int max_class = -1;
  for (int c = 0; c < RING_CLASSES; c++)
    if (ring_fits(cs, 32 << c)) max_class = c;
  classify_full_tasks_kernel<<<(n + 127) / 128, 128, 0, st>>>(FP.tasks, n, FP.anchor_width, FP.match, FP.local, FP.Tflag,
                                                              max_class, d_key, d_key_count);
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_FULL);
  full_key_offsets_kernel<<<1, 1, 0, st>>>(d_key_count, d_key_off);
  group_full_tasks_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_key, n, d_key_off, perm);
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_FULL);
  SH_LAUNCHED(ctx, ST_FULL);
  uint32_t kc[FULL_KEYS];
  SH_CUDA(cudaMemcpyAsync(kc, d_key_count, sizeof(kc), cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaStreamSynchronize(st));
  uint32_t cls[RING_CLASSES + 1], cls_rev0[RING_CLASSES + 1], cls_first[RING_CLASSES + 1];
  {
    uint32_t run = 0;
    for (int c = 0; c <= RING_CLASSES; c++) {
      cls_first[c] = run;
      cls[c] = cls_rev0[c] = 0;
      for (int k = 0; k < 32; k++) {
        cls[c] += kc[c * 32 + k];
        if (k < 16) cls_rev0[c] += kc[c * 32 + k];
      }
      run += cls[c];
    }
  }
  // the classes are independent: class c runs on its own stream (class 0, the bulk, on the main one) so the
  // small wide-band launches, each bounded by the latency of one alignment, overlap the bulk
  const size_t budget = (size_t)2 << 30;
  SH_CUDA(cudaEventRecord(ctx->fork_ev, st));
  ctx_stream_guard guard{ctx, st};
  cudaStream_t side_st = ctx->aux[SHRIMP_AUX_STREAMS - 1];
  bool side_used = false;
  for (int c = 0; c <= RING_CLASSES; c++) {
    if (cls[c] == 0) continue;
    const int count = (int)cls[c];
    const bool ring = c < RING_CLASSES;
    const int W = 32 << c;
    const size_t states = cs ? 12 : 3;
    const size_t per_task = ring ? (size_t)FP.max_rlen * W * (cs ? 8 : 1)
                                 : states * (FP.max_glen + 1) * 4 + (size_t)FP.max_rlen * FP.max_glen * (cs ? 12 : 1);
    int batch = (int)std::min<size_t>((size_t)count, std::max<size_t>(1024, budget / per_task));
    batch = (batch + 127) & ~127;
    if (ring) {
      SH_TRY(d_bp[c].ensure((size_t)FP.max_rlen * W * (cs ? 8 : 1) * batch));
    } else {
      SH_TRY(d_row.ensure(states * (FP.max_glen + 1) * 4 * batch));
      SH_TRY(d_bp[c].ensure((size_t)FP.max_rlen * FP.max_glen * (cs ? 12 : 1) * batch));
    }
    cudaStream_t cst = c == 0 ? st : ctx->aux[(c - 1) % SHRIMP_AUX_STREAMS];
    if (c > 0) SH_CUDA(cudaStreamWaitEvent(cst, ctx->fork_ev, 0));
    ctx->stream = cst;  // the launchers use ctx->stream
    // the forward-order tasks first, then the reverse-complement-order ones: a launch holds one kind.  Within a
    // kind the tasks are ordered by band width; runs of width buckets (at least 8192 tasks each) get a ring just wide
    // enough for them -- the ring is what limits the resident warps of these kernels.
    for (int rev = 0; rev < 2; rev++) {
      int seg0 = rev ? (int)cls_rev0[c] : 0;
      int acc = 0, kmax = -1;
      for (int k = 0; k < 16; k++) {
        const int nk = (int)kc[c * 32 + rev * 16 + k];
        if (nk > 0) kmax = k;
        acc += nk;
        if (acc == 0 || ((acc < 8192 || !cs) && k < 15)) continue;   // letter space: one run (thread-per-alignment rings are cheap)
        const int Wl = (ring && cs) ? std::max(4, (kmax + 1) * W / 16) : W;
        const size_t per_task_l = ring ? (size_t)FP.max_rlen * Wl * (cs ? 8 : 1) : per_task;
        int batch_l = ring ? (int)std::min<size_t>((size_t)acc, std::max<size_t>(1024, budget / per_task_l)) : batch;
        batch_l = (batch_l + 127) & ~127;
        if (ring && (size_t)batch_l * per_task_l > (size_t)batch * per_task) batch_l = batch;
        // a short run (the widest bands of a kind: a handful of alignments, a launch bounded by the latency of one)
        // goes to a side stream with back-pointers of its own and overlaps the bulk
        const bool side = ring && cs && acc < 8192 && acc <= batch_l;
        DevBuf &bpbuf = side ? d_bp[RING_CLASSES + 1] : d_bp[c];
        if (side) {
          SH_TRY(bpbuf.ensure((size_t)batch_l * per_task_l));   // grows only: earlier short runs in flight keep their bytes
          if (!side_used) SH_CUDA(cudaStreamWaitEvent(side_st, ctx->fork_ev, 0));
          side_used = true;
          ctx->stream = side_st;
        }
        for (int b0 = seg0; b0 < seg0 + acc; b0 += batch_l) {
          FullParams Q = FP;
          Q.perm = perm + cls_first[c] + b0;
          Q.n_tasks = std::min(batch_l, seg0 + acc - b0);
          Q.rev = rev;
          Q.NT = batch_l;
          Q.W = Wl;
          Q.row = Q.row_cs = d_row.as<int32_t>();
          Q.bp = Q.bp_cs = bpbuf.as<uint8_t>();
          Q.bp64 = bpbuf.as<unsigned long long>();
          int rc;
          if (ring) rc = launch_sw_full_ring(ctx, Q, cs);
          else if (cs) rc = launch_sw_full_cs(ctx, Q);
          else rc = launch_sw_full_ls(ctx, Q);
          if (rc != SHRIMP_OK) return rc;
        }
        if (side) ctx->stream = cst;
        seg0 += acc;
        acc = 0;
        kmax = -1;
      }
    }
    ctx->stream = st;
    if (c > 0) {
      SH_CUDA(cudaEventRecord(ctx->join_ev[(c - 1) % SHRIMP_AUX_STREAMS], cst));
      SH_CUDA(cudaStreamWaitEvent(st, ctx->join_ev[(c - 1) % SHRIMP_AUX_STREAMS], 0));
    }
  }
  if (side_used) {
    SH_CUDA(cudaEventRecord(ctx->join_ev[SHRIMP_AUX_STREAMS - 1], side_st));
    SH_CUDA(cudaStreamWaitEvent(st, ctx->join_ev[SHRIMP_AUX_STREAMS - 1], 0));
  }
  return SHRIMP_OK;
}

// ---- host stage ---------------------------------------------------------------------------------
static int cmp_gen_start(const void *e1, const void *e2) {  // mapping.c:1485-1494
  const HostHit *a = *(HostHit *const *)e1, *b = *(HostHit *const *)e2;
  if (a->info.cn != b->info.cn) return a->info.cn - b->info.cn;
  if (a->info.gen_st != b->info.gen_st) return a->info.gen_st - b->info.gen_st;
  return a->res.genome_start - b->res.genome_start;
}
static int cmp_gen_end(const void *e1, const void *e2) {  // mapping.c:1496-1506
  const HostHit *a = *(HostHit *const *)e1, *b = *(HostHit *const *)e2;
  if (a->info.cn != b->info.cn) return a->info.cn - b->info.cn;
  if (a->info.gen_st != b->info.gen_st) return a->info.gen_st - b->info.gen_st;
  return (-a->res.genome_start - a->res.rmapped + a->res.deletions - a->res.insertions) -
         (-b->res.genome_start - b->res.rmapped + b->res.deletions - b->res.insertions);
}
static int cmp_score(const void *e1, const void *e2) {  // mapping.c:1479-1482
  return (*(HostHit *const *)e2)->pass2_key - (*(HostHit *const *)e1)->pass2_key;
}
static void dedup_pass(HostHit **h, int *n, int (*cmp)(const void *, const void *)) {  // mapping.c:1552-1575
  qsort(h, *n, sizeof(h[0]), cmp);
  int i = 0, k = 0;
  while (i < *n) {
    int max = h[i]->pass2_key, max_idx = i, j = i + 1;
    while (j < *n && !cmp(&h[i], &h[j])) {
      if (h[j]->pass2_key > max) {
        max = h[j]->pass2_key;
        max_idx = j;
      }
      j++;
    }
    if (max_idx != k) h[k] = h[max_idx];
    k++;
    i = j;
  }
  *n = k;
}


// ---- chunk stages -------------------------------------------------------------------------------
// Validation, parameter block, upload of the reads and their reverse complements.
int chunk_begin(Chunk &C, shrimp_gpu_ctx *ctx, const shrimp_map_params *mp, int n_reads, const uint32_t *reads,
                int stride, const int32_t *read_len, const int8_t *initbp, bool resident, const char *who) {
  if (resident) {
    Pipeline *pp = ctx ? (Pipeline *)ctx->pipeline : nullptr;
    if (!pp || pp->res_n_reads <= 0) {
      set_error("%s: no reads resident; call shrimp_gpu_map_reads first", who);
      return SHRIMP_E_STATE;
    }
    n_reads = pp->res_n_reads;
    stride = pp->res_stride;
    read_len = pp->res_read_len.data();
    reads = (const uint32_t *)1;  // not dereferenced
    if (genome_of(ctx) && genome_of(ctx)->colour_space) initbp = (const int8_t *)1;
  }
  if (!ctx || !mp || !reads || !read_len || n_reads < 0 || stride <= 0) {
    set_error("%s: invalid argument", who);
    return SHRIMP_E_ARG;
  }
  DeviceGenome *g = genome_of(ctx);
  if (!g || !g->have_index) {
    set_error("%s: genome/index not resident (shrimp_gpu_genome_load + shrimp_gpu_index_build)", who);
    return SHRIMP_E_STATE;
  }
  if (!ctx->sw.valid) {
    set_error("%s: shrimp_gpu_sw_setup() has not been called", who);
    return SHRIMP_E_STATE;
  }
  const bool cs = g->colour_space != 0;
  if (cs != (ctx->sw.use_colours != 0)) {
    set_error("%s: genome and scoring set-up disagree about colour space", who);
    return SHRIMP_E_STATE;
  }
  if (cs && !initbp) {
    set_error("%s: colour-space reads need initbp", who);
    return SHRIMP_E_ARG;
  }
  C.ctx = ctx;
  C.g = g;
  C.mp = mp;
  C.n_reads = n_reads;
  C.stride = stride;
  C.cs = cs;
  C.read_len = read_len;
  C.n_ori = cs ? 2 : 1;
  C.max_rl = 0;
  C.sum_rl = 0;
  for (int r = 0; r < n_reads; r++) {
    if (read_len[r] < 0 || read_len[r] > stride * 8) {
      set_error("%s: read %d has length %d (stride holds %d)", who, r, read_len[r], stride * 8);
      return SHRIMP_E_ARG;
    }
    if (read_len[r] > C.max_rl) C.max_rl = read_len[r];
    C.sum_rl += read_len[r];
  }
  if (C.max_rl > ctx->sw.max_read_len) {
    set_error("%s: read length %d exceeds the qrlen given at setup (%d)", who, C.max_rl, ctx->sw.max_read_len);
    return SHRIMP_E_ARG;
  }
  if (mp->num_outputs < 1 || mp->num_tmp_outputs < mp->num_outputs || mp->num_tmp_outputs > 4096) {
    set_error("%s: num_outputs %d / num_tmp_outputs %d (need 1 <= num_outputs <= num_tmp_outputs)", who, mp->num_outputs,
              mp->num_tmp_outputs);
    return SHRIMP_E_ARG;
  }
  if (mp->region_bits < 4 || mp->region_bits > 24 || mp->region_overlap < 0 ||
      mp->region_overlap >= (1 << mp->region_bits)) {
    set_error("%s: region_bits %d / region_overlap %d out of range", who, mp->region_bits, mp->region_overlap);
    return SHRIMP_E_ARG;
  }
  if (!resident && cs && mp->crossover_scores && mp->crossover_stride > 0 && mp->crossover_stride < C.max_rl) {
    set_error("%s: crossover_stride %d is shorter than the longest read (%d)", who, mp->crossover_stride, C.max_rl);
    return SHRIMP_E_ARG;
  }
  if (!resident && cs && mp->read_quals && mp->qual_stride > 0 &&
      mp->qual_stride < C.max_rl + (mp->qual_vector_offset > 0 ? mp->qual_vector_offset : 0)) {
    set_error("%s: qual_stride %d is shorter than the longest read (%d)", who, mp->qual_stride, C.max_rl);
    return SHRIMP_E_ARG;
  }
  SH_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->pipeline) ctx->pipeline = new Pipeline();
  Pipeline *pl = C.pl = (Pipeline *)ctx->pipeline;
  cudaStream_t st = ctx->stream;
  const SwScores &sw = ctx->sw;

  MapParamsDev &M = C.M;
  memset(&M, 0, sizeof(M));
  M.colour_space = cs;
  M.match_mode = mp->match_mode;
  M.gapless = mp->gapless;
  M.hash_filter_calls = mp->hash_filter_calls;
  M.use_region_counts = (mp->match_mode == 2 && mp->use_regions) ? 1 : 0;  // gmapper.c:2610-2615
  M.region_bits = mp->region_bits;
  M.region_overlap = mp->region_overlap;
  M.list_cutoff = mp->list_cutoff;
  M.num_tmp_outputs = mp->num_tmp_outputs;
  M.min_matches = mp->match_mode;  // gmapper.c:2624
  M.match = sw.match;
  M.b_gap_open = -sw.b_open;
  M.b_gap_ext = -sw.b_ext;
  M.window_len = mp->window_len;
  M.window_len_frac = mp->window_len / 100.0;
  M.wgen_thr = mp->window_gen_threshold;
  M.wgen_frac = mp->window_gen_threshold / 100.0;
  M.vect_thr = mp->sw_vect_threshold;
  M.vect_frac = mp->sw_vect_threshold / 100.0;
  M.full_thr = mp->sw_full_threshold;
  M.full_frac = mp->sw_full_threshold / 100.0;
  M.overlap_thr = mp->window_overlap;
  M.overlap_frac = mp->window_overlap / 100.0;
  M.Gflag = mp->Gflag;
  M.Tflag = mp->Tflag;
  M.anchor_width = sw.anchor_width;
  C.max_wl = (int)(unsigned short)abs_or_pct_d(M.window_len, M.window_len_frac, (double)C.max_rl);
  if (C.max_wl > sw.max_window_len) {
    set_error("%s: window length %d exceeds the dblen given at setup (%d)", who, C.max_wl, sw.max_window_len);
    return SHRIMP_E_ARG;
  }
  C.ops_stride = (size_t)C.max_rl + C.max_wl;
  C.post_sw = cs && mp->compute_mapping_qualities;

  GenomeView &G = C.G;
  G.ls = g->d_ls.as<uint32_t>();
  G.ls_rc = g->d_ls_rc.as<uint32_t>();
  G.cs = g->d_cs.as<uint32_t>();
  G.cs_rc = g->d_cs_rc.as<uint32_t>();
  G.contig_off = g->d_off.as<uint32_t>();
  G.contig_len = g->d_len.as<uint32_t>();
  G.num_contigs = g->num_contigs;
  memset(&C.IV, 0, sizeof(C.IV));
  for (int sn = 0; sn < g->seeds.n_seeds; sn++) {
    C.IV.offs[sn] = g->d_offs[sn].as<uint32_t>();
    C.IV.pos[sn] = g->d_pos[sn].as<uint32_t>();
    C.IV.head_off[sn] = g->head_off[sn];
  }
  if (n_reads == 0) return SHRIMP_OK;

  // ---- upload + reverse complements ------------------------------------------------------------
  const size_t in_bytes = (size_t)n_reads * stride * 4;
  SH_TRY(pl->d_in.ensure(in_bytes));
  SH_TRY(pl->d_reads.ensure(in_bytes * 2));
  SH_TRY(pl->d_read_len.ensure((size_t)n_reads * 4));
  if (!resident) {
    SH_CUDA(cudaMemcpyAsync(pl->d_in.p, reads, in_bytes, cudaMemcpyHostToDevice, st));
    SH_CUDA(cudaMemcpyAsync(pl->d_read_len.p, read_len, (size_t)n_reads * 4, cudaMemcpyHostToDevice, st));
    if (cs) {
      SH_TRY(pl->d_initbp.ensure((size_t)n_reads));
      SH_CUDA(cudaMemcpyAsync(pl->d_initbp.p, initbp, (size_t)n_reads, cudaMemcpyHostToDevice, st));
    }
    pl->qual_stride = 0;
    if (cs && mp->compute_mapping_qualities && mp->read_quals && mp->qual_stride > 0) {
      SH_TRY(pl->d_quals.ensure((size_t)n_reads * mp->qual_stride));
      SH_CUDA(cudaMemcpyAsync(pl->d_quals.p, mp->read_quals, (size_t)n_reads * mp->qual_stride, cudaMemcpyHostToDevice, st));
      pl->qual_stride = mp->qual_stride;
    }
    pl->xover_stride = 0;
    if (cs && mp->crossover_scores && mp->crossover_stride > 0) {
      // per-position crossover scores (reads with qualities): 16 bits each on the device
      const int xs = mp->crossover_stride;
      SH_TRY(pl->h_xover.ensure((size_t)n_reads * xs * 2));
      int16_t *hx = pl->h_xover.as<int16_t>();
      for (size_t q = 0; q < (size_t)n_reads * xs; q++) {
        const int32_t v = mp->crossover_scores[q];
        hx[q] = (int16_t)(v < -32768 ? -32768 : v > 0 ? 0 : v);
      }
      SH_TRY(pl->d_xover.ensure((size_t)n_reads * xs * 2));
      SH_CUDA(cudaMemcpyAsync(pl->d_xover.p, hx, (size_t)n_reads * xs * 2, cudaMemcpyHostToDevice, st));
      pl->xover_stride = xs;
    }
    pl->res_n_reads = n_reads;
    pl->res_stride = stride;
    pl->res_read_len.assign(read_len, read_len + n_reads);
    C.read_len = pl->res_read_len.data();
    pl->h2d_bytes = in_bytes + (size_t)n_reads * 4 + (cs ? (size_t)n_reads : 0);
  }
  SH_TRY(pl->d_counters.ensure(64 * 4));
  SH_TRY(pl->d_rs_range.ensure((size_t)n_reads * 2 * sizeof(uint2)));
  C.cnt = pl->d_counters.as<uint32_t>();  // [0] hits_used [1] n_overflow [2] status [8..15] stats [16,17] full cells
                                          // [20..23] vector task stats [32..36] ring-class counts
  return SHRIMP_OK;
}

// read reverse complement + seed scan -> hits, rs_range
int chunk_scan(Chunk &C) {
  shrimp_gpu_ctx *ctx = C.ctx;
  Pipeline *pl = C.pl;
  DeviceGenome *g = C.g;
  cudaStream_t st = ctx->stream;
  const int n_reads = C.n_reads, stride = C.stride, max_rl = C.max_rl;
  uint32_t *cnt = C.cnt;
  ScopedStage ss(ctx, ST_SCAN);
  revcomp_reads_kernel<<<(n_reads + 127) / 128, 128, 0, st>>>(pl->d_in.as<uint32_t>(), pl->d_reads.as<uint32_t>(),
                                                              stride, n_reads, pl->d_read_len.as<int32_t>(),
                                                              C.cs ? pl->d_initbp.as<int8_t>() : nullptr, C.cs,
                                                              C.M.rev_mate[0], C.M.rev_mate[1]);
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_SCAN);
  // expected number of list entries per read strand: K(r) * L / 4^W
  double est = 0;
  const double avg_rl = (double)C.sum_rl / n_reads;
  const int mkp = C.cs ? 1 : 0;
  int K_max = 0;
  for (int sn = 0; sn < g->seeds.n_seeds; sn++) {
    est += std::max(0.0, avg_rl - g->seeds.span[sn] + 1) * ((double)g->total[sn] / (double)g->nbuckets[sn]);
    K_max += std::max(0, max_rl - g->seeds.span[sn] + 1 - mkp);
  }
  const bool filt = C.M.use_region_counts != 0;
  // warp kernel: bitmaps of >= 16 bits per expected entry (up to 2^17 bits) and candidate slots for twice the
  // expected survivors (true: entries that share a 2 kb region, ~3 region touches each; false: bitmap collisions);
  // without a region filter every entry is a candidate
  const double L_total = (double)g->total_len;
  int bm_log2 = 10, cap = 128;
  bool small_useful = true;
  if (filt) {
    while (bm_log2 < 17 && (1 << bm_log2) < est * 16) bm_log2++;
    const double s_true = est * std::min(1.0, est * 2100.0 / std::max(L_total, 1.0)) * 3.0;
    const double s_false = est * std::min(1.0, est * 1.1 / (double)(1 << bm_log2));
    cap = 256;
    // + the read's own locus: every k-mer may hit it
    const double want = 2.0 * (s_true + s_false) + 1.5 * K_max + 32;
    // in steps of 64: the slab decides how many warps an SM holds.  (The bitonic sort pads to the next power of two
    // of the survivors, possibly past `cap`: the bytes behind ent are the bitmaps, dead by then, scan_layout.)
    cap = 128;
    while (cap < 2048 && cap < want) cap += 64;
    small_useful = est * 8 <= (double)(1 << bm_log2) && want <= 2048;
  } else {
    bm_log2 = 5;
    while (cap < 2048 && cap < est * 3 + 64) cap <<= 1;
  }
  // the anchors (16 B per candidate) reuse the bitmaps once those are at least as large
  const bool alias_rec = filt && ((size_t)2 << bm_log2) / 8 >= (size_t)cap * 16;
  // the k-mer tables double as first_of[] of the tie replay, indexed by slot = sn * max_n_kmers + i
  const int K_slots = g->seeds.n_seeds * std::max(1, max_rl - g->seeds.min_span + 1);
  const int k_cap = std::max(32, std::min(std::max(K_max, K_slots), 1024));
  // dense regime (thousands of list entries per strand): every strand through the CTA kernel
  double cta_min_est = 1500.0;
  if (const char *e = getenv("SHRIMP_SCAN_CTA_MIN_EST")) cta_min_est = atof(e);
  if (est >= cta_min_est) small_useful = false;
  if (getenv("SHRIMP_SCAN_FORCE_BIG")) small_useful = false;  // test hook: every strand through the CTA kernel
  if (C.mp_mode) small_useful = false;   // mate-pair region counts live in the CTA kernel
  // CTA kernel: exact region bitmaps over partitions of 2^cta_bm_log2 regions, candidate slots for twice the
  // expected survivors
  const int big_k_cap = std::max(32, std::max(K_max, K_slots));
  int cta_bm_log2 = 5, cta_n_part = 1, cta_cap = 512, cta_win = 1024;
  bool cta_hashed = false;
  // cursor walk over the index lists (scan.cu) unless the mate-pair region tables are in play; SHRIMP_SCAN_WALK=0
  // brings the staged-window passes back (test hook)
  bool cta_walk = !C.mp_mode;
  if (const char *e = getenv("SHRIMP_SCAN_WALK")) cta_walk = cta_walk && atoi(e) != 0;
  {
    const double n_regions = L_total / (double)(1u << C.M.region_bits) + 2.0;
    if (filt && C.mp_mode) {
      // exact region tables in global memory: no bitmaps; every candidate is a survivor
      const double lam = est / n_regions;
      const double want = 3.0 * est * std::min(1.0, 1.1 * lam) + 2.0 * K_max + 64;
      while (cta_cap < 8192 && cta_cap < want) cta_cap <<= 1;
    } else if (filt) {
      int exact_log2 = 5;
      while (exact_log2 < 31 && (double)(1u << exact_log2) < n_regions) exact_log2++;
      // exact bitmaps (one bit per 2 kb region, partitions of at most 2^18 regions) unless they would be far
      // larger than the strand's entries call for: then hashed bitmaps of >= 8 bits per expected entry, whose
      // false candidates the exact neighbour test removes
      cta_hashed = (double)(1u << exact_log2) > 16.0 * std::max(est, 64.0) && 8.0 * est <= (double)(1u << 17);
      if (const char *e = getenv("SHRIMP_SCAN_HASHED")) cta_hashed = atoi(e) != 0;
      int bm_max = 18;
      if (const char *e = getenv("SHRIMP_SCAN_BM_LOG2")) {  // test hook
        bm_max = std::max(5, std::min(19, atoi(e)));
        if (!getenv("SHRIMP_SCAN_HASHED")) cta_hashed = false;
      }
      double s_false = 0;
      if (cta_hashed) {
        cta_bm_log2 = 10;
        while (cta_bm_log2 < bm_max && (double)(1u << cta_bm_log2) < 8.0 * est) cta_bm_log2++;
        if (getenv("SHRIMP_SCAN_BM_LOG2")) cta_bm_log2 = bm_max;
        s_false = est * std::min(1.0, 1.1 * est / (double)(1u << cta_bm_log2));
      } else {
        cta_bm_log2 = std::min(10, bm_max);
        while (cta_bm_log2 < bm_max && (double)(1u << cta_bm_log2) < n_regions) cta_bm_log2++;
        cta_n_part = (int)std::ceil(n_regions / (double)(1u << cta_bm_log2));
      }
      const double lam = est / n_regions;   // entries per region
      const double s_true = est * std::min(1.0, 1.1 * lam);
      const double want = 1.3 * (s_true + s_false) + K_max + 64;
      while (cta_cap < 8192 && cta_cap < want) cta_cap <<= 1;
    } else {
      while (cta_cap < 8192 && cta_cap < est * 3 + 64) cta_cap <<= 1;
    }
    // staging window: about half of a partition's entries, 1024..4096
    while (cta_win < 4096 && cta_win < 0.4 * est / cta_n_part) cta_win <<= 1;
    if (est / cta_n_part > 6000.0) cta_win = 8192;   // hg18 scale: a partition in one window, staged once for both passes
    if (const char *e = getenv("SHRIMP_SCAN_WIN")) cta_win = std::max(64, atoi(e));  // test hook: several windows
    if (cta_walk) cta_win = 16;   // the cursor walk stages nothing
    if (const char *e = getenv("SHRIMP_SCAN_CTA_CAP")) cta_cap = std::max(32, atoi(e));  // test hook: global slabs
    while (cta_cap > 256 && scan_cta_smem_bytes(cta_cap, max_rl, big_k_cap, cta_bm_log2, cta_n_part, cta_win, false) > 200 * 1024)
      cta_cap >>= 1;
    // a slab a little smaller than the power of two when that lets one more CTA live on an SM (the serial phases of a
    // strand -- the collapse runs on one warp -- then overlap another strand's parallel ones)
    if (!getenv("SHRIMP_SCAN_CTA_CAP")) {
      auto per_sm_of = [&](int c) {
        return (int)((size_t)(227 * 1024) / (scan_cta_smem_bytes(c, max_rl, big_k_cap, cta_bm_log2, cta_n_part, cta_win, false) + 1024));
      };
      const int per0 = per_sm_of(cta_cap);
      for (int c2 = cta_cap - cta_cap / 16; c2 >= cta_cap - cta_cap / 8 && per0 < 8; c2 -= cta_cap / 16)
        if (per_sm_of(c2) > per0) {
          cta_cap = c2;
          break;
        }
    }
  }
  const int g_cap = 65535;   // global-slab pass: 16-bit candidate indices
  if (scan_cta_smem_bytes(cta_cap, max_rl, big_k_cap, cta_bm_log2, cta_n_part, cta_win, false) > 226 * 1024) {
    set_error("seed scan: reads of %d bases with these seeds need more shared memory than a CTA has", max_rl);
    return SHRIMP_E_RANGE;
  }
  int k_max = g->seeds.n_seeds * std::max(1, max_rl);
  if (pl->hits_cap == 0) pl->hits_cap = (uint32_t)std::max<long long>(1 << 20, (long long)n_reads * 2 * 16);
  for (int attempt = 0;; attempt++) {
    SH_TRY(pl->d_hits.ensure((size_t)pl->hits_cap * sizeof(DevHit)));
    SH_TRY(pl->d_overflow.ensure((size_t)n_reads * 2 * 4));
    SH_CUDA(cudaMemsetAsync(cnt, 0, 64 * 4, st));
    ScanParams P;
    memset(&P, 0, sizeof(P));
    P.G = C.G;
    P.I = C.IV;
    P.S = g->seeds;
    P.M = C.M;
    P.reads = pl->d_reads.as<uint32_t>();
    P.stride = stride;
    P.n_reads = n_reads;
    P.read_len = pl->d_read_len.as<int32_t>();
    P.hits = pl->d_hits.as<DevHit>();
    P.hits_cap = pl->hits_cap;
    P.hits_used = cnt + 0;
    P.rs_range = pl->d_rs_range.as<uint2>();
    P.overflow = pl->d_overflow.as<uint32_t>();
    P.n_overflow = cnt + 1;
    P.status = cnt + 2;
    P.stats = cnt + 8;
    P.stats64 = (unsigned long long *)(cnt + 40);
    P.k_max = k_max;
    P.max_rl = max_rl;
    uint32_t h3[3] = {0, 0, 0};
    if (small_useful) {
      // warp-per-strand pass over all read strands
      P.cap = cap;
      P.k_cap = k_cap;
      P.bm_log2 = bm_log2;
      P.alias_rec = alias_rec ? 1 : 0;
      P.stream = est >= 8.0 * std::max(1, K_max) ? 1 : 0;   // average list of 8+ positions
      // staging slab for the strand's list entries: the expected random entries + one per k-mer for the read's own
      // locus, 128..2048 (strands with more take the two-pass path)
      int stash = 0;
      if (filt) {
        stash = 128;
        while (stash < 2048 && stash < 1.2 * est + K_max) stash += 32;
      }
      if (const char *e = getenv("SHRIMP_SCAN_STASH")) stash = std::max(0, std::min(4096, atoi(e)));
      P.stash = stash;
      // CTA size that keeps the most warps resident (227 KB of shared memory, 32 CTAs and 64 warps per SM)
      int warps = 1, ctas_per_sm = 1, best = 0;
      for (int w = SCAN_WARPS_HOST; w >= 1; w >>= 1) {
        const size_t sm_w = scan_smem_bytes(cap, max_rl, k_cap, bm_log2, w, stash) + 1024;
        if (sm_w > 220 * 1024) continue;
        const int c = (int)std::min<size_t>(std::min<size_t>((size_t)(226 * 1024) / sm_w, 32), (size_t)(64 / w));
        if (c * w > best) {
          best = c * w;
          warps = w;
          ctas_per_sm = c;
        }
      }
      const size_t smem = scan_smem_bytes(cap, max_rl, k_cap, bm_log2, warps, stash);
      (void)smem;
      int n_ctas = ctx->sm_count * ctas_per_sm;
      n_ctas = std::min<long long>(n_ctas, ((long long)n_reads * 2 + warps - 1) / warps);
      P.scratch_ints = 2 * k_max + 2 * cap;
      SH_TRY(pl->d_scratch.ensure((size_t)n_ctas * warps * P.scratch_ints * 4));
      P.scratch = pl->d_scratch.as<int32_t>();
      SH_TRY(launch_scan(ctx, P, warps, n_ctas));
      SH_CUDA(cudaMemcpyAsync(h3, cnt, 12, cudaMemcpyDeviceToHost, st));
      SH_CUDA(cudaStreamSynchronize(st));
    }
    if ((!small_useful || h3[1] > 0) && !(h3[2] & 1u)) {
      // CTA-per-strand passes: the strands the warp kernel passed on, or every strand when the lists are long.
      // Level 0: slab sized for the expected survivors; level 1: the largest shared-memory slab (repeats,
      // low-complexity reads); level 2: candidate arrays in global slabs.
      SH_TRY(pl->d_overflow2.ensure((size_t)n_reads * 2 * 4));
      uint32_t *lists[2] = {pl->d_overflow.as<uint32_t>(), pl->d_overflow2.as<uint32_t>()};
      int cur = 0;   // list holding the current work (valid when `have_list`)
      bool have_list = small_useful;
      uint32_t n_work = small_useful ? h3[1] : C.mp_mode ? (uint32_t)n_reads / 2u : 2u * (uint32_t)n_reads;
      C.scan_big = small_useful ? h3[1] : 2u * (uint32_t)n_reads;
      P.k_cap = big_k_cap;
      P.bm_log2 = cta_bm_log2;
      P.n_part = cta_n_part;
      P.win = cta_win;
      P.bm_hashed = cta_hashed ? 1 : 0;
      // lanes per index list: a warp streams 4 positions per lane and step
      double avg_list = est / std::max(1, K_max);
      P.lanes_per_list_log2 = avg_list > 64 ? 5 : avg_list > 32 ? 4 : avg_list > 12 ? 3 : 2;
      P.walk = cta_walk ? 1 : 0;
      {   // 64 sort bins over the genome's positions
        int sh = 0;
        while (sh < 31 && ((unsigned long long)(L_total > 1 ? L_total - 1 : 0) >> sh) >= 64ull) sh++;
        P.sort_shift = sh;
        if (getenv("SHRIMP_SCAN_NO_BINS")) P.sort_shift = 32;   // test hook: everything in one bin -> the CTA-wide network
      }
      if (cta_walk) {   // lanes per list by the entries a list has per tile; 8 lanes read one 32-byte sector
        const double per_tile = avg_list / std::max(1, (filt && !cta_hashed) ? cta_n_part : 1);
        P.lanes_per_list_log2 = per_tile > 96 ? 5 : per_tile > 40 ? 4 : per_tile > 5 ? 3 : 2;
      }
      if (const char *e = getenv("SHRIMP_SCAN_LANES_LOG2")) P.lanes_per_list_log2 = std::max(0, std::min(5, atoi(e)));
      if (pl->tie_cap == 0) pl->tie_cap = (uint32_t)std::max<long long>(1 << 20, (long long)n_reads * 2 * 32);
      SH_TRY(pl->d_tie_ent.ensure((size_t)pl->tie_cap * 8));
      SH_TRY(pl->d_tie_order.ensure((size_t)pl->tie_cap * 2));
      SH_TRY(pl->d_tie_rec.ensure((size_t)n_reads * 2 * sizeof(uint4)));
      P.tie_ent = pl->d_tie_ent.as<unsigned long long>();
      P.tie_order = pl->d_tie_order.as<uint16_t>();
      P.tie_rec = pl->d_tie_rec.as<uint4>();
      P.tie_used = cnt + 6;
      P.n_tie = cnt + 5;
      P.tie_cap = pl->tie_cap;
      P.tie_rec_cap = 2u * (uint32_t)n_reads;
      P.resume = 0;
      P.mp_mode = C.mp_mode;
      P.pair_mode = C.pair_mode;
      P.min_insert = C.min_insert;
      P.max_insert = C.max_insert;
      if (C.mp_mode) {
        // four region tables per CTA, (epoch << 8) | flags per 2 kb region; zeroed when (re)allocated, epochs persist
        const int mp_regions = (int)(L_total / (double)(1u << C.M.region_bits)) + 2;
        const size_t ctas_max = (size_t)ctx->sm_count * 8;
        const size_t ints = ctas_max * 4 * (size_t)mp_regions;
        if (pl->mp_tab_ints != ints) {
          SH_TRY(pl->d_mp_tab.ensure(ints * 4));
          SH_TRY(pl->d_mp_epoch.ensure(ctas_max * 4));
          SH_CUDA(cudaMemsetAsync(pl->d_mp_tab.p, 0, ints * 4, st));
          SH_CUDA(cudaMemsetAsync(pl->d_mp_epoch.p, 0, ctas_max * 4, st));
          pl->mp_tab_ints = ints;
        }
        P.mp_tab = pl->d_mp_tab.as<uint32_t>();
        P.mp_epoch = pl->d_mp_epoch.as<uint32_t>();
        P.mp_regions = mp_regions;
        if (K_slots >= 0x8000) {
          set_error("seed scan: reads too long for the paired match modes that use mate-pair region counts");
          return SHRIMP_E_RANGE;
        }
      }
      P.prof = nullptr;
      if (getenv("SHRIMP_SCAN_PROF")) {
        SH_TRY(pl->d_prof.ensure(16 * 8));
        SH_CUDA(cudaMemsetAsync(pl->d_prof.p, 0, 16 * 8, st));
        P.prof = pl->d_prof.as<unsigned long long>();
      }
      uint32_t level_work[3] = {0, 0, 0};
      // global-slab launches: the candidate arrays are in global memory, shared memory holds the bitmaps and the k-mer
      // tables only -- as many 512-thread CTAs per SM as that leaves room for (three at most: 1536 threads)
      const int glob_per_sm = (int)std::max<size_t>(1, std::min<size_t>(3, (size_t)(227 * 1024) /
          (scan_cta_smem_bytes(1, max_rl, big_k_cap, cta_bm_log2, cta_n_part, cta_win, true) + 1024)));
      int big_slab = 8192;
      while (big_slab > cta_cap && scan_cta_smem_bytes(big_slab, max_rl, big_k_cap, cta_bm_log2, cta_n_part, cta_win, false) > 224 * 1024)
        big_slab >>= 1;
      for (int level = 0; level < 3 && n_work > 0 && !(h3[2] & 1u); level++) {
        if (level == 1 && (big_slab <= cta_cap || getenv("SHRIMP_SCAN_CTA_CAP"))) continue;
        level_work[level] = n_work;
        P.work = have_list ? lists[cur] : nullptr;
        P.n_work = n_work;
        P.overflow = level < 2 ? lists[cur ^ 1] : nullptr;
        P.n_overflow = cnt + 4;
        P.work_counter = cnt + 3;
        SH_CUDA(cudaMemsetAsync(cnt + 3, 0, 8, st));
        int ctas, threads;
        if (level < 2) {
          P.cap = level == 0 ? cta_cap : big_slab;
          P.g_ent = nullptr;
          const size_t smem = scan_cta_smem_bytes(P.cap, max_rl, big_k_cap, cta_bm_log2, cta_n_part, cta_win, false) + 1024;
          const int per_sm = (int)std::max<size_t>(1, std::min<size_t>((size_t)(227 * 1024) / smem, 8));
          threads = std::max(128, std::min(768, (1536 / per_sm) & ~31));
          if (const char *e = getenv("SHRIMP_SCAN_CTA_THREADS")) threads = std::max(32, std::min(768, atoi(e) & ~31));
          ctas = (int)std::min<uint32_t>((uint32_t)(ctx->sm_count * per_sm), n_work);
        } else {
          ctas = (int)std::min<uint32_t>((uint32_t)(ctx->sm_count * glob_per_sm), n_work);
          threads = 512;
          const size_t per_cta = (size_t)(g_cap + 1) * 8 + (size_t)(g_cap + 1) * sizeof(AnchorRec) +
                                 (size_t)((g_cap + 15) & ~7) * 2 + (size_t)(g_cap / 32 + 2) * 4;
          SH_TRY(pl->d_scan_slab.ensure(per_cta * (size_t)ctas + 64));
          unsigned char *base = pl->d_scan_slab.as<unsigned char>();
          P.g_ent = (unsigned long long *)base;
          base += (size_t)ctas * (g_cap + 1) * 8;
          P.g_rec = (AnchorRec *)base;
          base += (size_t)ctas * (g_cap + 1) * sizeof(AnchorRec);
          P.g_keep = (uint32_t *)base;
          base += (size_t)ctas * (g_cap / 32 + 2) * 4;
          P.g_order = (uint16_t *)base;
          P.g_cap = g_cap;
          C.scan_global = n_work;
        }
        SH_TRY(launch_scan_cta(ctx, P, ctas, threads));
        uint32_t h5[5] = {0, 0, 0, 0, 0};
        SH_CUDA(cudaMemcpyAsync(h5, cnt, 20, cudaMemcpyDeviceToHost, st));
        SH_CUDA(cudaStreamSynchronize(st));
        h3[0] = h5[0];
        h3[2] = h5[2];
        n_work = h5[4];
        cur ^= 1;
        have_list = true;
      }
      if (P.prof) {
        unsigned long long hp[16];
        SH_CUDA(cudaMemcpyAsync(hp, P.prof, sizeof(hp), cudaMemcpyDeviceToHost, st));
        SH_CUDA(cudaStreamSynchronize(st));
        unsigned long long tot = 0;
        for (int i = 0; i < 16; i++) tot += hp[i];
        fprintf(stderr, "scan_cta phases (%% of CTA cycles):");
        for (int i = 0; i < 16; i++) if (hp[i]) fprintf(stderr, " [%d] %.1f", i, 100.0 * (double)hp[i] / (double)tot);
        fprintf(stderr, "\n");
        P.prof = nullptr;
      }
      // parked strands: replay the reference's heap order (a warp each), then resume them at the anchor step
      uint32_t h8[8];
      SH_CUDA(cudaMemcpyAsync(h8, cnt, 32, cudaMemcpyDeviceToHost, st));
      SH_CUDA(cudaStreamSynchronize(st));
      if ((h8[2] & 4u) && !(h8[2] & 3u)) {  // tie slab too small: the cursor counted what is needed
        pl->tie_cap = (uint32_t)std::min<unsigned long long>(0xfffffff0ull, (unsigned long long)h8[6] + h8[6] / 4 + 1024);
        if (attempt > 8) {
          set_error("seed scan: tie slab overflow");
          return SHRIMP_E_NOMEM;
        }
        continue;
      }
      if (h8[5] > 0 && !(h8[2] & 3u)) {
        const int ks_cap = std::max(32, K_slots);
        SH_TRY(launch_scan_replay(ctx, P, h8[5], ks_cap));
        P.resume = 1;
        P.work = nullptr;
        P.n_work = h8[5];
        P.overflow = nullptr;
        for (int level = 0; level < 3; level++) {
          if (level_work[level] == 0) continue;
          int ctas, threads;
          SH_CUDA(cudaMemsetAsync(cnt + 3, 0, 4, st));
          if (level < 2) {
            P.cap = level == 0 ? cta_cap : big_slab;
            P.resume_min = level == 0 ? 0 : cta_cap;
            P.g_ent = nullptr;
            const size_t smem = scan_cta_smem_bytes(P.cap, max_rl, big_k_cap, cta_bm_log2, cta_n_part, cta_win, false) + 1024;
            const int per_sm = (int)std::max<size_t>(1, std::min<size_t>((size_t)(227 * 1024) / smem, 8));
            threads = std::max(128, std::min(768, (1536 / per_sm) & ~31));
            ctas = (int)std::min<uint32_t>((uint32_t)(ctx->sm_count * per_sm), P.n_work);
          } else {
            // the global slabs of the level-2 launch above
            P.resume_min = level_work[1] ? big_slab : cta_cap;
            ctas = (int)std::min<uint32_t>((uint32_t)(ctx->sm_count * glob_per_sm), level_work[2]);
            threads = 512;
            unsigned char *base = pl->d_scan_slab.as<unsigned char>();
            P.g_ent = (unsigned long long *)base;
            base += (size_t)ctas * (g_cap + 1) * 8;
            P.g_rec = (AnchorRec *)base;
            base += (size_t)ctas * (g_cap + 1) * sizeof(AnchorRec);
            P.g_keep = (uint32_t *)base;
            base += (size_t)ctas * (g_cap / 32 + 2) * 4;
            P.g_order = (uint16_t *)base;
            P.g_cap = g_cap;
          }
          SH_TRY(launch_scan_cta(ctx, P, ctas, threads));
        }
        SH_CUDA(cudaMemcpyAsync(h3, cnt, 12, cudaMemcpyDeviceToHost, st));
        SH_CUDA(cudaStreamSynchronize(st));
      }
    }
    if (h3[2] & 2u) {
      set_error("seed scan: a read strand kept more than %d index positions after the region filter", g_cap);
      return SHRIMP_E_RANGE;
    }
    if (h3[2] & 1u) {  // hit buffer too small: grow and redo the scan
      if (attempt > 8) {
        set_error("seed scan: hit buffer overflow");
        return SHRIMP_E_NOMEM;
      }
      // hits_used kept counting past the capacity: it is the exact demand of this chunk (unless it wrapped)
      const unsigned long long need = h3[0] > pl->hits_cap ? (unsigned long long)h3[0] + h3[0] / 16 + 1024
                                                           : (unsigned long long)pl->hits_cap * 4;
      if (need > 0xfffffff0ull) {
        set_error("seed scan: more than 2^32 candidate windows in one chunk; map fewer reads per call");
        return SHRIMP_E_RANGE;
      }
      pl->hits_cap = (uint32_t)need;
      continue;
    }
    C.hits_used = h3[0];
    break;
  }
  return SHRIMP_OK;
}

// sw_vector over every eligible window (matches >= min_matches): true scores per hit slot in d_vtrue[0]
int chunk_vector(Chunk &C) {
  shrimp_gpu_ctx *ctx = C.ctx;
  Pipeline *pl = C.pl;
  cudaStream_t st = ctx->stream;
  const bool cs = C.cs;
  const size_t HU = std::max<uint32_t>(C.hits_used, 1);
  const size_t task_bytes = HU * 4 * 5 + ((HU + 3) & ~(size_t)3);
  VecTaskArrays VT[2];
  for (int o = 0; o < C.n_ori; o++) SH_TRY(pl->d_task[o].ensure(task_bytes));
  SH_TRY(pl->d_vtrue[0].ensure(HU * 4));
  SH_CUDA(cudaMemsetAsync(pl->d_vtrue[0].p, 0xff, HU * 4, st));
  SH_TRY(pl->d_slot.ensure(HU * 4));
  SH_TRY(pl->d_writer.ensure(HU));
  uint32_t n_dense[2] = {0, 0};
  {
    ScopedStage ss(ctx, ST_PASS1);
    TaskBuildParams TB;
    memset(&TB, 0, sizeof(TB));
    TB.G = C.G;
    TB.M = C.M;
    TB.hits = pl->d_hits.as<DevHit>();
    TB.rs_range = pl->d_rs_range.as<uint2>();
    TB.read_len = pl->d_read_len.as<int32_t>();
    TB.n_reads = C.n_reads;
    for (int o = 0; o < 2; o++) {
      char *tb = (char *)pl->d_task[o < C.n_ori ? o : 0].p;
      TB.goff[o] = (uint32_t *)tb;
      TB.glen[o] = (int32_t *)(tb + HU * 4);
      TB.ridx[o] = (int32_t *)(tb + HU * 8);
      TB.rlen[o] = (int32_t *)(tb + HU * 12);
      TB.out[o] = (uint32_t *)(tb + HU * 16);
      TB.initbp_out[o] = (int8_t *)(tb + HU * 20);
      VT[o].goff = TB.goff[o];
      VT[o].glen = TB.glen[o];
      VT[o].ridx = TB.ridx[o];
      VT[o].rlen = TB.rlen[o];
      VT[o].out = TB.out[o];
      VT[o].initbp = cs ? TB.initbp_out[o] : nullptr;
    }
    TB.initbp = cs ? pl->d_initbp.as<int8_t>() : nullptr;
    TB.slot = C.M.hash_filter_calls ? pl->d_slot.as<uint32_t>() : nullptr;
    TB.task_stats = C.cnt + 20;
    SH_TRY(launch_build_vec_tasks(ctx, TB));
    SH_CUDA(cudaMemcpyAsync(n_dense, C.cnt + 24, 8, cudaMemcpyDeviceToHost, st));
    if (TB.slot) SH_CUDA(cudaMemsetAsync(TB.slot, 0xff, (size_t)HU * 4, st));   // hits without a task: no cache slot
    SH_CUDA(cudaStreamSynchronize(st));
    if (TB.slot)
      for (int o = 0; o < C.n_ori; o++)
        SH_TRY(launch_window_slots(ctx, cs ? (o ? C.G.cs_rc : C.G.cs) : C.G.ls, VT[o].goff, (const int32_t *)VT[o].glen,
                                   VT[o].out, (uint32_t)n_dense[o], TB.slot));
  }
  {
    ScopedStage ss(ctx, ST_VECTOR);
    for (int o = 0; o < C.n_ori; o++) {
      if (n_dense[o] == 0) continue;
      const uint32_t *gen = cs ? (o ? C.G.cs_rc : C.G.cs) : C.G.ls;
      const uint32_t *gen_ls = cs ? (o ? C.G.ls_rc : C.G.ls) : nullptr;
      SH_TRY(launch_sw_vector(ctx, gen, gen_ls, pl->d_reads.as<uint32_t>(), C.stride, (int)n_dense[o], C.max_rl,
                              C.max_wl, VT[o], pl->d_vtrue[0].as<int32_t>(), ST_VECTOR));
    }
    if (C.M.gapless) {
      // -U / mirna: pass 1 ranks by sw_gapless; the sw_vector scores above still feed hit_run_full_sw (:386)
      SH_TRY(pl->d_vtrue[1].ensure(HU * 4));
      SH_CUDA(cudaMemsetAsync(pl->d_vtrue[1].p, 0xff, HU * 4, st));
      for (int o = 0; o < C.n_ori; o++) {
        GaplessParams GP;
        memset(&GP, 0, sizeof(GP));
        GP.G = C.G;
        GP.hits = pl->d_hits.as<DevHit>();
        GP.reads = pl->d_reads.as<uint32_t>();
        GP.stride = C.stride;
        GP.out = VT[o].out;
        GP.ridx = VT[o].ridx;
        GP.rlen = VT[o].rlen;
        GP.initbp = VT[o].initbp;
        GP.n_tasks = n_dense[o];
        GP.match = ctx->sw.match;
        GP.mismatch = cs ? ctx->sw.vec_mismatch : ctx->sw.mismatch;
        GP.cs = cs ? 1 : 0;
        GP.ori = o;
        GP.scores = pl->d_vtrue[1].as<int32_t>();
        SH_TRY(launch_sw_gapless(ctx, GP));
      }
    }
  }
  return SHRIMP_OK;
}

Pass1Params chunk_pass1_params(Chunk &C) {
  Pipeline *pl = C.pl;
  Pass1Params PP;
  memset(&PP, 0, sizeof(PP));
  PP.M = C.M;
  PP.hits = pl->d_hits.as<DevHit>();
  PP.rs_range = pl->d_rs_range.as<uint2>();
  PP.read_len = pl->d_read_len.as<int32_t>();
  PP.n_reads = C.n_reads;
  // every hit has one orientation: both sw_vector launches scatter into one array; gapless mode ranks by sw_gapless
  PP.vtrue[0] = PP.vtrue[1] = (C.M.gapless ? pl->d_vtrue[1] : pl->d_vtrue[0]).as<int32_t>();
  PP.slot = pl->d_slot.as<uint32_t>();
  PP.writer = pl->d_writer.as<uint8_t>();
  PP.sel = pl->d_sel.as<int32_t>();
  PP.n_sel = pl->d_nsel.as<int32_t>();
  PP.stats = C.cnt + 8;
  return PP;
}

// dense full-SW task list of the hits selected per read (d_sel / d_nsel): exclusive scan (cub, plumbing) ->
// task offsets -> one FullTask per selected hit, thresholds from full_thr (abs_or_pct form)
int chunk_full_tasks_unpaired(Chunk &C, double full_thr, int *n_slots_out) {
  shrimp_gpu_ctx *ctx = C.ctx;
  Pipeline *pl = C.pl;
  cudaStream_t st = ctx->stream;
  const int n_reads = C.n_reads, NT = C.mp->num_tmp_outputs;
  SH_TRY(pl->d_taskoff.ensure(((size_t)n_reads + 1) * 4));
  int n_slots = 0;
  size_t tmp_bytes = 0;
  SH_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, pl->d_nsel.as<int32_t>(), pl->d_taskoff.as<int32_t>(),
                                        n_reads + 1, st));
  SH_TRY(pl->d_scan_tmp.ensure(tmp_bytes));
  // n_sel has n_reads entries; entry n_reads of the scan needs a readable (zero) input slot
  SH_CUDA(cudaMemsetAsync(pl->d_nsel.as<int32_t>() + n_reads, 0, 4, st));
  SH_CUDA(cub::DeviceScan::ExclusiveSum(pl->d_scan_tmp.p, tmp_bytes, pl->d_nsel.as<int32_t>(),
                                        pl->d_taskoff.as<int32_t>(), n_reads + 1, st));
  ctx->launches += 1;
  SH_CUDA(cudaMemcpyAsync(&n_slots, pl->d_taskoff.as<int32_t>() + n_reads, 4, cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaStreamSynchronize(st));
  const int n_grid = n_reads * NT;
  SH_TRY(pl->d_ftasks.ensure((size_t)std::max(n_slots, 1) * sizeof(FullTask)));
  SH_TRY(pl->d_finfo.ensure((size_t)std::max(n_slots, 1) * sizeof(SelInfo)));
  SH_TRY(pl->d_fresults.ensure((size_t)std::max(n_slots, 1) * sizeof(FullResult)));
  ScopedStage ss(ctx, ST_FULL);
  FullBuildParams FB;
  memset(&FB, 0, sizeof(FB));
  FB.G = C.G;
  FB.M = C.M;
  FB.M.full_thr = full_thr;
  FB.M.full_frac = full_thr / 100.0;
  FB.hits = pl->d_hits.as<DevHit>();
  FB.rs_range = pl->d_rs_range.as<uint2>();
  FB.read_len = pl->d_read_len.as<int32_t>();
  FB.sel = pl->d_sel.as<int32_t>();
  FB.n_sel = pl->d_nsel.as<int32_t>();
  FB.vtrue0 = pl->d_vtrue[0].as<int32_t>();
  FB.initbp = C.cs ? pl->d_initbp.as<int8_t>() : nullptr;
  FB.task_off = pl->d_taskoff.as<int32_t>();
  FB.n_reads = n_reads;
  FB.tasks = pl->d_ftasks.as<FullTask>();
  FB.info = pl->d_finfo.as<SelInfo>();
  build_full_tasks_kernel<<<(n_grid + 127) / 128, 128, 0, st>>>(FB);
  SH_CUDA(cudaGetLastError());
  SH_LAUNCHED(ctx, ST_FULL);
  *n_slots_out = n_slots;
  return SHRIMP_OK;
}

// full SW with traceback over d_ftasks[0, n_slots)
int chunk_run_full(Chunk &C, int n_slots) {
  shrimp_gpu_ctx *ctx = C.ctx;
  Pipeline *pl = C.pl;
  const SwScores &sw = ctx->sw;
  SH_TRY(pl->d_fops.ensure(C.ops_stride * (size_t)std::max(n_slots, 1)));
  FullParams FP;
  memset(&FP, 0, sizeof(FP));
  FP.genome_fwd = C.G.ls;
  FP.genome_rc = C.G.ls_rc;
  FP.reads = pl->d_reads.as<uint32_t>();
  FP.stride = C.stride;
  FP.tasks = pl->d_ftasks.as<FullTask>();
  FP.results = pl->d_fresults.as<FullResult>();
  FP.ops = pl->d_fops.as<uint8_t>();
  FP.max_glen = C.max_wl;
  FP.max_rlen = C.max_rl;
  FP.match = sw.match;
  FP.mismatch = sw.mismatch;
  FP.a_open = sw.a_open;
  FP.a_ext = sw.a_ext;
  FP.b_open = sw.b_open;
  FP.b_ext = sw.b_ext;
  FP.anchor_width = sw.anchor_width;
  FP.Tflag = C.mp->Tflag;
  FP.local = C.mp->Gflag ? 0 : 1;
  FP.cells = (unsigned long long *)(C.cnt + 16);
  FP.xover = sw.xover;
  FP.xover_pos = pl->xover_stride ? pl->d_xover.as<int16_t>() : nullptr;
  FP.xover_stride = pl->xover_stride;
  FP.indel_taboo_len = sw.indel_taboo_len;
  {
    ScopedStage ss(ctx, ST_FULL);
    SH_TRY(run_full_sw(ctx, pl->d_perm, pl->d_frow, pl->d_fbp, FP, n_slots, C.cs, C.cnt + 32));
  }
  if (C.post_sw && n_slots > 0) {
    ScopedStage ss(ctx, ST_POST);
    // hit_run_post_sw (mapping.c:1609-1625) for every alignment with a positive score: emission terms from the host
    // (libm, gmapper.c:2561-2572 and post_sw_setup), recurrences on the device (post_sw.cu)
    const shrimp_map_params *mp = C.mp;
    const double pr_xover = mp->pr_xover > 0 ? mp->pr_xover : 0.03;
    const double alpha = mp->score_alpha, beta = mp->score_beta;
    const double pr_snp = 1.0 / (1.0 + 1.0 / 3.0 * pow(2.0, ((double)sw.match - (double)sw.mismatch) / alpha));
    PostParams PS;
    memset(&PS, 0, sizeof(PS));
    PS.genome_fwd = C.G.ls;
    PS.genome_rc = C.G.ls_rc;
    PS.reads = pl->d_reads.as<uint32_t>();
    PS.stride = C.stride;
    PS.tasks = pl->d_ftasks.as<FullTask>();
    PS.results = pl->d_fresults.as<FullResult>();
    PS.ops = pl->d_fops.as<uint8_t>();
    PS.ops_stride = (int)C.ops_stride;
    SH_TRY(pl->d_fqual.ensure((size_t)n_slots * (size_t)C.max_rl));
    PS.quals_out = pl->d_fqual.as<uint8_t>();
    PS.max_rlen = C.max_rl;
    PS.n_tasks = n_slots;
    PS.columns = (unsigned long long *)(C.cnt + 46);
    PS.score_alpha = alpha;
    PS.score_2ab = 2.0 * alpha + beta;
    PS.log2v = log(2.0);
    PS.la1 = log(1 - pr_snp);
    PS.la2 = log(pr_snp / 3.0);
    PS.lc1 = log(1 - pr_xover);
    PS.lc2 = log(pr_xover / 3.0);
    PS.ln1 = log(1 - .75);
    PS.ln2 = log(.75 / 3.0);
    PS.pr_del_open = pow(2.0, (double)(-sw.a_open) / alpha);      // the CLI (negative) gap scores
    PS.pr_ins_open = pow(2.0, (double)(-sw.b_open) / alpha);
    PS.pr_del_extend = pow(2.0, (double)(-sw.a_ext) / alpha);
    PS.pr_ins_extend = pow(2.0, ((double)(-sw.b_ext) - beta) / alpha);
    {   // tables of the libm transcription, once per context
      if (!pl->gm_tab_ready) {
        std::vector<unsigned long long> gm(8 + 256 + 18 + 256);
        memcpy(gm.data(), GLIBC_EXP_CONST, 8 * 8);
        memcpy(gm.data() + 8, GLIBC_EXP_TAB, 256 * 8);
        memcpy(gm.data() + 8 + 256, GLIBC_LOG_CONST, 18 * 8);
        memcpy(gm.data() + 8 + 256 + 18, GLIBC_LOG_TAB, 256 * 8);
        SH_TRY(pl->d_gmtab.ensure(gm.size() * 8));
        SH_CUDA(cudaMemcpy(pl->d_gmtab.p, gm.data(), gm.size() * 8, cudaMemcpyHostToDevice));
        pl->gm_tab_ready = true;
      }
      PS.gm_tab = pl->d_gmtab.as<unsigned long long>();
    }
    if (pl->qual_stride) {
      double tab[512];
      for (int q = 0; q < 256; q++) {
        const int qv = q - mp->qual_delta;
        double e = qv <= 0 ? .99999999 : qv >= 250 ? 1E-25 : pow(10.0, -(double)qv / 10.0);   // util.h:285-293
        if (!mp->use_sanger_qvs) e /= (1 + e);
        if (e > .75) e = .75;
        tab[q] = log(1 - e);
        tab[256 + q] = log(e / 3.0);
      }
      SH_TRY(pl->d_pstab.ensure(sizeof(tab)));
      SH_CUDA(cudaMemcpyAsync(pl->d_pstab.p, tab, sizeof(tab), cudaMemcpyHostToDevice, ctx->stream));
      SH_CUDA(cudaStreamSynchronize(ctx->stream));   // tab lives on this stack frame
      PS.lc1_tab = pl->d_pstab.as<double>();
      PS.lc2_tab = PS.lc1_tab + 256;
      PS.read_quals = pl->d_quals.as<uint8_t>();
      PS.qual_stride = pl->qual_stride;
      PS.qual_vector_offset = mp->qual_vector_offset;
    }
    SH_TRY(launch_post_sw(ctx, PS, pl->d_scratch));   // the scan scratch is dead by now
  }
  return SHRIMP_OK;
}

// results of the n_slots full-SW tasks (+ n_sel per read and the counters) back to pinned host memory
int chunk_fetch_full(Chunk &C, int n_slots, bool with_nsel) {
  Pipeline *pl = C.pl;
  cudaStream_t st = C.ctx->stream;
  const int n_reads = C.n_reads;
  pl->d2h_bytes += (size_t)n_slots * (sizeof(SelInfo) + sizeof(FullResult)) + C.ops_stride * (size_t)n_slots +
                   (with_nsel ? (size_t)n_reads * 4 : 0) + 64 * 4;
  SH_TRY(pl->h_info.ensure((size_t)n_slots * sizeof(SelInfo)));
  SH_TRY(pl->h_results.ensure((size_t)n_slots * sizeof(FullResult)));
  SH_TRY(pl->h_ops.ensure(C.ops_stride * (size_t)n_slots));
  SH_TRY(pl->h_nsel.ensure((size_t)n_reads * 4 + 64 * 4));
  SH_CUDA(cudaMemcpyAsync(pl->h_info.p, pl->d_finfo.p, (size_t)n_slots * sizeof(SelInfo), cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaMemcpyAsync(pl->h_results.p, pl->d_fresults.p, (size_t)n_slots * sizeof(FullResult),
                          cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaMemcpyAsync(pl->h_ops.p, pl->d_fops.p, C.ops_stride * (size_t)n_slots, cudaMemcpyDeviceToHost, st));
  if (C.post_sw && n_slots > 0) {
    SH_TRY(pl->h_fqual.ensure((size_t)n_slots * (size_t)C.max_rl));
    SH_CUDA(cudaMemcpyAsync(pl->h_fqual.p, pl->d_fqual.p, (size_t)n_slots * (size_t)C.max_rl, cudaMemcpyDeviceToHost, st));
    pl->d2h_bytes += (size_t)n_slots * (size_t)C.max_rl;
  }
  if (with_nsel)
    SH_CUDA(cudaMemcpyAsync(pl->h_nsel.p, pl->d_nsel.p, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, st));
  uint32_t *h_cnt = (uint32_t *)((char *)pl->h_nsel.p + (size_t)n_reads * 4);
  SH_CUDA(cudaMemcpyAsync(h_cnt, C.cnt, 64 * 4, cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaStreamSynchronize(st));
  return SHRIMP_OK;
}

// Results of the tasks [0, n_slots) of the unpaired flow (tasks of read r at d_taskoff[r] ..): compacted on the device
// (see pack_flag_kernel), then to the host.  h_nsel receives the number of alignments per read, h_info / h_results /
// h_ops the packed records; C.packed tells the host stage how to find an edit script.
int chunk_fetch_full_packed(Chunk &C, int n_slots) {
  Pipeline *pl = C.pl;
  shrimp_gpu_ctx *ctx = C.ctx;
  cudaStream_t st = ctx->stream;
  const int n_reads = C.n_reads;
  const size_t n1 = (size_t)n_slots + 1;
  SH_TRY(pl->d_pk.ensure(n1 * 4 * 4 + (size_t)n_reads * 4));
  uint32_t *keep = pl->d_pk.as<uint32_t>(), *units = keep + n1, *keep_pos = units + n1, *unit_off = keep_pos + n1;
  int32_t *d_nkept = (int32_t *)(unit_off + n1);
  SH_TRY(pl->h_nsel.ensure((size_t)n_reads * 4 + 64 * 4));
  uint32_t *h_cnt = (uint32_t *)((char *)pl->h_nsel.p + (size_t)n_reads * 4);
  uint32_t tot[2] = {0, 0};
  {
    ScopedStage ss(ctx, ST_OTHER);
    SH_CUDA(cudaMemsetAsync(C.cnt + 52, 0, 8, st));
    SH_CUDA(cudaMemsetAsync(keep + n_slots, 0, 4, st));
    SH_CUDA(cudaMemsetAsync(units + n_slots, 0, 4, st));
    if (n_slots > 0) {
      pack_flag_kernel<<<(n_slots + 255) / 256, 256, 0, st>>>(pl->d_fresults.as<FullResult>(), pl->d_finfo.as<SelInfo>(),
                                                              pl->d_read_len.as<int32_t>(), n_slots, C.post_sw ? 1 : 0,
                                                              C.cs ? 1 : 0, keep, units, (unsigned long long *)(C.cnt + 52));
      SH_CUDA(cudaGetLastError());
      SH_LAUNCHED(ctx, ST_OTHER);
    }
    size_t tmp_bytes = 0;
    SH_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, keep, keep_pos, (int)n1, st));
    SH_TRY(pl->d_scan_tmp.ensure(tmp_bytes));
    SH_CUDA(cub::DeviceScan::ExclusiveSum(pl->d_scan_tmp.p, tmp_bytes, keep, keep_pos, (int)n1, st));
    SH_CUDA(cub::DeviceScan::ExclusiveSum(pl->d_scan_tmp.p, tmp_bytes, units, unit_off, (int)n1, st));
    ctx->launches += 2;
    SH_CUDA(cudaMemcpyAsync(&tot[0], keep_pos + n_slots, 4, cudaMemcpyDeviceToHost, st));
    SH_CUDA(cudaMemcpyAsync(&tot[1], unit_off + n_slots, 4, cudaMemcpyDeviceToHost, st));
    SH_CUDA(cudaStreamSynchronize(st));
  }
  const size_t n_kept = tot[0], pool_bytes = 4 * (size_t)tot[1];
  SH_TRY(pl->d_pk_info.ensure(std::max<size_t>(n_kept, 1) * sizeof(SelInfo)));
  SH_TRY(pl->d_pk_res.ensure(std::max<size_t>(n_kept, 1) * sizeof(FullResult)));
  SH_TRY(pl->d_pk_pool.ensure(std::max<size_t>(pool_bytes, 4)));
  SH_TRY(pl->h_info.ensure(std::max<size_t>(n_kept, 1) * sizeof(SelInfo)));
  SH_TRY(pl->h_results.ensure(std::max<size_t>(n_kept, 1) * sizeof(FullResult)));
  SH_TRY(pl->h_ops.ensure(std::max<size_t>(pool_bytes, 4)));
  {
    ScopedStage ss(ctx, ST_OTHER);
    if (n_slots > 0) {
      pack_scatter_kernel<<<(n_slots + 127) / 128, 128, 0, st>>>(
          pl->d_fresults.as<FullResult>(), pl->d_finfo.as<SelInfo>(), pl->d_fops.as<uint8_t>(), C.ops_stride,
          C.post_sw ? pl->d_fqual.as<uint8_t>() : nullptr, C.max_rl, keep_pos, unit_off, n_slots, C.post_sw ? 1 : 0,
          pl->d_pk_info.as<SelInfo>(), pl->d_pk_res.as<FullResult>(), pl->d_pk_pool.as<uint8_t>());
      SH_CUDA(cudaGetLastError());
      SH_LAUNCHED(ctx, ST_OTHER);
    }
    pack_counts_kernel<<<(n_reads + 255) / 256, 256, 0, st>>>(pl->d_taskoff.as<int32_t>(), keep_pos, n_reads, d_nkept);
    SH_CUDA(cudaGetLastError());
    SH_LAUNCHED(ctx, ST_OTHER);
  }
  SH_CUDA(cudaMemcpyAsync(pl->h_info.p, pl->d_pk_info.p, n_kept * sizeof(SelInfo), cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaMemcpyAsync(pl->h_results.p, pl->d_pk_res.p, n_kept * sizeof(FullResult), cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaMemcpyAsync(pl->h_ops.p, pl->d_pk_pool.p, pool_bytes, cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaMemcpyAsync(pl->h_nsel.p, d_nkept, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaMemcpyAsync(h_cnt, C.cnt, 64 * 4, cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaStreamSynchronize(st));
  pl->d2h_bytes += n_kept * (sizeof(SelInfo) + sizeof(FullResult)) + pool_bytes + (size_t)n_reads * 4 + 64 * 4 + 8;
  C.packed = true;
  C.packed_slots = n_slots;
  C.packed_kept = (int64_t)n_kept;
  return SHRIMP_OK;
}

void chunk_stats(const Chunk &C, const uint32_t *hc, shrimp_map_stats *stats) {
  if (!stats) return;
  stats->heap_replays = hc[8 + 0];
  stats->list_entries = *(const unsigned long long *)(hc + 40);
  stats->surviving_entries = *(const unsigned long long *)(hc + 42);
  stats->anchors = *(const unsigned long long *)(hc + 44);
  stats->hits = C.hits_used;
  stats->vector_tasks = hc[20];
  stats->device_vector_cells = *(const unsigned long long *)(hc + 22);
  stats->vector_calls = hc[8 + 4];
  stats->vector_bypassed = hc[8 + 5];
  stats->vector_cells = *(const unsigned long long *)(hc + 8 + 6);
  stats->full_cells = *(const unsigned long long *)(hc + 16);
  stats->post_sw_columns = *(const unsigned long long *)(hc + 46);
  stats->scan_big_strands = C.scan_big;
  stats->scan_global_strands = C.scan_global;
}

// hit_run_full_sw's scores + hit_run_post_sw (mapping.c:1609-1625, letter space) for task idx
void host_score_hit(const Chunk &C, int idx, HostHit &h) {
  const shrimp_map_params *mp = C.mp;
  h.info = C.pl->h_info.as<SelInfo>()[idx];
  h.res = C.pl->h_results.as<FullResult>()[idx];
  h.task_idx = idx;
  h.posterior = 0.0;
  h.score_full = h.res.score;
  h.pct_score_full = (1000 * 100 * h.score_full) / h.info.score_max;
  if (mp->compute_mapping_qualities && h.score_full > 0) {
    int ps;
    if (C.cs) {   // post_sw ran on the device; it also took the logarithm of the score below (FullResult::post_score)
      h.posterior = h.res.posterior;
      ps = h.res.post_score;
    } else {
      h.posterior = pow(2.0, ((double)h.res.score - (double)h.res.rmapped * (2.0 * mp->score_alpha + mp->score_beta)) /
                                 mp->score_alpha);
      ps = (int)rint(mp->score_alpha * log(h.posterior) / log(2.0) +
                     (double)h.res.rmapped * (2.0 * mp->score_alpha + mp->score_beta));
      if (ps < 0) ps = 0;
    }
    const int pct = (1000 * 100 * ps) / h.info.score_max;
    h.score_full = ps;
    h.pct_score_full = pct;
  }
}

void host_fill_hit(const Chunk &C, const HostHit &h, int r, HostOut &O) {
  if (O.n_out >= O.hits_cap) {
    O.hits_short = true;
    return;
  }
  shrimp_hit &o = O.hits[O.n_out++];
  o.read_idx = r;
  o.cn = h.info.cn;
  o.gen_st = h.info.gen_st;
  o.w_len = h.info.w_len;
  o.g_off = h.info.g_off;
  o.score_vector = h.info.score_vector;
  o.score_full = h.score_full;
  o.pass2_key = h.pass2_key;
  o.score_max = h.info.score_max;
  o.matches = h.info.matches;
  o.sw_score = h.res.score;
  o.posterior = h.posterior;
  o.read_start = h.res.read_start;
  o.rmapped = h.res.rmapped;
  o.genome_start = h.res.genome_start;
  o.gmapped = h.res.gmapped;
  o.sfr_matches = h.res.matches;
  o.mismatches = h.res.mismatches;
  o.insertions = h.res.insertions;
  o.deletions = h.res.deletions;
  o.crossovers = h.res.crossovers;
  o.edit_len = h.res.ops_len;
  o.edit_off = O.e_used;
  o.hit_slot = h.info.hit_slot;
  o.st = h.info.st;
  o.score_window_gen = h.info.wg;
  o.reserved = 0;
  const uint8_t *OPS = C.pl->h_ops.as<uint8_t>();
  if (O.edits && O.e_used + h.res.ops_len <= O.edits_cap)
    memcpy(O.edits + O.e_used, OPS + C.ops_stride * (size_t)h.task_idx + h.res.ops_start, (size_t)h.res.ops_len);   // never packed: pairs
  else if (h.res.ops_len > 0)
    O.edits_short = true;
  O.e_used += h.res.ops_len;
  if (C.post_sw) {   // sfrp->qual right after the edit script
    if (O.edits && O.e_used + h.res.rmapped <= O.edits_cap)
      memcpy(O.edits + O.e_used, C.pl->h_fqual.as<uint8_t>() + (size_t)C.max_rl * (size_t)h.task_idx, (size_t)h.res.rmapped);
    else if (h.res.rmapped > 0)
      O.edits_short = true;
    O.e_used += h.res.rmapped;
  }
}

// read_pass2 after the DP (mapping.c:1644-1722) for one read: tasks [task_base, task_base + n1) -> the hits it
// keeps, in output order, as (task, scores) records.  hh / h2 are caller-owned scratch of >= n1 entries.
struct KeptRec {
  int task_idx, score_full, pass2_key;
  double posterior;
};
static int host_pass2_select(const Chunk &C, int r, int n1, int task_base, double full_thr, HostHit *hh, HostHit **h2,
                             KeptRec *out, uint64_t &full_calls, uint64_t &vcalls, uint64_t &vcells) {
  const shrimp_map_params *mp = C.mp;
  int n2 = 0;
  for (int k = 0; k < n1; k++) {
    HostHit &h = hh[k];
    host_score_hit(C, task_base + k, h);
    if (!C.cs && !C.packed) {
      vcalls++;
      vcells += (uint64_t)h.info.w_len * (uint64_t)C.read_len[r];
    }
    if (!C.packed && (h.res.score > 0 || h.res.ops_len > 0)) full_calls++;
    h.pass2_key = full_thr < 0 ? h.score_full : (int)h.pct_score_full;
    const double thr = full_thr < 0 ? -full_thr : h.info.score_max * (full_thr / 100.0);
    if (h.score_full >= thr) h2[n2++] = &h;
  }
  if (n2 > 1) {  // with one survivor the duplicate removal and the ranking are the identity
    dedup_pass(h2, &n2, cmp_gen_start);
    dedup_pass(h2, &n2, cmp_gen_end);
    qsort(h2, n2, sizeof(HostHit *), cmp_score);
  }
  if (n2 > mp->num_outputs) n2 = mp->num_outputs;
  if (mp->strata && n2 > 0) {
    int i;
    for (i = 1; i < n2 && h2[0]->score_full == h2[i]->score_full; i++)
      ;
    n2 = i;
  }
  if (n2 > 0 && !(mp->max_alignments == 0 || n2 <= mp->max_alignments)) n2 = 0;
  for (int i = 0; i < n2; i++) {
    out[i].task_idx = h2[i]->task_idx;
    out[i].score_full = h2[i]->score_full;
    out[i].pass2_key = h2[i]->pass2_key;
    out[i].posterior = h2[i]->posterior;
  }
  return n2;
}

// host threads of the OpenMP stages; 0 = omp_get_max_threads() (launchers such as torchrun set OMP_NUM_THREADS=1)
static std::atomic<int> g_host_threads{0};
extern "C" int shrimp_gpu_set_host_threads(int n) {
  g_host_threads.store(n < 0 ? 0 : n);
  return SHRIMP_OK;
}

// read_pass2 for every read of the chunk (n_sel[r] tasks each, in task order), OpenMP over contiguous blocks
// of reads: the duplicate removal and ranking stay the reference's qsort-based code, only the
// loop over reads is parallel.  Appends to O in read order; n_per_read[r] = hits kept for read r.
int host_pass2_all(const Chunk &C, const int32_t *n_sel, double full_thr, HostOut &O, int32_t *n_per_read) {
  const int n_reads = C.n_reads, NT = C.mp->num_tmp_outputs, NO = C.mp->num_outputs;
  const int want_T = g_host_threads.load() > 0 ? g_host_threads.load() : omp_get_max_threads();
  const int T = std::max(1, std::min(want_T, 64));
  std::vector<std::vector<KeptRec>> kept((size_t)T);
  std::vector<std::vector<int32_t>> counts((size_t)T);
  std::vector<int64_t> nh((size_t)T + 1, 0), ne((size_t)T + 1, 0);
  std::vector<uint64_t> fc((size_t)T, 0), vc((size_t)T, 0), vl((size_t)T, 0);
  std::vector<int64_t> tb((size_t)T + 1, 0);  // first task of each block
  {
    int64_t acc = 0;
    int t = 0;
    for (int r = 0; r <= n_reads; r++) {
      while (t <= T && r == (int)((int64_t)n_reads * t / T)) tb[t++] = acc;
      if (r < n_reads) acc += n_sel[r];
    }
  }
  const FullResult *RES = C.pl->h_results.as<FullResult>();
  // the block index is the loop variable, never the team size: inside gmapper's own parallel region (nested
  // parallelism off) or under OMP_THREAD_LIMIT the team has fewer than T threads and one thread takes several blocks
#pragma omp parallel for schedule(static, 1) num_threads(T)
  for (int t = 0; t < T; t++) {
    const int r0 = (int)((int64_t)n_reads * t / T), r1 = (int)((int64_t)n_reads * (t + 1) / T);
    std::vector<HostHit> hh((size_t)NT + 1);
    std::vector<HostHit *> h2((size_t)NT + 1);
    std::vector<KeptRec> tmp((size_t)NO + 1);
    std::vector<KeptRec> &K = kept[t];
    K.reserve((size_t)(r1 - r0) * 2 + 64);
    counts[t].resize((size_t)(r1 - r0));
    int64_t task_base = tb[t], e = 0;
    uint64_t my_fc = 0, my_vc = 0, my_vl = 0;  // thread-local: the shared counters would share cache lines
    for (int r = r0; r < r1; r++) {
      const int n2 = host_pass2_select(C, r, n_sel[r], (int)task_base, full_thr, hh.data(), h2.data(), tmp.data(), my_fc,
                                       my_vc, my_vl);
      for (int i = 0; i < n2; i++) {
        K.push_back(tmp[i]);
        e += RES[tmp[i].task_idx].ops_len + (C.post_sw ? RES[tmp[i].task_idx].rmapped : 0);
      }
      counts[t][r - r0] = n2;
      task_base += n_sel[r];
    }
    nh[t + 1] = (int64_t)K.size();
    ne[t + 1] = e;
    fc[t] = my_fc;
    vc[t] = my_vc;
    vl[t] = my_vl;
  }
  for (int t = 0; t < T; t++) {
    nh[t + 1] += nh[t];
    ne[t + 1] += ne[t];
    O.full_calls += fc[t];
    O.pass2_vector_calls += vc[t];
    O.pass2_vector_cells += vl[t];
  }
  if (C.packed) {   // counted on the device over every task, kept or not (pack_flag_kernel)
    const uint32_t *hc = (const uint32_t *)((const char *)C.pl->h_nsel.p + (size_t)n_reads * 4);
    O.full_calls += (uint64_t)C.packed_kept;
    if (!C.cs) {
      O.pass2_vector_calls += (uint64_t)C.packed_slots;
      O.pass2_vector_cells += *(const unsigned long long *)(hc + 52);
    }
  }
  const int64_t hit_base = O.n_out, edit_base = O.e_used;
  if (hit_base + nh[T] > O.hits_cap) O.hits_short = true;
  if (O.edits && edit_base + ne[T] > O.edits_cap) O.edits_short = true;
  const bool fill = !O.hits_short;
  const bool fill_edits = O.edits && !O.edits_short;
  const SelInfo *INFO = C.pl->h_info.as<SelInfo>();
  const uint8_t *OPS = C.pl->h_ops.as<uint8_t>();
  const uint8_t *FQ = C.pl->h_fqual.as<uint8_t>();
#pragma omp parallel for schedule(static, 1) num_threads(T)
  for (int t = 0; t < T; t++) {
    const int r0 = (int)((int64_t)n_reads * t / T), r1 = (int)((int64_t)n_reads * (t + 1) / T);
    int64_t hi = hit_base + nh[t], eo = edit_base + ne[t];
    size_t q = 0;
    for (int r = r0; r < r1; r++) {
      const int n2 = counts[t][r - r0];
      if (n_per_read) n_per_read[r] = n2;
      for (int i = 0; i < n2 && fill; i++, q++, hi++) {
        const KeptRec &kr = kept[t][q];
        const SelInfo &info = INFO[kr.task_idx];
        const FullResult &res = RES[kr.task_idx];
        shrimp_hit &o = O.hits[hi];
        o.read_idx = r;
        o.cn = info.cn;
        o.gen_st = info.gen_st;
        o.w_len = info.w_len;
        o.g_off = info.g_off;
        o.score_vector = info.score_vector;
        o.score_full = kr.score_full;
        o.pass2_key = kr.pass2_key;
        o.score_max = info.score_max;
        o.matches = info.matches;
        o.sw_score = res.score;
        o.posterior = kr.posterior;
        o.read_start = res.read_start;
        o.rmapped = res.rmapped;
        o.genome_start = res.genome_start;
        o.gmapped = res.gmapped;
        o.sfr_matches = res.matches;
        o.mismatches = res.mismatches;
        o.insertions = res.insertions;
        o.deletions = res.deletions;
        o.crossovers = res.crossovers;
        o.edit_len = res.ops_len;
        o.edit_off = eo;
        o.hit_slot = info.hit_slot;
        o.st = info.st;
        o.score_window_gen = info.wg;
        o.reserved = 0;
        const uint8_t *rec = C.packed ? OPS + 4 * (size_t)(uint32_t)res.ops_start
                                      : OPS + C.ops_stride * (size_t)kr.task_idx + res.ops_start;
        if (fill_edits) memcpy(O.edits + eo, rec, (size_t)res.ops_len);
        eo += res.ops_len;
        if (C.post_sw) {   // sfrp->qual right after the edit script
          if (fill_edits)
            memcpy(O.edits + eo, C.packed ? rec + res.ops_len : FQ + (size_t)C.max_rl * (size_t)kr.task_idx,
                   (size_t)res.rmapped);
          eo += res.rmapped;
        }
      }
    }
  }
  O.n_out = hit_base + nh[T];
  O.e_used = edit_base + ne[T];
  return SHRIMP_OK;
}

}  // namespace shrimp

using namespace shrimp;

// resident = reuse the reads uploaded by the previous call (bench: inputs already in HBM);
// device_only = stop after the last device stage (no D2H of results, no host pass 2).
static int map_impl(shrimp_gpu_ctx *ctx, const shrimp_map_params *mp, int n_reads, const uint32_t *reads, int stride,
                    const int32_t *read_len, const int8_t *initbp, shrimp_hit *hits_out, int64_t hits_cap,
                    int32_t *n_hits_per_read, uint8_t *edits, int64_t edits_cap, int64_t *n_hits, int64_t *edits_used,
                    shrimp_stage_hit *stage, int64_t stage_cap, int64_t *n_stage, shrimp_map_stats *stats,
                    bool resident, bool device_only) {
  const char *who = resident ? "shrimp_gpu_map_resident" : "shrimp_gpu_map_reads";
  int64_t dummy_hits = 0;
  if (device_only && !n_hits) n_hits = &dummy_hits;
  if ((!hits_out && !device_only) || !n_hits) {
    set_error("%s: invalid argument", who);
    return SHRIMP_E_ARG;
  }
  if (mp && mp->match_mode != 1 && mp->match_mode != 2) {
    set_error("%s: unpaired match_mode must be 1 or 2", who);
    return SHRIMP_E_ARG;
  }
  Chunk C;
  const bool timing = getenv("SHRIMP_TIMING") != nullptr;
  const double t_b0 = omp_get_wtime();
  SH_TRY(chunk_begin(C, ctx, mp, n_reads, reads, stride, read_len, initbp, resident, who));
  n_reads = C.n_reads;
  read_len = C.read_len;
  *n_hits = 0;
  if (edits_used) *edits_used = 0;
  if (n_stage) *n_stage = 0;
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n_hits_per_read) memset(n_hits_per_read, 0, sizeof(int32_t) * (size_t)n_reads);
  if (n_reads == 0) return SHRIMP_OK;
  if (!device_only && hits_cap < (int64_t)n_reads * mp->num_outputs) {
    set_error("%s: hits_cap must be at least n_reads * num_outputs", who);
    return SHRIMP_E_ARG;
  }
  Pipeline *pl = C.pl;
  cudaStream_t st = ctx->stream;
  pl->d2h_bytes = 0;
  const double t_b1 = omp_get_wtime();
  SH_TRY(chunk_scan(C));
  const double t_b2 = omp_get_wtime();
  SH_TRY(chunk_vector(C));
  const double t_b3 = omp_get_wtime();

  // ---- pass-1 replay + top-k -------------------------------------------------------------------
  const int NT = mp->num_tmp_outputs;
  SH_TRY(pl->d_sel.ensure((size_t)n_reads * NT * 4));
  SH_TRY(pl->d_nsel.ensure(((size_t)n_reads + 1) * 4));
  {
    ScopedStage ss(ctx, ST_PASS1);
    Pass1Params PP = chunk_pass1_params(C);
    SH_TRY(launch_pass1_replay(ctx, PP));
    SH_TRY(launch_select_unpaired(ctx, PP));
  }

  // ---- stage dump (tests) ---------------------------------------------------------------------
  if (stage) {
    const size_t HU = std::max<uint32_t>(C.hits_used, 1);
    SH_TRY(pl->h_hits.ensure(HU * sizeof(DevHit)));
    SH_TRY(pl->h_range.ensure((size_t)n_reads * 2 * sizeof(uint2)));
    SH_CUDA(cudaMemcpyAsync(pl->h_hits.p, pl->d_hits.p, (size_t)C.hits_used * sizeof(DevHit), cudaMemcpyDeviceToHost, st));
    SH_CUDA(cudaMemcpyAsync(pl->h_range.p, pl->d_rs_range.p, (size_t)n_reads * 2 * sizeof(uint2),
                            cudaMemcpyDeviceToHost, st));
    SH_CUDA(cudaStreamSynchronize(st));
    const DevHit *H = pl->h_hits.as<DevHit>();
    const uint2 *RG = pl->h_range.as<uint2>();
    int64_t ns = 0;
    for (int rs = 0; rs < 2 * n_reads; rs++) {
      for (uint32_t k = 0; k < RG[rs].y; k++) {
        if (ns >= stage_cap) {
          set_error("%s: stage_cap too small", who);
          return SHRIMP_E_NOMEM;
        }
        const DevHit &h = H[RG[rs].x + k];
        shrimp_stage_hit &s = stage[ns++];
        s.read_idx = rs >> 1;
        s.st = rs & 1;
        s.cn = h.cn;
        s.w_len = h.w_len;
        s.g_off = h.g_off;
        s.score_window_gen = h.wg;
        s.matches = h.matches;
        s.score_max = h.score_max;
        s.score_vector = h.score_vector;
        s.pct_score_vector = h.pct_vector;
        s.ax = h.ax;
        s.ay = h.ay;
        s.alen = h.alen;
        s.awidth = h.awidth;
      }
    }
    if (n_stage) *n_stage = ns;
  }

  // ---- full SW on the selected hits ------------------------------------------------------------
  int n_slots = 0;
  SH_TRY(chunk_full_tasks_unpaired(C, mp->sw_full_threshold, &n_slots));
  SH_TRY(chunk_run_full(C, n_slots));

  if (device_only) {
    SH_TRY(pl->h_nsel.ensure((size_t)n_reads * 4 + 64 * 4));
    uint32_t *hc = (uint32_t *)((char *)pl->h_nsel.p + (size_t)n_reads * 4);
    SH_CUDA(cudaMemcpyAsync(hc, C.cnt, 64 * 4, cudaMemcpyDeviceToHost, st));
    SH_CUDA(cudaStreamSynchronize(st));
    chunk_stats(C, hc, stats);
    return SHRIMP_OK;
  }
  // ---- results back to the host + host stage: read_pass2 after the DP ----------------------------
  const double t_f0 = omp_get_wtime();
  SH_TRY(chunk_fetch_full_packed(C, n_slots));
  const double t_f1 = omp_get_wtime();
  const uint32_t *h_cnt = (const uint32_t *)((char *)pl->h_nsel.p + (size_t)n_reads * 4);
  const int32_t *NSEL = pl->h_nsel.as<int32_t>();
  HostOut O;
  memset(&O, 0, sizeof(O));
  O.hits = hits_out;
  O.hits_cap = hits_cap;
  O.edits = edits;
  O.edits_cap = edits_cap;
  SH_TRY(host_pass2_all(C, NSEL, mp->sw_full_threshold, O, n_hits_per_read));
  if (timing)
    fprintf(stderr, "[shrimp_b200] map_reads: %d reads, n_slots %d: begin %.2f, scan %.2f, vector %.2f, pass 1 + full SW %.2f, "
            "D2H %.2f ms (%.1f MB), host pass 2 %.2f ms\n", n_reads, n_slots, 1e3 * (t_b1 - t_b0), 1e3 * (t_b2 - t_b1),
            1e3 * (t_b3 - t_b2), 1e3 * (t_f0 - t_b3), 1e3 * (t_f1 - t_f0), pl->d2h_bytes / 1e6, 1e3 * (omp_get_wtime() - t_f1));
  *n_hits = O.n_out;
  if (edits_used) *edits_used = O.e_used;
  if (stats) {
    chunk_stats(C, h_cnt, stats);
    stats->vector_calls += O.pass2_vector_calls;
    stats->vector_cells += O.pass2_vector_cells;
    stats->full_calls = O.full_calls;
  }
  if (O.hits_short) {
    set_error("%s: hits_cap too small", who);
    return SHRIMP_E_NOMEM;
  }
  if (O.edits_short && edits) {
    set_error("%s: edits_cap too small, %lld bytes needed", who, (long long)O.e_used);
    return SHRIMP_E_NOMEM;
  }
  return SHRIMP_OK;
}

extern "C" int shrimp_gpu_map_reads(shrimp_gpu_ctx *ctx, const shrimp_map_params *mp, int n_reads,
                                    const uint32_t *reads, int stride, const int32_t *read_len, const int8_t *initbp,
                                    shrimp_hit *hits_out, int64_t hits_cap, int32_t *n_hits_per_read, uint8_t *edits,
                                    int64_t edits_cap, int64_t *n_hits, int64_t *edits_used, shrimp_stage_hit *stage,
                                    int64_t stage_cap, int64_t *n_stage, shrimp_map_stats *stats) {
  return map_impl(ctx, mp, n_reads, reads, stride, read_len, initbp, hits_out, hits_cap, n_hits_per_read, edits,
                  edits_cap, n_hits, edits_used, stage, stage_cap, n_stage, stats, false, false);
}

// Measurement entry: re-runs every device stage on the reads left in HBM by the previous
// shrimp_gpu_map_reads call; results stay on the device (bench.py's device-resident `value`).
extern "C" int shrimp_gpu_map_resident(shrimp_gpu_ctx *ctx, const shrimp_map_params *mp, shrimp_map_stats *stats) {
  return map_impl(ctx, mp, 0, nullptr, 0, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr,
                  0, nullptr, stats, true, true);
}

// Host<->device bytes moved by the last shrimp_gpu_map_reads call.
extern "C" int shrimp_gpu_last_transfer_bytes(shrimp_gpu_ctx *ctx, uint64_t *h2d, uint64_t *d2h) {
  Pipeline *pl = ctx ? (Pipeline *)ctx->pipeline : nullptr;
  if (!pl) {
    set_error("shrimp_gpu_last_transfer_bytes: no mapping call yet");
    return SHRIMP_E_STATE;
  }
  if (h2d) *h2d = pl->h2d_bytes;
  if (d2h) *d2h = pl->d2h_bytes;
  return SHRIMP_OK;
}

// Batched sw_full_ls / sw_full_cs (common/sw-full-ls.c:637-683, common/sw-full-cs.c:1146-1236): one
// task per call the reference would make, against one packed letter genome array.
extern "C" int shrimp_gpu_sw_full_batch(shrimp_gpu_ctx *ctx, const uint32_t *genome, size_t genome_words,
                                        const uint32_t *reads, int read_stride_words, int n_reads, int n_tasks,
                                        const shrimp_full_task *tasks, int local_alignment,
                                        shrimp_full_result *results, uint8_t *edits, int64_t edits_cap,
                                        int64_t *edits_used) {
  return shrimp_gpu_sw_full_batch_xover(ctx, genome, genome_words, reads, read_stride_words, n_reads, n_tasks, tasks,
                                        local_alignment, nullptr, 0, results, edits, edits_cap, edits_used);
}

// ... with the per-position crossover scores sw_full_cs takes as its last argument (sw-full-cs.c:1149): row
// read_idx of crossover_scores[n_reads][crossover_stride], or NULL for the global score of the set-up.
extern "C" int shrimp_gpu_sw_full_batch_xover(shrimp_gpu_ctx *ctx, const uint32_t *genome, size_t genome_words,
                                              const uint32_t *reads, int read_stride_words, int n_reads, int n_tasks,
                                              const shrimp_full_task *tasks, int local_alignment,
                                              const int32_t *crossover_scores, int crossover_stride,
                                              shrimp_full_result *results, uint8_t *edits, int64_t edits_cap,
                                              int64_t *edits_used) {
  if (!ctx || !genome || !reads || !tasks || !results || n_tasks < 0 || n_reads <= 0 || read_stride_words <= 0 ||
      (crossover_scores && crossover_stride <= 0)) {
    set_error("shrimp_gpu_sw_full_batch: invalid argument");
    return SHRIMP_E_ARG;
  }
  if (!ctx->sw.valid) {
    set_error("shrimp_gpu_sw_full_batch: shrimp_gpu_sw_setup() has not been called");
    return SHRIMP_E_STATE;
  }
  if (edits_used) *edits_used = 0;
  if (n_tasks == 0) return SHRIMP_OK;
  const SwScores &sw = ctx->sw;
  const bool cs = sw.use_colours != 0;
  int max_rl = 1, max_gl = 1;
  // the kernels find a read's crossover row at ridx >> 1 (the chunk layout: row 2r = read r, 2r + 1 = its reverse
  // complement), so with crossover rows the reads of the batch go to the even rows
  const bool xpos = cs && crossover_scores != nullptr && crossover_stride > 0;
  const int rmul = xpos ? 2 : 1;
  std::vector<FullTask> ft((size_t)n_tasks);
  for (int t = 0; t < n_tasks; t++) {
    const shrimp_full_task &a = tasks[t];
    if (xpos && a.rlen > crossover_stride) {
      set_error("shrimp_gpu_sw_full_batch: task %d is longer than a crossover row", t);
      return SHRIMP_E_ARG;
    }
    if (a.glen <= 0 || a.rlen <= 0 || a.read_idx < 0 || a.read_idx >= n_reads || a.rlen > read_stride_words * 8 ||
        (uint64_t)a.goff + (uint64_t)a.glen > (uint64_t)genome_words * 8 || a.rlen > sw.max_read_len ||
        a.glen > sw.max_window_len) {
      set_error("shrimp_gpu_sw_full_batch: task %d out of range", t);
      return SHRIMP_E_ARG;
    }
    FullTask &T = ft[t];
    memset(&T, 0, sizeof(T));
    T.goff_global = a.goff;
    T.goff_contig = a.goff;
    T.glen = a.glen;
    T.rlen = a.rlen;
    T.ridx = a.read_idx * rmul;
    T.ax = a.ax;
    T.ay = a.ay;
    T.alen = a.alen;
    T.awidth = a.awidth;
    T.thresh = a.threshscore;
    T.maxscore = a.maxscore;
    T.gen_st = a.revcmpl ? 1 : 0;
    T.run = 1;
    T.initbp = a.initbp;
    max_rl = std::max(max_rl, a.rlen);
    max_gl = std::max(max_gl, a.glen);
  }
  SH_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevBuf d_gen, d_reads, d_tasks, d_res, d_ops, d_perm, d_row, d_bp[RING_CLASSES + 2], d_cnt, d_xo;
  struct Rel {
    DevBuf *b[15];
    ~Rel() {
      for (DevBuf *x : b) x->release();
    }
  } rel{{&d_gen, &d_reads, &d_tasks, &d_res, &d_ops, &d_perm, &d_row, &d_bp[0], &d_bp[1], &d_bp[2], &d_bp[3], &d_bp[4],
         &d_bp[5], &d_cnt, &d_xo}};
  const size_t ops_stride = (size_t)max_rl + max_gl;
  SH_TRY(d_gen.ensure(genome_words * 4 + 16));
  SH_TRY(d_reads.ensure((size_t)n_reads * rmul * read_stride_words * 4));
  SH_TRY(d_tasks.ensure((size_t)n_tasks * sizeof(FullTask)));
  SH_TRY(d_res.ensure((size_t)n_tasks * sizeof(FullResult)));
  SH_TRY(d_ops.ensure(ops_stride * (size_t)n_tasks));
  SH_TRY(d_cnt.ensure(64 * 4));
  SH_CUDA(cudaMemsetAsync((char *)d_gen.p + genome_words * 4, 0, 16, st));
  SH_CUDA(cudaMemcpyAsync(d_gen.p, genome, genome_words * 4, cudaMemcpyHostToDevice, st));
  SH_CUDA(cudaMemcpy2DAsync(d_reads.p, (size_t)rmul * read_stride_words * 4, reads, (size_t)read_stride_words * 4,
                            (size_t)read_stride_words * 4, (size_t)n_reads, cudaMemcpyHostToDevice, st));
  std::vector<int16_t> hx;
  if (xpos) {
    hx.resize((size_t)n_reads * crossover_stride);
    for (size_t q = 0; q < hx.size(); q++) {
      const int32_t v = crossover_scores[q];
      hx[q] = (int16_t)(v < -32768 ? -32768 : v > 0 ? 0 : v);
    }
    SH_TRY(d_xo.ensure(hx.size() * 2));
    SH_CUDA(cudaMemcpyAsync(d_xo.p, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice, st));
  }
  SH_CUDA(cudaMemcpyAsync(d_tasks.p, ft.data(), (size_t)n_tasks * sizeof(FullTask), cudaMemcpyHostToDevice, st));
  SH_CUDA(cudaMemsetAsync(d_cnt.p, 0, 64 * 4, st));
  FullParams FP;
  memset(&FP, 0, sizeof(FP));
  FP.genome_fwd = FP.genome_rc = d_gen.as<uint32_t>();
  FP.reads = d_reads.as<uint32_t>();
  FP.stride = read_stride_words;
  FP.tasks = d_tasks.as<FullTask>();
  FP.results = d_res.as<FullResult>();
  FP.ops = d_ops.as<uint8_t>();
  FP.max_glen = max_gl;
  FP.max_rlen = max_rl;
  FP.match = sw.match;
  FP.mismatch = sw.mismatch;
  FP.a_open = sw.a_open;
  FP.a_ext = sw.a_ext;
  FP.b_open = sw.b_open;
  FP.b_ext = sw.b_ext;
  FP.anchor_width = sw.anchor_width;
  FP.Tflag = 1;
  FP.local = local_alignment ? 1 : 0;
  FP.cells = (unsigned long long *)(d_cnt.as<uint32_t>() + 16);
  FP.xover = sw.xover;
  FP.indel_taboo_len = sw.indel_taboo_len;
  FP.xover_pos = xpos ? d_xo.as<int16_t>() : nullptr;
  FP.xover_stride = xpos ? crossover_stride : 0;
  {
    ScopedStage ss(ctx, ST_FULL);
    SH_TRY(run_full_sw(ctx, d_perm, d_row, d_bp, FP, n_tasks, cs, d_cnt.as<uint32_t>() + 32));
  }
  std::vector<FullResult> hr((size_t)n_tasks);
  std::vector<uint8_t> hops(ops_stride * (size_t)n_tasks);
  SH_CUDA(cudaMemcpyAsync(hr.data(), d_res.p, (size_t)n_tasks * sizeof(FullResult), cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaMemcpyAsync(hops.data(), d_ops.p, hops.size(), cudaMemcpyDeviceToHost, st));
  SH_CUDA(cudaStreamSynchronize(st));
  int64_t used = 0;
  bool short_pool = false;
  for (int t = 0; t < n_tasks; t++) {
    const FullResult &r = hr[t];
    shrimp_full_result &o = results[t];
    o.score = r.score;
    o.read_start = r.read_start;
    o.rmapped = r.rmapped;
    o.genome_start = r.genome_start;
    o.gmapped = r.gmapped;
    o.matches = r.matches;
    o.mismatches = r.mismatches;
    o.insertions = r.insertions;
    o.deletions = r.deletions;
    o.crossovers = r.crossovers;
    o.edit_len = r.ops_len;
    o.edit_off = used;
    if (edits && used + r.ops_len <= edits_cap)
      memcpy(edits + used, hops.data() + ops_stride * (size_t)t + r.ops_start, (size_t)r.ops_len);
    else if (r.ops_len > 0)
      short_pool = true;
    used += r.ops_len;
  }
  if (edits_used) *edits_used = used;
  if (short_pool && edits) {
    set_error("shrimp_gpu_sw_full_batch: edits_cap too small, %lld bytes needed", (long long)used);
    return SHRIMP_E_NOMEM;
  }
  return SHRIMP_OK;
}
