"""Host-side mirror of the reference's operator interface for the hot path.

Names, argument meaning and error behaviour follow the reference (compbio-UofT/shrimp 2.2.3):
``sw_vector_setup``/``sw_vector`` (common/sw-vector.c:388,453) etc., but every call takes a batch
because a GPU call per window would be pointless.  All arithmetic happens in libshrimp_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

import math

from ._lib import (FullResultC, FullTaskC, HitC, MapParamsC, MapStatsC, PairC, PairParamsC, StageHitC, SwParams,
                   check, lib)

# fasta.h:26-42
_LS_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate("ACGTUMRWSYKVHDB"):
    _LS_CODE[ord(_c)] = _i
    _LS_CODE[ord(_c.lower())] = _i
for _c in "NnXx.-":
    _LS_CODE[ord(_c)] = 15
_CS_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate("0123"):
    _CS_CODE[ord(_c)] = _i
for _c in "4NnXx.":
    _CS_CODE[ord(_c)] = 15


def _pack_codes(codes: np.ndarray, n_words: int | None = None) -> np.ndarray:
    """4 bits per base, 8 per uint32, base i in bits 4*(i%8) of word i/8 (util.h:41-42)."""
    n = codes.size
    nw = (n + 7) // 8 if n_words is None else n_words
    buf = np.zeros(nw * 8, dtype=np.uint32)
    buf[:n] = codes
    buf = buf.reshape(nw, 8)
    sh = (4 * np.arange(8, dtype=np.uint32))[None, :]
    return np.bitwise_or.reduce(buf << sh, axis=1).astype(np.uint32)


def pack_bases(seq, n_words: int | None = None) -> np.ndarray:
    """ASCII letters -> packed 4-bit codes (fasta_sequence_to_bitfield, fasta.c:617-676)."""
    a = np.frombuffer(seq.encode() if isinstance(seq, str) else bytes(seq), dtype=np.uint8) \
        if not isinstance(seq, np.ndarray) else seq
    codes = _LS_CODE[a]
    if (codes == 255).any():
        raise ValueError("invalid character in letter-space sequence")
    return _pack_codes(codes, n_words)


def pack_colours(seq, n_words: int | None = None) -> np.ndarray:
    a = np.frombuffer(seq.encode() if isinstance(seq, str) else bytes(seq), dtype=np.uint8) \
        if not isinstance(seq, np.ndarray) else seq
    codes = _CS_CODE[a]
    if (codes == 255).any():
        raise ValueError("invalid character in colour-space sequence")
    return _pack_codes(codes, n_words)


@dataclass
class Scores:
    match: int
    mismatch: int
    a_gap_open: int
    a_gap_ext: int
    b_gap_open: int
    b_gap_ext: int
    crossover: int = 0


# gmapper-defaults.h:45-58
LS_DEFAULT_SCORES = Scores(10, -15, -33, -7, -33, -3, 0)
CS_DEFAULT_SCORES = Scores(10, -24, -33, -7, -33, -3, -20)


def score_alpha_beta(scores: Scores, colour_space: bool, pr_xover: float = 0.03):
    """score_alpha / score_beta exactly as gmapper.c:2559-2568 derives them (double, same libm)."""
    if colour_space:
        alpha = float(scores.crossover) / (math.log(pr_xover / 3) / math.log(2.0))
        pr_mismatch = 1.0 / (1.0 + 1.0 / 3.0 * math.pow(2.0, (float(scores.match) - float(scores.mismatch)) / alpha))
    else:
        pr_mismatch = .01
        alpha = (float(scores.match) - float(scores.mismatch)) / (
            math.log((1 - pr_mismatch) / (pr_mismatch / 3.0)) / math.log(2.0))
    beta = float(scores.match) - 2 * alpha - alpha * math.log(1 - pr_mismatch) / math.log(2.0)
    return alpha, beta


def crossover_scores_from_quals(quals, scores: Scores, qual_delta: int = 33, pr_xover: float = 0.03) -> np.ndarray:
    """read_entry::crossover_score from the quality strings of colour-space reads, exactly as gmapper.c:532-543
    (double arithmetic, same libm): (int)(alpha * log(pr_err(q) / 3) / log 2), clamped to [2 * crossover, -1]."""
    alpha, _ = score_alpha_beta(scores, True, pr_xover)
    width = max(len(q) for q in quals)
    out = np.zeros((len(quals), width), dtype=np.int32)
    for r, q in enumerate(quals):
        for j, ch in enumerate(q):
            qv = int(ch) - qual_delta
            pr = .99999999 if qv <= 0 else 1e-25 if qv >= 250 else math.pow(10.0, -float(qv) / 10.0)   # util.h:285-293
            v = int(alpha * math.log(pr / 3.0) / math.log(2.0))
            out[r, j] = -1 if v > -1 else max(v, 2 * scores.crossover)
    return out


def set_host_threads(n: int) -> None:
    """host threads of the stages the reference also runs on the CPU (read_pass2 ranking); 0 = OpenMP default"""
    from ._lib import lib
    lib().shrimp_gpu_set_host_threads(int(n))


def auto_list_cutoff(total_genome_len: int, max_seed_weight: int) -> int:
    """automatic index trimming, gmapper.c:2811-2837: max(1000, 100*L/4^W)"""
    c = ((100 * total_genome_len) // (4 ** max_seed_weight)) & 0xFFFFFFFF
    return c if c > 1000 else 1000


@dataclass
class MapParams:
    """Mapping options with the defaults of gmapper.h:50-141 / gmapper-defaults.h (unpaired)."""
    window_len: float = 140.0
    window_overlap: float = 90.0
    window_gen_threshold: float = 55.0
    sw_vect_threshold: float | None = None   # None: 47 % in colour space, = sw_full_threshold in letter space (gmapper.c:2464-2466)
    sw_full_threshold: float = 50.0
    match_mode: int = 2
    num_outputs: int = 10
    num_tmp_outputs: int = 30
    gapless: bool = False
    hash_filter_calls: bool = True
    use_regions: bool = True
    region_bits: int = 11
    region_overlap: int = 50
    Gflag: bool = True
    Tflag: bool = True
    strata: bool = False
    max_alignments: int = 0
    compute_mapping_qualities: bool = True
    list_cutoff: int = 0xFFFFFFFF

    def to_c(self, scores: Scores, colour_space: bool, crossover_scores: np.ndarray | None = None,
             read_quals: np.ndarray | None = None, qual_delta: int = 33) -> MapParamsC:
        alpha, beta = score_alpha_beta(scores, colour_space)
        vect = self.sw_vect_threshold
        if vect is None:
            vect = 47.0 if colour_space else self.sw_full_threshold
        return MapParamsC(self.window_len, self.window_overlap, self.window_gen_threshold, vect,
                          self.sw_full_threshold, alpha, beta, self.match_mode, self.num_outputs,
                          self.num_tmp_outputs, int(self.gapless), int(self.hash_filter_calls), int(self.use_regions),
                          self.region_bits, self.region_overlap, int(self.Gflag), int(self.Tflag), int(self.strata),
                          self.max_alignments, int(self.compute_mapping_qualities), self.list_cutoff & 0xFFFFFFFF,
                          crossover_scores.ctypes.data if crossover_scores is not None else None,
                          int(crossover_scores.shape[1]) if crossover_scores is not None else 0,
                          read_quals.ctypes.data if read_quals is not None else None,
                          int(read_quals.shape[1]) if read_quals is not None else 0, qual_delta, 0, 1, 0.0)


@dataclass
class MapResult:
    hits: np.ndarray            # structured array of shrimp_hit
    n_hits_per_read: np.ndarray
    edits: np.ndarray           # uint8 pool
    stage: np.ndarray | None
    stats: dict


PAIR_MODES = {"opp-in": 1, "opp-out": 2, "col-fw": 3, "col-bw": 4}


@dataclass
class PairResult:
    hits: np.ndarray             # shrimp_hit pool: members of the pairs first, then the unpaired hits in read order
    pairs: np.ndarray            # shrimp_pair records (hit_idx into hits)
    n_pairs_per_pair: np.ndarray
    n_unpaired_per_read: np.ndarray
    n_paired_hits: int           # hits[:n_paired_hits] belong to pairs
    edits: np.ndarray
    stats: dict


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class GpuContext:
    """One per process/GPU (the reference's per-thread *_setup state, gmapper.c:2907-2965)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        self._L = lib()
        check(self._L.shrimp_gpu_create(device, C.byref(self._h)), "shrimp_gpu_create")

    def close(self):
        if self._h:
            self._L.shrimp_gpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- set-up -------------------------------------------------------------------------------
    def sw_setup(self, dblen: int, qrlen: int, scores: Scores, use_colours: bool = False,
                 anchor_width: int = 8, indel_taboo_len: int = 0):
        """sw_vector_setup + sw_full_{ls,cs}_setup in one (sw-vector.c:388, sw-full-ls.c:573)."""
        mismatch = scores.mismatch
        p = SwParams(scores.match, mismatch, scores.a_gap_open, scores.a_gap_ext, scores.b_gap_open,
                     scores.b_gap_ext, scores.crossover, int(use_colours), anchor_width, indel_taboo_len,
                     qrlen, dblen)
        check(self._L.shrimp_gpu_sw_setup(self._h, C.byref(p)), "shrimp_gpu_sw_setup")

    # alias with the reference's name
    sw_vector_setup = sw_setup

    # ---- sw_vector ----------------------------------------------------------------------------
    def sw_vector(self, genome: np.ndarray, goff, glen, reads: np.ndarray, read_idx, rlen,
                  genome_ls: np.ndarray | None = None, initbp=None) -> np.ndarray:
        """Batched sw_vector(genome, goff, glen, read, rlen, genome_ls, initbp) (sw-vector.c:453)."""
        genome = np.ascontiguousarray(genome, dtype=np.uint32)
        reads = np.ascontiguousarray(reads, dtype=np.uint32)
        if reads.ndim != 2:
            raise ValueError("reads must be [n_reads, stride_words]")
        goff = np.ascontiguousarray(goff, dtype=np.uint32)
        glen = np.ascontiguousarray(glen, dtype=np.int32)
        read_idx = np.ascontiguousarray(read_idx, dtype=np.int32)
        rlen = np.ascontiguousarray(rlen, dtype=np.int32)
        n = goff.size
        if not (glen.size == read_idx.size == rlen.size == n):
            raise ValueError("task arrays must have equal length")
        if genome_ls is not None:
            genome_ls = np.ascontiguousarray(genome_ls, dtype=np.uint32)
            initbp = np.ascontiguousarray(initbp, dtype=np.int8)
        out = np.empty(n, dtype=np.int32)
        check(self._L.shrimp_gpu_sw_vector_batch(
            self._h, _ptr(genome), genome.size, _ptr(genome_ls), _ptr(reads), reads.shape[1], reads.shape[0],
            n, _ptr(goff), _ptr(glen), _ptr(read_idx), _ptr(rlen), _ptr(initbp), _ptr(out)),
            "shrimp_gpu_sw_vector_batch")
        return out

    # ---- sw_gapless ---------------------------------------------------------------------------
    def sw_gapless(self, genome: np.ndarray, goff, glen, reads: np.ndarray, read_idx, rlen, g_idx, r_idx,
                   genome_ls: np.ndarray | None = None, initbp=None) -> np.ndarray:
        """Batched sw_gapless(genome, glen, read, rlen, g_idx, r_idx, genome_ls, init_bp) (sw-gapless.c:57); the
        genome piece of task t starts at nibble goff[t] of `genome`."""
        genome = np.ascontiguousarray(genome, dtype=np.uint32)
        reads = np.ascontiguousarray(reads, dtype=np.uint32)
        arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (glen, read_idx, rlen, g_idx, r_idx)]
        goff = np.ascontiguousarray(goff, dtype=np.uint32)
        n = goff.size
        if genome_ls is not None:
            genome_ls = np.ascontiguousarray(genome_ls, dtype=np.uint32)
            initbp = np.ascontiguousarray(initbp, dtype=np.int8)
        out = np.empty(n, dtype=np.int32)
        check(self._L.shrimp_gpu_sw_gapless_batch(
            self._h, _ptr(genome), genome.size, _ptr(genome_ls), _ptr(reads), reads.shape[1], reads.shape[0], n,
            _ptr(goff), *[_ptr(a) for a in arrs], _ptr(initbp), _ptr(out)), "shrimp_gpu_sw_gapless_batch")
        return out

    # ---- sw_full_ls / sw_full_cs ----------------------------------------------------------------
    def sw_full(self, genome: np.ndarray, reads: np.ndarray, tasks: np.ndarray, local: bool = False):
        """Batched sw_full_ls (sw-full-ls.c:637) / sw_full_cs (sw-full-cs.c:1146, after a colour set-up).
        tasks: structured array of FullTaskC.  Returns (results structured array, edit-script pool)."""
        genome = np.ascontiguousarray(genome, dtype=np.uint32)
        reads = np.ascontiguousarray(reads, dtype=np.uint32)
        tasks = np.ascontiguousarray(tasks, dtype=FullTaskC)
        n = tasks.size
        res = np.zeros(max(1, n), dtype=FullResultC)
        cap = max(1024, int((tasks["glen"].astype(np.int64) + tasks["rlen"]).sum()))
        edits = np.zeros(cap, dtype=np.uint8)
        used = C.c_int64(0)
        check(self._L.shrimp_gpu_sw_full_batch(self._h, _ptr(genome), genome.size, _ptr(reads), reads.shape[1],
                                               reads.shape[0], n, _ptr(tasks), int(local), _ptr(res), _ptr(edits),
                                               cap, C.byref(used)), "shrimp_gpu_sw_full_batch")
        return res[:n], edits[: used.value]

    # ---- genome + index ---------------------------------------------------------------------
    def load_genome(self, contigs_packed: list, genome_len, colour_space: bool = False):
        """load_genome's arrays (genome.c:1092-1124): packed letter contigs -> HBM (+ rc, colour arrays)."""
        arrs = [np.ascontiguousarray(c, dtype=np.uint32) for c in contigs_packed]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        lens = np.ascontiguousarray(genome_len, dtype=np.uint32)
        check(self._L.shrimp_gpu_genome_load(self._h, len(arrs), C.cast(ptrs, C.c_void_p), _ptr(lens),
                                             int(colour_space)), "shrimp_gpu_genome_load")
        self.genome_len = lens.copy()
        self.total_len = int(lens.astype(np.int64).sum())
        self.colour_space = bool(colour_space)

    def share_genome_from(self, other: "GpuContext"):
        """Borrow the genome + projection resident in `other` (same GPU): one context per host thread, like
        gmapper's -N threads sharing the process-wide index."""
        check(self._L.shrimp_gpu_share_genome(self._h, other._h), "shrimp_gpu_share_genome")
        self._lender = other           # keep it alive
        for k in ("genome_len", "total_len", "colour_space", "seeds"):
            if hasattr(other, k):
                setattr(self, k, getattr(other, k))

    def genome_export(self, which: int) -> np.ndarray:
        out = np.zeros((self.total_len + 7) // 8, dtype=np.uint32)
        check(self._L.shrimp_gpu_genome_export(self._h, which, _ptr(out), out.size), "shrimp_gpu_genome_export")
        return out

    def build_index(self, seeds, hflag: bool = False):
        """projection loop of load_genome (genome.c:1138-1166) for seeds from shrimp_b200.seeds."""
        masks = np.array([s.mask for s in seeds], dtype=np.uint64)
        spans = np.array([s.span for s in seeds], dtype=np.int32)
        weights = np.array([s.weight for s in seeds], dtype=np.int32)
        check(self._L.shrimp_gpu_index_build(self._h, len(seeds), _ptr(masks), _ptr(spans), _ptr(weights),
                                             int(hflag)), "shrimp_gpu_index_build")
        self.seeds = list(seeds)

    def export_index(self, sn: int):
        """(genomemap_len[sn], concatenated genomemap[sn] lists) as in the -S files (genome.c:37-63)."""
        nb = C.c_uint32()
        tot = C.c_uint64()
        check(self._L.shrimp_gpu_index_nbuckets(self._h, sn, C.byref(nb), C.byref(tot)), "shrimp_gpu_index_nbuckets")
        lens = np.zeros(nb.value, dtype=np.uint32)
        pos = np.zeros(max(1, tot.value), dtype=np.uint32)
        check(self._L.shrimp_gpu_index_export(self._h, sn, _ptr(lens), _ptr(pos), C.byref(tot)),
              "shrimp_gpu_index_export")
        return lens, pos[: tot.value]

    def save_projection(self, prefix: str, contig_names):
        """save_genome_map (genome.c:185-272): <prefix>.genome + <prefix>.seed.N for `gmapper -L <prefix>`."""
        names = (C.c_char_p * len(contig_names))(*[n.encode() for n in contig_names])
        check(self._L.shrimp_gpu_projection_save(self._h, prefix.encode(), C.cast(names, C.c_void_p)),
              "shrimp_gpu_projection_save")

    def load_projection(self, prefix: str):
        """load_genome_map + load_genome_map_seed (genome.c:670-832, :69-182): the files of `gmapper -S` or of
        save_projection straight into HBM; returns the contig names stored in the file."""
        check(self._L.shrimp_gpu_projection_load(self._h, prefix.encode()), "shrimp_gpu_projection_load")
        n = int(self._L.shrimp_gpu_num_contigs(self._h))
        names = [self._L.shrimp_gpu_contig_name(self._h, c).decode() for c in range(n)]
        return names

    # ---- chunk mapping ------------------------------------------------------------------------
    def _buf(self, key, n, dtype, reuse):
        """output buffer; with reuse the same pages serve every call (results are views valid until the next call)"""
        if not reuse:
            return np.empty(n, dtype=dtype)
        cache = self.__dict__.setdefault("_out_bufs", {})
        b = cache.get(key)
        if b is None or b.size < n:
            b = cache[key] = np.empty(n, dtype=dtype)
        return b[:n]

    def map_reads(self, params: MapParams, scores: Scores, reads: np.ndarray, read_len, initbp=None,
                  want_stage: bool = False, stage_cap_per_read: int = 256, reuse_buffers: bool = False,
                  crossover_scores: np.ndarray | None = None, quals=None, qual_delta: int = 33) -> MapResult:
        """handle_read (mapping.c:1773) for a chunk: returns what read_output would receive.  crossover_scores
        [n, >= max read length] int32: read_entry::crossover_score of colour-space reads that came with qualities."""
        reads = np.ascontiguousarray(reads, dtype=np.uint32)
        read_len = np.ascontiguousarray(read_len, dtype=np.int32)
        n = reads.shape[0]
        if crossover_scores is not None:
            crossover_scores = np.ascontiguousarray(crossover_scores, dtype=np.int32)
        qbuf = None
        if quals is not None:   # re->qual of every read (gmapper -Q), for post_sw
            qw = max(len(q) for q in quals) + 1
            qbuf = np.zeros((len(quals), qw), dtype=np.uint8)
            for r, q in enumerate(quals):
                qbuf[r, :len(q)] = np.frombuffer(bytes(q), dtype=np.uint8)
        pc = params.to_c(scores, getattr(self, "colour_space", False), crossover_scores, qbuf, qual_delta)
        hits = self._buf("hits", max(1, n * params.num_outputs), HitC, reuse_buffers)   # untouched pages cost nothing
        n_per = self._buf("n_per", max(1, n), np.int32, reuse_buffers)
        max_rl = int(read_len.max()) if n else 0
        pool_cap = max(1024, n * 3 * max(1, max_rl))
        if initbp is not None:
            initbp = np.ascontiguousarray(initbp, dtype=np.int8)
        stage = np.empty(max(1, n * stage_cap_per_read), dtype=StageHitC) if want_stage else None
        while True:
            edits = self._buf("edits", pool_cap, np.uint8, reuse_buffers)
            n_hits, e_used, n_stage = C.c_int64(0), C.c_int64(0), C.c_int64(0)
            st = MapStatsC()
            rc = self._L.shrimp_gpu_map_reads(
                self._h, C.byref(pc), n, _ptr(reads), reads.shape[1] if n else 1, _ptr(read_len), _ptr(initbp),
                _ptr(hits), hits.size, _ptr(n_per), _ptr(edits), pool_cap, C.byref(n_hits), C.byref(e_used),
                _ptr(stage), stage.size if want_stage else 0, C.byref(n_stage), C.byref(st))
            if rc == -5 and e_used.value > pool_cap:   # SHRIMP_E_NOMEM: edit pool too small, size is returned
                pool_cap = int(e_used.value) + 1024
                continue
            check(rc, "shrimp_gpu_map_reads")
            break
        stats = {k: int(getattr(st, k)) for k, _ in MapStatsC._fields_}
        return MapResult(hits[: n_hits.value], n_per[:n], edits[: e_used.value],
                         stage[: n_stage.value] if want_stage else None, stats)

    def map_pairs(self, params: MapParams, scores: Scores, reads: np.ndarray, read_len, pair_mode: str = "opp-in",
                  min_insert: int = 0, max_insert: int = 1000, half_paired: bool = True, initbp=None,
                  reuse_buffers: bool = False) -> "PairResult":
        """handle_readpair (mapping.c:2504) for a chunk of pairs: rows 2k and 2k+1 of `reads` are mates."""
        reads = np.ascontiguousarray(reads, dtype=np.uint32)
        read_len = np.ascontiguousarray(read_len, dtype=np.int32)
        n = reads.shape[0]
        if n % 2:
            raise ValueError("map_pairs needs an even number of reads (mates interleaved)")
        npairs = n // 2
        pc = params.to_c(scores, getattr(self, "colour_space", False))
        pp = PairParamsC(PAIR_MODES[pair_mode], min_insert, max_insert, int(half_paired))
        hits = self._buf("p_hits", max(1, npairs * params.num_outputs * 4), HitC, reuse_buffers)
        pairs = self._buf("p_pairs", max(1, npairs * params.num_outputs), PairC, reuse_buffers)
        n_per_pair = self._buf("p_npp", max(1, npairs), np.int32, reuse_buffers)
        n_unp = self._buf("p_nunp", max(1, n), np.int32, reuse_buffers)
        max_rl = int(read_len.max()) if n else 0
        pool_cap = max(1024, n * 4 * max(1, max_rl))
        if initbp is not None:
            initbp = np.ascontiguousarray(initbp, dtype=np.int8)
        while True:
            edits = self._buf("p_edits", pool_cap, np.uint8, reuse_buffers)
            n_hits, n_pairs_out, e_used = C.c_int64(0), C.c_int64(0), C.c_int64(0)
            st = MapStatsC()
            rc = self._L.shrimp_gpu_map_pairs(
                self._h, C.byref(pc), C.byref(pp), npairs, _ptr(reads), reads.shape[1] if n else 1, _ptr(read_len),
                _ptr(initbp), _ptr(hits), hits.size, C.byref(n_hits), _ptr(pairs), pairs.size, C.byref(n_pairs_out),
                _ptr(n_per_pair), _ptr(n_unp), _ptr(edits), pool_cap, C.byref(e_used), C.byref(st))
            if rc == -5 and e_used.value > pool_cap:
                pool_cap = int(e_used.value) + 1024
                continue
            check(rc, "shrimp_gpu_map_pairs")
            break
        stats = {k: int(getattr(st, k)) for k, _ in MapStatsC._fields_}
        n_paired_hits = 2 * n_pairs_out.value
        return PairResult(hits[: n_hits.value], pairs[: n_pairs_out.value], n_per_pair[:npairs], n_unp[:n],
                          n_paired_hits, edits[: e_used.value], stats)

    def map_resident(self, params: MapParams, scores: Scores) -> dict:
        """Device stages only, on the reads the last map_reads call left in HBM (bench.py `value`)."""
        pc = params.to_c(scores, getattr(self, "colour_space", False))
        st = MapStatsC()
        check(self._L.shrimp_gpu_map_resident(self._h, C.byref(pc), C.byref(st)), "shrimp_gpu_map_resident")
        return {k: int(getattr(st, k)) for k, _ in MapStatsC._fields_}

    def map_pairs_resident(self, params: MapParams, scores: Scores, pair_mode: str = "opp-in", min_insert: int = 0,
                           max_insert: int = 1000) -> dict:
        """All stages of map_pairs on the pairs the last map_pairs call left in HBM (bench.py `value`)."""
        pc = params.to_c(scores, getattr(self, "colour_space", False))
        pp = PairParamsC(PAIR_MODES[pair_mode], min_insert, max_insert, 1)
        st = MapStatsC()
        check(self._L.shrimp_gpu_map_pairs_resident(self._h, C.byref(pc), C.byref(pp), C.byref(st)),
              "shrimp_gpu_map_pairs_resident")
        return {k: int(getattr(st, k)) for k, _ in MapStatsC._fields_}

    def last_transfer_bytes(self):
        a, b = C.c_uint64(), C.c_uint64()
        check(self._L.shrimp_gpu_last_transfer_bytes(self._h, C.byref(a), C.byref(b)), "shrimp_gpu_last_transfer_bytes")
        return int(a.value), int(b.value)

    def event_record(self, which: int):
        check(self._L.shrimp_gpu_event_record(self._h, which), "shrimp_gpu_event_record")

    def event_elapsed_ms(self) -> float:
        v = C.c_float()
        check(self._L.shrimp_gpu_event_elapsed_ms(self._h, C.byref(v)), "shrimp_gpu_event_elapsed_ms")
        return float(v.value)

    def flush_l2(self):
        check(self._L.shrimp_gpu_flush_l2(self._h), "shrimp_gpu_flush_l2")

    def dpx_peak(self) -> float:
        """Measured integer-pipe peak in G thread-instructions/s (VIADDMNMX.S16x2)."""
        v = C.c_double()
        check(self._L.shrimp_gpu_dpx_peak(self._h, C.byref(v)), "shrimp_gpu_dpx_peak")
        return float(v.value)

    def fp64_peak(self) -> float:
        """Measured FP64 issue peak in G thread-instructions/s (DFMA)."""
        v = C.c_double()
        check(self._L.shrimp_gpu_fp64_peak(self._h, C.byref(v)), "shrimp_gpu_fp64_peak")
        return float(v.value)

    # ---- accounting ---------------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self._L.shrimp_gpu_launch_count(self._h))

    def stage_times(self) -> dict:
        names = (C.c_char_p * 16)()
        ms = (C.c_float * 16)()
        ln = (C.c_uint64 * 16)()
        n = self._L.shrimp_gpu_stage_times(self._h, names, ms, ln, 16)
        return {names[i].decode(): (float(ms[i]), int(ln[i])) for i in range(n)}

    def stage_times_reset(self):
        self._L.shrimp_gpu_stage_times_reset(self._h)
