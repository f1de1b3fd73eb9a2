"""TEST INFRASTRUCTURE ONLY: Python face of the oracle's per-read pipeline (oracle_pipeline.inc)
plus the SAM-field derivation of gmapper/output.c used to compare against reference SAM files."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import OrcScores, _p, oracle_lib

ALN_CAP = 640


class OrcSfr(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("read_start", "rmapped", "genome_start", "gmapped", "matches", "mismatches",
                                       "insertions", "deletions", "score", "crossovers")] + \
               [("dbalign", C.c_char * ALN_CAP), ("qralign", C.c_char * ALN_CAP), ("qual", C.c_char * ALN_CAP)]


class OrcParams(C.Structure):
    _fields_ = [("sc", OrcScores), ("colour_space", C.c_int),
                ("window_len", C.c_double), ("window_overlap", C.c_double), ("window_gen_threshold", C.c_double),
                ("sw_vect_threshold", C.c_double), ("sw_full_threshold", C.c_double),
                ("match_mode", C.c_int), ("num_outputs", C.c_int), ("num_tmp_outputs", C.c_int),
                ("anchor_width", C.c_int), ("indel_taboo_len", C.c_int), ("gapless", C.c_int),
                ("hash_filter_calls", C.c_int), ("use_regions", C.c_int), ("region_bits", C.c_int),
                ("region_overlap", C.c_int), ("Gflag", C.c_int), ("Tflag", C.c_int), ("strata", C.c_int),
                ("max_alignments", C.c_int), ("compute_mapping_qualities", C.c_int), ("list_cutoff", C.c_uint32),
                ("score_alpha", C.c_double), ("score_beta", C.c_double)]


class OrcStageHit(C.Structure):
    _fields_ = [("read_idx", C.c_int), ("st", C.c_int), ("cn", C.c_int), ("w_len", C.c_int), ("g_off", C.c_longlong),
                ("score_window_gen", C.c_int), ("matches", C.c_int), ("score_max", C.c_int),
                ("score_vector", C.c_int), ("pct_score_vector", C.c_int),
                ("ax", C.c_int), ("ay", C.c_int), ("alen", C.c_int), ("awidth", C.c_int)]


class OrcHitOut(C.Structure):
    _fields_ = [("read_idx", C.c_int), ("cn", C.c_int), ("gen_st", C.c_int), ("st", C.c_int), ("g_off", C.c_longlong),
                ("w_len", C.c_int), ("score_vector", C.c_int), ("score_full", C.c_int), ("pass2_key", C.c_int),
                ("score_max", C.c_int), ("matches", C.c_int), ("sw_score", C.c_int), ("posterior", C.c_double),
                ("sfr", OrcSfr)]


class OrcPairParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("pair_mode", "min_insert_size", "max_insert_size", "half_paired")]


PAIR_MODES = {"opp-in": 1, "opp-out": 2, "col-fw": 3, "col-bw": 4}


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("vector_calls", "vector_cells", "vector_bypassed", "full_calls",
                                          "full_cells", "n_anchors", "n_hits", "eq_x_ties")]


def score_alpha_beta(scores, colour_space: bool, pr_xover: float = 0.03):
    """gmapper.c:2559-2568 (double arithmetic, same libm)."""
    if colour_space:
        alpha = float(scores.crossover) / (math.log(pr_xover / 3) / math.log(2.0))
        pr_mismatch = 1.0 / (1.0 + 1.0 / 3.0 * math.pow(2.0, (float(scores.match) - float(scores.mismatch)) / alpha))
    else:
        pr_mismatch = .01
        alpha = (float(scores.match) - float(scores.mismatch)) / (math.log((1 - pr_mismatch) / (pr_mismatch / 3.0)) / math.log(2.0))
    beta = float(scores.match) - 2 * alpha - alpha * math.log(1 - pr_mismatch) / math.log(2.0)
    return alpha, beta


def auto_list_cutoff(total_genome_len: int, max_seed_weight: int) -> int:
    """gmapper.c:2811-2837: max(1000, 100*L/4^W)."""
    c = (100 * total_genome_len) // (4 ** max_seed_weight)
    c &= 0xFFFFFFFF
    return c if c > 1000 else 1000


@dataclass
class MapOptions:
    """Defaults of gmapper.h:50-141 / gmapper-defaults.h for unpaired mapping."""
    colour_space: bool = False
    window_len: float = 140.0
    window_overlap: float = 90.0
    window_gen_threshold: float = 55.0
    sw_vect_threshold: float | None = None   # None: 47 % in colour space, = sw_full_threshold in letter space (gmapper.c:2464-2466)
    sw_full_threshold: float = 50.0
    match_mode: int = 2
    num_outputs: int = 10
    num_tmp_outputs: int = 30
    anchor_width: int = 8
    indel_taboo_len: int = 0
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
    scores: object = None
    extra: dict = field(default_factory=dict)

    def to_struct(self) -> OrcParams:
        s = self.scores
        alpha, beta = score_alpha_beta(s, self.colour_space)
        vect = self.sw_vect_threshold
        if vect is None:
            vect = 47.0 if self.colour_space else self.sw_full_threshold
        return OrcParams(OrcScores(s.match, s.mismatch, s.a_gap_open, s.a_gap_ext, s.b_gap_open, s.b_gap_ext,
                                   s.crossover),
                         int(self.colour_space), self.window_len, self.window_overlap, self.window_gen_threshold,
                         vect, self.sw_full_threshold, self.match_mode, self.num_outputs,
                         self.num_tmp_outputs, self.anchor_width, self.indel_taboo_len, int(self.gapless),
                         int(self.hash_filter_calls), int(self.use_regions), self.region_bits, self.region_overlap,
                         int(self.Gflag), int(self.Tflag), int(self.strata), self.max_alignments,
                         int(self.compute_mapping_qualities), self.list_cutoff & 0xFFFFFFFF, alpha, beta)


class Genome:
    def __init__(self, contig_codes: list[np.ndarray], colour_space: bool):
        L = oracle_lib()
        L.orc_genome_create.restype = C.c_void_p
        L.orc_genome_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        self.lens = np.array([c.size for c in contig_codes], dtype=np.uint32)
        codes = np.ascontiguousarray(np.concatenate(contig_codes).astype(np.uint8))
        self.h = L.orc_genome_create(len(contig_codes), _p(codes), _p(self.lens), int(colour_space))
        self.colour_space = colour_space
        self.total_len = int(self.lens.sum())

    def __del__(self):
        try:
            L = oracle_lib()
            L.orc_genome_destroy.argtypes = [C.c_void_p]
            L.orc_genome_destroy(self.h)
        except Exception:
            pass


class OrcIndexStruct(C.Structure):
    _fields_ = [("n_seeds", C.c_int), ("mask", C.c_uint64 * 16), ("span", C.c_int * 16), ("weight", C.c_int * 16),
                ("max_span", C.c_int), ("min_span", C.c_int), ("hflag", C.c_int), ("nbuckets", C.c_uint32 * 16),
                ("len", C.POINTER(C.c_uint32) * 16), ("start", C.POINTER(C.c_uint32) * 16),
                ("pos", C.POINTER(C.c_uint32) * 16), ("total", C.c_uint64 * 16)]


class Index:
    def __init__(self, genome: Genome, seeds, hflag: bool = False):
        L = oracle_lib()
        L.orc_index_build.restype = C.POINTER(OrcIndexStruct)
        L.orc_index_build.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        masks = np.array([s.mask for s in seeds], dtype=np.uint64)
        spans = np.array([s.span for s in seeds], dtype=np.int32)
        weights = np.array([s.weight for s in seeds], dtype=np.int32)
        self.seeds = list(seeds)
        self.genome = genome
        self.h = L.orc_index_build(genome.h, len(seeds), _p(masks), _p(spans), _p(weights), int(hflag))

    def bucket_lens(self, sn: int) -> np.ndarray:
        s = self.h.contents
        return np.ctypeslib.as_array(s.len[sn], shape=(int(s.nbuckets[sn]),)).copy()

    def positions(self, sn: int) -> np.ndarray:
        s = self.h.contents
        return np.ctypeslib.as_array(s.pos[sn], shape=(int(s.total[sn]),)).copy()

    def __del__(self):
        try:
            L = oracle_lib()
            L.orc_index_destroy.argtypes = [C.c_void_p]
            L.orc_index_destroy(self.h)
        except Exception:
            pass


def map_reads(genome: Genome, index: Index, opts: MapOptions, reads: np.ndarray, read_len: np.ndarray,
              initbp: np.ndarray | None = None, want_stage: bool = False, stage_cap_per_read: int = 256,
              crossover_scores: np.ndarray | None = None, quals=None, qual_delta: int = 33):
    """Returns (hits structured array, n_out_per_read, stage array or None, stats dict).  crossover_scores [n, w]
    int32: read_entry::crossover_score of colour-space reads with qualities (gmapper.c:532-543)."""
    L = oracle_lib()
    L.orc_map_reads.restype = C.c_longlong
    L.orc_set_crossover_scores.argtypes = [C.c_void_p, C.c_int]
    L.orc_set_crossover_scores.restype = None
    if crossover_scores is not None:
        crossover_scores = np.ascontiguousarray(crossover_scores, dtype=np.int32)
        L.orc_set_crossover_scores(_p(crossover_scores), int(crossover_scores.shape[1]))
    L.orc_set_read_quals.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_set_read_quals.restype = None
    qbuf = None
    if quals is not None:   # post_sw reads re->qual (gmapper -Q)
        w = max(len(q) for q in quals) + 1
        qbuf = np.zeros((len(quals), w), dtype=np.uint8)
        for r, q in enumerate(quals):
            qbuf[r, :len(q)] = np.frombuffer(bytes(q), dtype=np.uint8)
        L.orc_set_read_quals(_p(qbuf), w, qual_delta, 0, 1)
    reads = np.ascontiguousarray(reads, dtype=np.uint32)
    read_len = np.ascontiguousarray(read_len, dtype=np.int32)
    n = reads.shape[0]
    p = opts.to_struct()
    out = (OrcHitOut * max(1, n * opts.num_outputs))()
    n_per = np.zeros(n, dtype=np.int32)
    stats = OrcStats()
    stage = None
    n_stage = C.c_longlong(0)
    cap = n * stage_cap_per_read
    if want_stage:
        stage = (OrcStageHit * max(1, cap))()
    if initbp is not None:
        initbp = np.ascontiguousarray(initbp, dtype=np.int8)
    L.orc_map_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
    rc = L.orc_map_reads(genome.h, index.h, C.byref(p), n, _p(reads), reads.shape[1], _p(read_len), _p(initbp),
                         C.cast(out, C.c_void_p), len(out), _p(n_per),
                         C.cast(stage, C.c_void_p) if want_stage else None, cap, C.byref(n_stage), C.byref(stats))
    L.orc_set_crossover_scores(None, 0)
    L.orc_set_read_quals(None, 0, 33, 0, 1)
    if rc < 0:
        raise RuntimeError("oracle capacity too small")
    hits = np.ctypeslib.as_array(out)[:rc] if rc > 0 else np.ctypeslib.as_array(out)[:0]
    st = None
    if want_stage:
        st = np.ctypeslib.as_array(stage)[: n_stage.value]
    sd = {k: int(getattr(stats, k)) for k, _ in OrcStats._fields_}
    return hits, n_per, st, sd


def map_pairs(genome: Genome, index: Index, opts: MapOptions, reads: np.ndarray, read_len: np.ndarray,
              pair_mode: str = "opp-in", min_insert: int = 0, max_insert: int = 1000, half_paired: bool = True,
              initbp: np.ndarray | None = None):
    """handle_readpair for pairs (reads 2k, 2k+1).  Returns (pair_hits [n,2] structured, pair_info [n,5]
    (pair, score, score_max, key, insert_size), n_pairs_per_pair, unpaired hits, n_unp_per_read, stats)."""
    L = oracle_lib()
    L.orc_map_pairs.restype = C.c_longlong
    reads = np.ascontiguousarray(reads, dtype=np.uint32)
    read_len = np.ascontiguousarray(read_len, dtype=np.int32)
    n = reads.shape[0]
    assert n % 2 == 0
    npairs = n // 2
    p = opts.to_struct()
    pp = OrcPairParams(PAIR_MODES[pair_mode], min_insert, max_insert, int(half_paired))
    pcap = max(1, npairs * opts.num_outputs)
    pout = (OrcHitOut * (2 * pcap))()
    pinfo = np.zeros((pcap, 5), dtype=np.int32)
    nper = np.zeros(max(1, npairs), dtype=np.int32)
    ucap = max(1, n * opts.num_outputs)
    uout = (OrcHitOut * ucap)()
    nunp_per = np.zeros(max(1, n), dtype=np.int32)
    nunp = C.c_longlong(0)
    stats = OrcStats()
    if initbp is not None:
        initbp = np.ascontiguousarray(initbp, dtype=np.int8)
    L.orc_map_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong,
                                C.c_void_p, C.c_void_p, C.c_void_p]
    rc = L.orc_map_pairs(genome.h, index.h, C.byref(p), C.byref(pp), npairs, _p(reads), reads.shape[1], _p(read_len),
                         _p(initbp), C.cast(pout, C.c_void_p), _p(pinfo), pcap, _p(nper), C.cast(uout, C.c_void_p),
                         ucap, _p(nunp_per), C.byref(nunp), C.byref(stats))
    if rc < 0:
        raise RuntimeError("oracle capacity too small")
    ph = np.ctypeslib.as_array(pout)[: 2 * rc].reshape(rc, 2) if rc > 0 else np.ctypeslib.as_array(pout)[:0].reshape(0, 2)
    uh = np.ctypeslib.as_array(uout)[: nunp.value]
    sd = {k: int(getattr(stats, k)) for k, _ in OrcStats._fields_}
    return ph, pinfo[:rc], nper[:npairs], uh, nunp_per[:n], sd


def pair_sam_records(pair_hits, pair_of, unp_hits, genome_lens, read_len, fields_of, n_pairs):
    """The mapped SAM records readpair_output / readpair_output_no_mqv + read_output print for each pair, in print
    order, as tuples (pair, mate, flag, cn, pos, cigar, mate_cn, mpos, isize, AS, NM) -- gmapper/output.c
    hit_output :228-700.  fields_of(hit, read_len, contig_len) -> (flag16, cn, pos, cigar, AS, NM, gmapped)."""
    out = []
    per_pair = [[] for _ in range(n_pairs)]
    per_read = [[] for _ in range(2 * n_pairs)]
    for k, hh in zip(pair_of, pair_hits):
        per_pair[int(k)].append(hh)
    for h in unp_hits:
        per_read[int(h["read_idx"])].append(h)

    def rec(k, mate, h, mp, paired_alignment):
        f = fields_of(h, int(read_len[int(h["read_idx"])]), int(genome_lens[int(h["cn"])]))
        rev = f[0] == 16
        flag = 1 | (2 if paired_alignment else 0) | (16 if rev else 0) | (0x40 if mate == 0 else 0x80)
        mcn, mpos, isize = -1, 0, 0
        if mp is None:
            flag |= 8
        else:
            fm = fields_of(mp, int(read_len[int(mp["read_idx"])]), int(genome_lens[int(mp["cn"])]))
            rev_mp = fm[0] == 16
            if rev_mp:
                flag |= 0x20
            mcn, mpos = fm[1], fm[2]
            if fm[1] == f[1]:
                fivep = f[2] + f[6] - 1 if rev else f[2] - 1
                fivep_mp = fm[2] + fm[6] - 1 if rev_mp else fm[2] - 1
                isize = fivep_mp - fivep
        return (k, mate, flag, f[1], f[2], f[3], mcn, mpos, isize, f[4], f[5])

    for k in range(n_pairs):
        for hh in per_pair[k]:
            out.append(rec(k, 0, hh[0], hh[1], True))
            out.append(rec(k, 1, hh[1], hh[0], True))
        for mate in range(2):
            for h in per_read[2 * k + mate]:
                out.append(rec(k, mate, h, None, False))
    return out


def parse_pair_sam(path: str, contig_idx: dict):
    """mapped records of a paired SAM as the tuples of pair_sam_records (qnames p<k>)."""
    out = []
    with open(path) as f:
        for line in f:
            if line.startswith("@"):
                continue
            t = line.rstrip("\n").split("\t")
            flag = int(t[1])
            if flag & 4:
                continue
            tags = {x[:2]: x[5:] for x in t[11:]}
            mcn = -1 if t[6] == "*" else (contig_idx[t[2]] if t[6] == "=" else contig_idx[t[6]])
            out.append((int(t[0][1:]), 0 if flag & 0x40 else 1, flag, contig_idx[t[2]], int(t[3]), t[5], mcn, int(t[7]),
                        int(t[8]), int(tags["AS"]), int(tags["NM"])))
    return out


# ------------------------------------------------------------------------------------------------
# SAM fields that the hot path determines (gmapper/output.c: make_cigar :15-65, hit_output :470-700)
# ------------------------------------------------------------------------------------------------
def cigar_from_alignment(read_start0: int, rmapped: int, read_len: int, qralign: bytes, dbalign: bytes,
                         reverse: bool, clip: str = "S") -> str:
    ops = []
    read_start = read_start0 + 1
    read_end = read_start + rmapped - 1
    if read_start > 1:
        ops.append((read_start - 1, clip))
    i, n = 0, len(qralign)
    while i < n:
        if qralign[i:i + 1] == b"-":
            op, test = "D", (lambda k: qralign[k:k + 1] == b"-")
        elif dbalign[i:i + 1] == b"-":
            op, test = "I", (lambda k: dbalign[k:k + 1] == b"-")
        else:
            op, test = "M", (lambda k: dbalign[k:k + 1] != b"-" and qralign[k:k + 1] != b"-")
        ln = 0
        while i + ln < n and test(i + ln):
            ln += 1
        ops.append((ln, op))
        i += ln
    if read_end != read_len:
        ops.append((read_len - read_end, clip))
    if reverse:
        ops = ops[::-1]
    return "".join(f"{l}{o}" for l, o in ops)


def sam_fields(hit, read_len: int, genome_len_cn: int, colour_space: bool = False):
    """(flag, cn, pos, cigar, AS, NM) as hit_output prints them for an unpaired read."""
    sfr = hit["sfr"]
    reverse = int(hit["gen_st"]) == 1
    read_start = int(sfr["read_start"]) + 1
    read_end = read_start + int(sfr["rmapped"]) - 1
    if not reverse:
        pos = int(sfr["genome_start"]) + 1
    else:
        right = genome_len_cn - int(sfr["genome_start"])
        pos = right - (read_end - read_start - int(sfr["deletions"]) + int(sfr["insertions"]))
    qr = bytes(sfr["qralign"]).split(b"\0")[0]
    db = bytes(sfr["dbalign"]).split(b"\0")[0]
    cigar = cigar_from_alignment(int(sfr["read_start"]), int(sfr["rmapped"]), read_len, qr, db, reverse,
                                 "H" if colour_space else "S")
    nm = int(sfr["mismatches"]) + int(sfr["deletions"]) + int(sfr["insertions"])
    return (16 if reverse else 0, int(hit["cn"]), pos, cigar, int(hit["score_full"]), nm, int(sfr["gmapped"]))


def parse_sam_seq_qual(path: str):
    """[(SEQ, QUAL)] of the mapped records, in file order (colour space with mapping qualities: the corrected base
    calls and base qualities of post_sw, gmapper/output.c:483-640)."""
    out = []
    with open(path) as f:
        for line in f:
            if line.startswith("@"):
                continue
            t = line.rstrip("\n").split("\t")
            if int(t[1]) & 4:
                continue
            out.append((t[9], t[10]))
    return out


_RC = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def seq_qual_of_alignment(qralign: bytes, qual: bytes, reverse: bool, have_quals: bool):
    """SEQ / QUAL columns hit_output prints for a colour-space alignment with mapping qualities (output.c:486-546,
    :607-616, :626-628): the letters of qralign upper-cased (reverse-complemented on the reverse strand), and
    sfrp->qual (reversed there) when the reads came with qualities, else '*'."""
    seq = bytes(c for c in qralign if c != ord("-")).upper()
    q = bytes(qual) if have_quals else b"*"
    if reverse:
        seq = seq.translate(_RC)[::-1]
        if have_quals:
            q = q[::-1]
    return seq.decode(), q.decode()


def parse_sam(path: str):
    """[(qname, flag, rname, pos, cigar, AS, NM)] for mapped records, in file order."""
    out = []
    with open(path) as f:
        for line in f:
            if line.startswith("@"):
                continue
            t = line.rstrip("\n").split("\t")
            if int(t[1]) & 4:
                continue
            tags = {x[:2]: x[5:] for x in t[11:]}
            out.append((t[0], int(t[1]), t[2], int(t[3]), t[5], int(tags["AS"]), int(tags["NM"])))
    return out
