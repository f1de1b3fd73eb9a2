"""Synthetic mapping configs shared by the oracle and GPU pipeline tests (seeded, see tools/gen_synth.py)."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_synth  # noqa: E402

from shrimp_b200 import seeds as S  # noqa: E402
from shrimp_b200.api import CS_DEFAULT_SCORES, LS_DEFAULT_SCORES, _CS_CODE, _LS_CODE, _pack_codes  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name -> (generator config, gmapper command-line options, option overrides for MapOptions)
MAP_CASES = {
    "c1_small": dict(gen="c1_small", args=[], opts={}),
    "c1_repeat": dict(gen="c1_repeat", args=[], opts={}),
    # colour space: post_sw is not on the path yet, so parity runs with --no-mapping-qualities
    "c2_small": dict(gen="c2_small", args=["--no-mapping-qualities"], opts={"compute_mapping_qualities": False}),
    # BASELINE.json configs[3]: 22 bp reads vs a miRNA-like database, default options and -M mirna (gmapper.c:
    # 1498-1515: hashed 5-seed set, gapless pass 1, gap opens -255, no window cache, one seed match, window 100 %,
    # local full SW, no mapping qualities)
    # -U (ungapped) in colour space: sw_gapless with the first colour forced through the letter genome
    # (sw-gapless.c:83-93), gap opens -255, anchor_width 0, no window cache; -U needs local mode (gmapper.c:2330)
    "c2_small_ungapped": dict(gen="c2_small", args=["--no-mapping-qualities", "-U", "--local"],
                              opts=dict(gapless=True, hash_filter_calls=False, Gflag=False,
                                        compute_mapping_qualities=False),
                              gap_open=-255, anchor_width=0),
    # colour-space reads WITH qualities (-Q): per-position crossover scores in sw_full_cs (gmapper.c:532-543,
    # sw-full-cs.c:312); qualities are seeded, 2..40, PHRED+33
    "c2_small_fastq": dict(gen="c2_small", args=["-Q", "--no-mapping-qualities"],
                           opts={"compute_mapping_qualities": False}, fastq=True),
    "c2_small_fastq_local": dict(gen="c2_small", args=["-Q", "--no-mapping-qualities", "--local"],
                                 opts={"compute_mapping_qualities": False, "Gflag": False}, fastq=True),
    # colour space with mapping qualities (the reference's default): post_sw (sw-post.c) rescoring every alignment
    "c2_small_mq": dict(gen="c2_small", args=[], opts={}),
    "c2_small_fastq_mq": dict(gen="c2_small", args=["-Q"], opts={}, fastq=True),
    "c4_small": dict(gen="c4_small", args=[], opts={}),
    "c4_small_mirna": dict(gen="c4_small", args=["-M", "mirna"],
                           opts=dict(match_mode=1, window_len=100.0, gapless=True, hash_filter_calls=False,
                                     Gflag=False, compute_mapping_qualities=False),
                           mirna=True, anchor_width=0),
    # BASELINE.json configs[4], "overly sensitive" mode (README:481-534, letter-space subset): four weight-11 seeds,
    # one seed match is enough (no region filter, every index position is an anchor), wide windows, threshold band
    # of the full SW (-a -1), no window cache (-Z), no index trimming (-V)
    "c5_small": dict(gen="c5_small",
                     args=["-s", "w11", "-n", "1", "-w", "150%", "-r", "50%", "-l", "40%", "-Z", "-h", "60%", "-a", "-1",
                           "-V"],
                     opts=dict(match_mode=1, window_len=150.0, window_gen_threshold=50.0, window_overlap=40.0,
                               hash_filter_calls=False, sw_full_threshold=60.0),
                     seeds_weight=11, anchor_width=-1, list_cutoff=0xFFFFFFFF),
}


class LsCase:
    def __init__(self, name: str):
        cfg = gen_synth.CONFIGS[MAP_CASES[name]["gen"]]
        self.name = name
        self.colour = bool(cfg.get("colour"))
        self.binary = "gmapper-cs" if self.colour else "gmapper-ls"
        self.contigs = gen_synth.make_genome(**cfg["genome"])
        self.contig_codes = [_LS_CODE[s] for _, s in self.contigs]
        self.contig_names = [n for n, _ in self.contigs]
        if self.colour:
            self.reads = gen_synth.simulate_cs_reads(self.contigs, **cfg["reads"])
            seqs = [np.frombuffer(r[1], dtype=np.uint8) for r in self.reads]
            self.read_len = np.array([s.size - 1 for s in seqs], dtype=np.int32)   # colours (gmapper.c:476)
            self.initbp = np.array([_LS_CODE[s[0]] for s in seqs], dtype=np.int8)
            self.stride = int((self.read_len.max() + 7) // 8)
            self.packed = np.stack([_pack_codes(_CS_CODE[s[1:]].astype(np.uint32), self.stride) for s in seqs])
            self.scores = CS_DEFAULT_SCORES
        else:
            self.reads = gen_synth.simulate_reads(self.contigs, **cfg["reads"])
            self.read_len = np.array([r[1].size for r in self.reads], dtype=np.int32)
            self.initbp = None
            self.stride = int((self.read_len.max() + 7) // 8)
            self.packed = np.stack([_pack_codes(_LS_CODE[r[1]].astype(np.uint32), self.stride) for r in self.reads])
            self.scores = LS_DEFAULT_SCORES
        self.read_names = [r[0] for r in self.reads]
        spec = MAP_CASES[name]
        self.hflag = bool(spec.get("mirna"))
        if self.hflag:
            self.seeds = S.load_default_mirna_seeds()
            from shrimp_b200.api import Scores
            self.scores = Scores(self.scores.match, self.scores.mismatch, -255, self.scores.a_gap_ext, -255,
                                 self.scores.b_gap_ext, self.scores.crossover)
        else:
            self.seeds = S.load_default_seeds(spec.get("seeds_weight", 0))
            if "gap_open" in spec:
                from shrimp_b200.api import Scores
                self.scores = Scores(self.scores.match, self.scores.mismatch, spec["gap_open"], self.scores.a_gap_ext,
                                     spec["gap_open"], self.scores.b_gap_ext, self.scores.crossover)
        self.anchor_width = spec.get("anchor_width", 8)
        self.quals = None
        self.crossover_scores = None
        if spec.get("fastq"):
            from shrimp_b200.api import crossover_scores_from_quals
            rq = np.random.default_rng(4242)
            self.quals = [(33 + rq.integers(2, 41, size=int(n))).astype(np.uint8) for n in self.read_len]
            self.crossover_scores = crossover_scores_from_quals(self.quals, self.scores, qual_delta=33)
        self.total_len = int(sum(c.size for c in self.contig_codes))
        from shrimp_b200.api import auto_list_cutoff
        self.list_cutoff = spec.get("list_cutoff", auto_list_cutoff(
            self.total_len, 12 if self.hflag else max(s.weight for s in self.seeds)))

    def write_fasta(self, d: str):
        os.makedirs(d, exist_ok=True)
        gen_synth.write_fasta(os.path.join(d, "genome.fa"), self.contigs, width=80)
        if self.quals is None:
            gen_synth.write_fasta(os.path.join(d, "reads.fa"), self.reads)
        else:   # FASTQ (gmapper -Q); colour-space reads carry one quality per colour, none for the primer base
            with open(os.path.join(d, "reads.fa"), "wb") as f:
                for (name, seq), q in zip(self.reads, self.quals):
                    f.write(b"@" + name.encode() + b"\n" + bytes(seq) + b"\n+\n" + bytes(q) + b"\n")


def stage_tuple_array(stage) -> np.ndarray:
    """[n, 12] int64: read, st, cn, g_off, w_len, wg, vc, matches, ax, ay, alen, awidth"""
    cols = ["read_idx", "st", "cn", "g_off", "w_len", "score_window_gen", "score_vector", "matches", "ax", "ay",
            "alen", "awidth"]
    if len(stage) == 0:
        return np.zeros((0, len(cols)), dtype=np.int64)
    return np.stack([np.asarray(stage[c], dtype=np.int64) for c in cols], axis=1)


# paired configs: name -> (generator config, gmapper options, MapOptions overrides)
PAIR_CASES = {
    "c3_small": dict(gen="c3_small", args=["-p", "opp-in", "-I", "0,1000"], opts={}),
    "c3_small_nomq": dict(gen="c3_small", args=["-p", "opp-in", "-I", "0,1000", "--no-mapping-qualities"],
                          opts={"compute_mapping_qualities": False}),
    # colour-space pairs (post_sw is not on the path: --no-mapping-qualities)
    "c2p_small": dict(gen="c2p_small", args=["-p", "opp-in", "-I", "0,1000", "--no-mapping-qualities"],
                      opts={"compute_mapping_qualities": False}),
    # colour-space pairs with mapping qualities (post_sw on every member of a pair)
    "c2p_small_mq": dict(gen="c2p_small", args=["-p", "opp-in", "-I", "0,1000"], opts={}),
    # the paired option sets that look at the mate's region counts (gmapper.c:2659-2662):
    # use_mp_region_counts 1 (-n 4 without half-paired), 2 (-n 3), 3 (-n 3 without half-paired); and -n 2 (no regions)
    "c3_small_nohp": dict(gen="c3_small", args=["-p", "opp-in", "-I", "0,1000", "--no-half-paired"], opts={},
                          pair=dict(half_paired=False)),
    "c3_small_n3": dict(gen="c3_small", args=["-p", "opp-in", "-I", "0,1000", "-n", "3"], opts={"match_mode": 3}),
    "c3_small_n3_nohp": dict(gen="c3_small", args=["-p", "opp-in", "-I", "0,1000", "-n", "3", "--no-half-paired"],
                             opts={"match_mode": 3}, pair=dict(half_paired=False)),
    "c3_small_n2": dict(gen="c3_small", args=["-p", "opp-in", "-I", "0,1000", "-n", "2"], opts={"match_mode": 2}),
}


class PairCase:
    """reads 2k and 2k+1 are the mates of pair k (files -1 / -2 of gmapper)"""

    def __init__(self, name: str):
        cfg = gen_synth.CONFIGS[PAIR_CASES[name]["gen"]]
        self.name = name
        self.colour = bool(cfg.get("colour"))
        self.binary = "gmapper-cs" if self.colour else "gmapper-ls"
        self.contigs = gen_synth.make_genome(**cfg["genome"])
        self.contig_codes = [_LS_CODE[s] for _, s in self.contigs]
        self.contig_names = [n for n, _ in self.contigs]
        self.m1, self.m2 = gen_synth.simulate_pairs(self.contigs, **cfg["reads"])
        self.n_pairs = len(self.m1)
        if self.colour:   # SOLiD encoding of every mate: primer base T + colours
            rng = np.random.default_rng(99)
            self.m1 = [(n, np.frombuffer(gen_synth.letters_to_colour_read(s, rng, 0.02), dtype=np.uint8)) for n, s in self.m1]
            self.m2 = [(n, np.frombuffer(gen_synth.letters_to_colour_read(s, rng, 0.02), dtype=np.uint8)) for n, s in self.m2]
        inter = [r for pair in zip(self.m1, self.m2) for r in pair]
        if self.colour:
            seqs = [r[1] for r in inter]
            self.read_len = np.array([s.size - 1 for s in seqs], dtype=np.int32)
            self.initbp = np.array([_LS_CODE[s[0]] for s in seqs], dtype=np.int8)
            self.stride = int((self.read_len.max() + 7) // 8)
            self.packed = np.stack([_pack_codes(_CS_CODE[s[1:]].astype(np.uint32), self.stride) for s in seqs])
            self.scores = CS_DEFAULT_SCORES
        else:
            self.read_len = np.array([r[1].size for r in inter], dtype=np.int32)
            self.stride = int((self.read_len.max() + 7) // 8)
            self.packed = np.stack([_pack_codes(_LS_CODE[r[1]].astype(np.uint32), self.stride) for r in inter])
            self.initbp = None
            self.scores = LS_DEFAULT_SCORES
        self.seeds = S.load_default_seeds()
        self.total_len = int(sum(c.size for c in self.contig_codes))

    def write_fasta(self, d: str):
        os.makedirs(d, exist_ok=True)
        gen_synth.write_fasta(os.path.join(d, "genome.fa"), self.contigs, width=80)
        gen_synth.write_fasta(os.path.join(d, "m1.fa"), self.m1)
        gen_synth.write_fasta(os.path.join(d, "m2.fa"), self.m2)
