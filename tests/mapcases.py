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
        self.seeds = S.load_default_seeds()
        self.total_len = int(sum(c.size for c in self.contig_codes))

    def write_fasta(self, d: str):
        os.makedirs(d, exist_ok=True)
        gen_synth.write_fasta(os.path.join(d, "genome.fa"), self.contigs, width=80)
        gen_synth.write_fasta(os.path.join(d, "reads.fa"), self.reads)


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
}


class PairCase:
    """reads 2k and 2k+1 are the mates of pair k (files -1 / -2 of gmapper)"""

    def __init__(self, name: str):
        cfg = gen_synth.CONFIGS[PAIR_CASES[name]["gen"]]
        self.name = name
        self.colour = False
        self.binary = "gmapper-ls"
        self.contigs = gen_synth.make_genome(**cfg["genome"])
        self.contig_codes = [_LS_CODE[s] for _, s in self.contigs]
        self.contig_names = [n for n, _ in self.contigs]
        self.m1, self.m2 = gen_synth.simulate_pairs(self.contigs, **cfg["reads"])
        self.n_pairs = len(self.m1)
        inter = [r for pair in zip(self.m1, self.m2) for r in pair]
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
