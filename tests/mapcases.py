"""Synthetic mapping configs shared by the oracle and GPU pipeline tests (seeded, see tools/gen_synth.py)."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_synth  # noqa: E402

from shrimp_b200 import seeds as S  # noqa: E402
from shrimp_b200.api import LS_DEFAULT_SCORES, _LS_CODE, _pack_codes  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name -> (generator config, gmapper command-line options, option overrides for MapOptions)
MAP_CASES = {
    "c1_small": dict(gen="c1_small", args=[], opts={}),
    "c1_repeat": dict(gen="c1_repeat", args=[], opts={}),
}


class LsCase:
    def __init__(self, name: str):
        cfg = gen_synth.CONFIGS[MAP_CASES[name]["gen"]]
        self.name = name
        self.contigs = gen_synth.make_genome(**cfg["genome"])
        self.reads = gen_synth.simulate_reads(self.contigs, **cfg["reads"])
        self.contig_codes = [_LS_CODE[s] for _, s in self.contigs]
        self.contig_names = [n for n, _ in self.contigs]
        self.read_names = [r[0] for r in self.reads]
        self.read_len = np.array([r[1].size for r in self.reads], dtype=np.int32)
        self.stride = int((self.read_len.max() + 7) // 8)
        self.packed = np.stack([_pack_codes(_LS_CODE[r[1]].astype(np.uint32), self.stride) for r in self.reads])
        self.seeds = S.load_default_seeds()
        self.scores = LS_DEFAULT_SCORES
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
