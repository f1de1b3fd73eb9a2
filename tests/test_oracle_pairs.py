"""Pins the oracle's paired pipeline (handle_readpair restatement, oracle_pipeline.inc) against golden vectors from
the reference gmapper: every mapped SAM record of the small 2x100 bp configuration, with and without mapping
qualities (the latter changes the hot path: no posterior-based score_full, direct output)."""
import os

import numpy as np
import pytest

from mapcases import GOLD, PAIR_CASES, PairCase
from oracle import pipeline as op


def record_arrays(recs):
    ints = np.array([[r[0], r[1], r[2], r[3], r[4], r[6], r[7], r[8], r[9], r[10]] for r in recs],
                    dtype=np.int64).reshape(-1, 10)
    return ints, np.array([r[5] for r in recs])


def run_oracle_pairs(case, pair=None, **over):
    g = op.Genome(case.contig_codes, case.colour)
    ix = op.Index(g, case.seeds)
    over.setdefault("match_mode", 4)
    opts = op.MapOptions(scores=case.scores, colour_space=case.colour,
                         list_cutoff=op.auto_list_cutoff(g.total_len, 12), **over)
    ph, pinfo, nper, uh, nunp, st = op.map_pairs(g, ix, opts, case.packed, case.read_len, initbp=case.initbp,
                                                 **(pair or {}))
    recs = op.pair_sam_records(ph, pinfo[:, 0], uh, g.lens, case.read_len,
                               lambda h, rl, gl: op.sam_fields(h, rl, gl, case.colour), case.n_pairs)
    return recs, (ph, pinfo, nper, uh, nunp, st)


@pytest.mark.parametrize("name", sorted(PAIR_CASES))
def test_oracle_pairs_match_reference_golden(name):
    gold = np.load(os.path.join(GOLD, f"pairs_{name}.npz"))
    case = PairCase(name)
    recs, _ = run_oracle_pairs(case, pair=PAIR_CASES[name].get("pair"), **PAIR_CASES[name]["opts"])
    ints, cig = record_arrays(recs)
    assert ints.shape == gold["recs"].shape
    assert np.array_equal(ints, gold["recs"])
    assert np.array_equal(cig, gold["cigars"])
    # the fixture exercises every class of output: proper pairs, half-mapped pairs, multiple pairs per read pair
    assert (gold["recs"][:, 2] & 2).any()
    if PAIR_CASES[name].get("pair", {}).get("half_paired", True):
        assert (gold["recs"][:, 2] & 8).any()
