"""Pins the oracle's per-read pipeline (oracle/oracle_pipeline.inc) against golden vectors produced by
the reference gmapper binary: the hot-path SAM fields of every record (flag, contig, pos, CIGAR, AS, NM)
and the post-pass1 hit lists of its DEBUG_HIT_LIST_PASS1 build (tests/golden/make_golden.py mapping)."""
import os

import numpy as np
import pytest

from mapcases import GOLD, MAP_CASES, LsCase, stage_tuple_array
from oracle import pipeline as op


def run_oracle(case: LsCase, **over):
    g = op.Genome(case.contig_codes, case.colour)
    ix = op.Index(g, case.seeds, hflag=case.hflag)
    opts = op.MapOptions(scores=case.scores, colour_space=case.colour, list_cutoff=case.list_cutoff,
                         anchor_width=case.anchor_width, **over)
    hits, nper, stage, stats = op.map_reads(g, ix, opts, case.packed, case.read_len, initbp=case.initbp,
                                            want_stage=True, crossover_scores=case.crossover_scores, quals=case.quals)
    return g, hits, nper, stage, stats


def sam_arrays(case, g, hits):
    rows, cig = [], []
    for h in hits:
        f = op.sam_fields(h, int(case.read_len[h["read_idx"]]), int(g.lens[h["cn"]]), case.colour)
        rows.append([int(h["read_idx"]), f[0], f[1], f[2], f[4], f[5]])
        cig.append(f[3])
    return np.array(rows, dtype=np.int64).reshape(-1, 6), np.array(cig)


@pytest.mark.parametrize("name", sorted(MAP_CASES))
def test_oracle_pipeline_matches_reference_golden(name):
    gold = np.load(os.path.join(GOLD, f"map_{name}.npz"))
    case = LsCase(name)
    g, hits, nper, stage, stats = run_oracle(case, **MAP_CASES[name]["opts"])
    sam, cig = sam_arrays(case, g, hits)
    assert sam.shape == gold["sam"].shape
    assert np.array_equal(sam, gold["sam"])
    assert np.array_equal(cig, gold["cigars"])
    assert np.array_equal(stage_tuple_array(stage), gold["stage"])
    if "seq" in gold:   # colour space with mapping qualities: post_sw's corrected base calls and base qualities
        sq = [op.seq_qual_of_alignment(bytes(h["sfr"]["qralign"]).split(b"\0")[0],
                                       bytes(h["sfr"]["qual"]).split(b"\0")[0], int(h["gen_st"]) == 1,
                                       case.quals is not None) for h in hits]
        assert [a for a, _ in sq] == gold["seq"].tolist()
        assert [b for _, b in sq] == gold["qual"].tolist()


def test_oracle_index_is_sorted_csr():
    case = LsCase("c1_small")
    g = op.Genome(case.contig_codes, False)
    ix = op.Index(g, case.seeds)
    for sn, s in enumerate(case.seeds):
        lens, pos = ix.bucket_lens(sn), ix.positions(sn)
        # every position once: L - span + 1 per contig (no N in this genome)
        assert lens.sum() == sum(c.size - s.span + 1 for c in case.contig_codes) == pos.size
        starts = np.concatenate([[0], np.cumsum(lens.astype(np.int64))[:-1]]).astype(np.int64)
        big = np.nonzero(lens > 1)[0][:2000]
        for m in big:
            seg = pos[int(starts[m]):int(starts[m]) + int(lens[m])]
            assert (np.diff(seg.astype(np.int64)) > 0).all()
