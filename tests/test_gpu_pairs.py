"""GPU parity of the paired chunk pipeline (shrimp_gpu_map_pairs = handle_readpair): every mapped SAM record's
hot-path fields vs golden vectors from the reference gmapper (default and --no-mapping-qualities), and the pair
records themselves vs the CPU oracle."""
import os

import numpy as np
import pytest

from mapcases import GOLD, PAIR_CASES, PairCase
from oracle import pipeline as op
from shrimp_b200 import align
from shrimp_b200.api import MapParams, _pack_codes, auto_list_cutoff

pytestmark = pytest.mark.gpu


def run_gpu_pairs(ctx, case, pair=None, **over):
    ctx.sw_setup(1400, 1000, case.scores, use_colours=case.colour, anchor_width=8)
    ctx.load_genome([_pack_codes(c.astype(np.uint32)) for c in case.contig_codes], [c.size for c in case.contig_codes],
                    colour_space=case.colour)
    ctx.build_index(case.seeds)
    over.setdefault("match_mode", 4)
    params = MapParams(list_cutoff=auto_list_cutoff(case.total_len, 12), **over)
    return ctx.map_pairs(params, case.scores, case.packed, case.read_len, initbp=case.initbp, **(pair or {}))


def gpu_records(case, res):
    lens = [c.size for c in case.contig_codes]

    def fields(h, rl, gl):
        e = res.edits[int(h["edit_off"]): int(h["edit_off"]) + int(h["edit_len"])]
        return align.sam_fields(h, e, rl, gl, case.colour)

    pair_hits = [(res.hits[int(p["hit_idx"][0])], res.hits[int(p["hit_idx"][1])]) for p in res.pairs]
    return op.pair_sam_records(pair_hits, res.pairs["pair_idx"], res.hits[res.n_paired_hits:], lens, case.read_len,
                               fields, case.n_pairs)


@pytest.mark.parametrize("name", sorted(PAIR_CASES))
def test_pairs_match_reference_golden(gpu_ctx, name):
    gold = np.load(os.path.join(GOLD, f"pairs_{name}.npz"))
    case = PairCase(name)
    res = run_gpu_pairs(gpu_ctx, case, pair=PAIR_CASES[name].get("pair"), **PAIR_CASES[name]["opts"])
    recs = gpu_records(case, res)
    ints = np.array([[r[0], r[1], r[2], r[3], r[4], r[6], r[7], r[8], r[9], r[10]] for r in recs],
                    dtype=np.int64).reshape(-1, 10)
    cig = np.array([r[5] for r in recs])
    assert ints.shape == gold["recs"].shape
    bad = np.nonzero((ints != gold["recs"]).any(axis=1))[0]
    assert bad.size == 0, (bad[:5], ints[bad[:5]], gold["recs"][bad[:5]])
    assert np.array_equal(cig, gold["cigars"])


def test_pairs_match_oracle_records(gpu_ctx):
    """pair scores, keys and insert sizes (readpair_compute_paired_hit) and the per-class counts vs the oracle"""
    from test_oracle_pairs import run_oracle_pairs
    case = PairCase("c3_small")
    res = run_gpu_pairs(gpu_ctx, case)
    _, (ph, pinfo, nper, uh, nunp, st) = run_oracle_pairs(case)
    assert np.array_equal(res.n_pairs_per_pair, nper)
    assert np.array_equal(res.n_unpaired_per_read, nunp)
    got = np.stack([res.pairs[k] for k in ("pair_idx", "score", "score_max", "key", "insert_size")], axis=1)
    assert np.array_equal(got.astype(np.int64), pinfo.astype(np.int64))
    assert res.stats["vector_calls"] == st["vector_calls"]
    assert res.stats["vector_cells"] == st["vector_cells"]
    assert res.stats["full_cells"] == st["full_cells"]


def test_pairs_reject_invalid_modes(gpu_ctx):
    from shrimp_b200._lib import ShrimpGpuError
    case = PairCase("c3_small")
    gpu_ctx.sw_setup(1400, 1000, case.scores)
    gpu_ctx.load_genome([_pack_codes(c.astype(np.uint32)) for c in case.contig_codes], [c.size for c in case.contig_codes])
    gpu_ctx.build_index(case.seeds)
    with pytest.raises(ShrimpGpuError):   # gmapper.c:2495-2499
        gpu_ctx.map_pairs(MapParams(list_cutoff=1000, match_mode=1), case.scores, case.packed[:4], case.read_len[:4])
    res = gpu_ctx.map_pairs(MapParams(list_cutoff=1000, match_mode=4), case.scores, case.packed[:0], case.read_len[:0])
    assert len(res.pairs) == 0 and len(res.hits) == 0


@pytest.mark.parametrize("name", ["c3_small_nohp", "c3_small_n3", "c3_small_n3_nohp"])
def test_mate_pair_region_counts_with_small_slabs(gpu_ctx, name, monkeypatch):
    """the option sets that look at the mate's region counts, with a candidate slab so small that strands overflow
    into the launches that rebuild the pair's region tables for a single strand (global slabs)"""
    monkeypatch.setenv("SHRIMP_SCAN_CTA_CAP", "32")
    monkeypatch.setenv("SHRIMP_SCAN_WIN", "128")
    gold = np.load(os.path.join(GOLD, f"pairs_{name}.npz"))
    case = PairCase(name)
    res = run_gpu_pairs(gpu_ctx, case, pair=PAIR_CASES[name].get("pair"), **PAIR_CASES[name]["opts"])
    assert res.stats["scan_global_strands"] > 0
    recs = gpu_records(case, res)
    ints = np.array([[r[0], r[1], r[2], r[3], r[4], r[6], r[7], r[8], r[9], r[10]] for r in recs],
                    dtype=np.int64).reshape(-1, 10)
    assert ints.shape == gold["recs"].shape and np.array_equal(ints, gold["recs"])
    assert np.array_equal(np.array([r[5] for r in recs]), gold["cigars"])
