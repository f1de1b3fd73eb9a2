"""GPU parity: the DPX sw_vector kernel (through the C ABI) vs the CPU oracle, bit-exact."""
import os

import numpy as np
import pytest

import oracle
from swcases import make_vector_cases
from test_oracle_sw_vector import SCORE_SETS, _oracle_scores

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gpu_scores(ctx, cases, sc, colour, dblen=400, qrlen=200):
    ctx.sw_setup(dblen, qrlen, sc, use_colours=colour)
    return ctx.sw_vector(cases["genome"], cases["goff"], cases["glen"], cases["reads"], cases["read_idx"],
                         cases["rlen"], cases["genome_ls"] if colour else None,
                         cases["initbp"] if colour else None)


@pytest.mark.parametrize("name", sorted(SCORE_SETS))
def test_gpu_matches_golden_vectors(gpu_ctx, name):
    gold = np.load(os.path.join(GOLD, f"sw_vector_{name}.npz"))
    sc, colour = SCORE_SETS[name]
    cases = make_vector_cases(seed=int(gold["seed"]), n_tasks=int(gold["n_tasks"]), colour=colour)
    got = _gpu_scores(gpu_ctx, cases, sc, colour)
    assert np.array_equal(got, gold["scores"])


@pytest.mark.parametrize("name", sorted(SCORE_SETS))
@pytest.mark.parametrize("rl", [(1, 9), (20, 64), (50, 50), (65, 130), (150, 200)])
def test_gpu_matches_oracle_random(gpu_ctx, name, rl):
    """ragged lengths: single strip, exact-fit strip, multi-strip; odd task count exercises the lone lane"""
    sc, colour = SCORE_SETS[name]
    if colour and rl[0] < 2:
        rl = (2, 9)
    cases = make_vector_cases(seed=rl[0] * 7 + rl[1], n_tasks=1001, rlen_range=rl, colour=colour)
    got = _gpu_scores(gpu_ctx, cases, sc, colour)
    want = _oracle_scores(cases, sc, colour)
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, (bad[:10], got[bad[:10]], want[bad[:10]])


def test_gpu_edge_cases(gpu_ctx):
    sc, _ = SCORE_SETS["ls_default"]
    cases = make_vector_cases(seed=3, n_tasks=64, rlen_range=(30, 40))
    # window of length 1, window at the very end of the genome, all-N read
    cases["glen"][0] = 1
    cases["goff"][1] = cases["genome"].size * 8 - cases["glen"][1]
    cases["reads"][2, :] = 0xFFFFFFFF
    got = _gpu_scores(gpu_ctx, cases, sc, False)
    want = _oracle_scores(cases, sc, False)
    assert np.array_equal(got, want)
    # empty batch
    e = np.zeros(0, dtype=np.int32)
    out = gpu_ctx.sw_vector(cases["genome"], e.astype(np.uint32), e, cases["reads"], e, e)
    assert out.size == 0


def test_gpu_rejects_too_long_reads(gpu_ctx):
    """same guard as sw-vector.c:393: match * qrlen must stay below 32768"""
    import shrimp_b200
    sc, _ = SCORE_SETS["ls_default"]
    with pytest.raises(shrimp_b200.ShrimpGpuError):
        gpu_ctx.sw_setup(5000, 3277, sc)


def test_gpu_large_uniform_batch_properties(gpu_ctx):
    """C1-shaped batch (50 bp vs 70 bp windows): score bounds + perfect implants score match*rlen."""
    sc, _ = SCORE_SETS["ls_default"]
    rng = np.random.default_rng(11)
    n = 200_000
    glen_total = 1 << 20
    from shrimp_b200.api import _pack_codes
    g = rng.integers(0, 4, size=glen_total).astype(np.uint32)
    off = rng.integers(0, glen_total - 70, size=n).astype(np.uint32)
    idx = (off[:, None] + 10 + np.arange(50)[None, :])
    rd = g[idx]
    stride = 7
    sh = (4 * np.arange(8, dtype=np.uint32))[None, None, :]
    buf = np.zeros((n, stride * 8), dtype=np.uint32)
    buf[:, :50] = rd
    reads = np.bitwise_or.reduce(buf.reshape(n, stride, 8) << sh, axis=2).astype(np.uint32)
    gpu_ctx.sw_setup(200, 100, sc)
    got = gpu_ctx.sw_vector(_pack_codes(g), off, np.full(n, 70, np.int32), reads, np.arange(n, dtype=np.int32),
                            np.full(n, 50, np.int32))
    assert (got == 500).all()
