"""Pins the oracle's sw_vector / sw_gapless / window hash against the reference's own objects
(oracle/_ref/libshrimp_ref.so = common/sw-vector.c etc. compiled in place) and against the
committed golden vectors (tests/golden/sw_vector_*.npz, made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import oracle
from shrimp_b200.api import CS_DEFAULT_SCORES, LS_DEFAULT_SCORES, Scores
from swcases import make_vector_cases

GOLD = os.path.join(os.path.dirname(__file__), "golden")

SCORE_SETS = {
    "ls_default": (LS_DEFAULT_SCORES, False),
    "ls_samegap": (Scores(10, -15, -33, -7, -33, -7, 0), False),
    "ls_odd": (Scores(7, -11, -20, -5, -13, -2, 0), False),
    # colour space: mismatch passed to sw_vector_setup is match + crossover (gmapper.c:2935)
    "cs_default": (Scores(10, 10 - 20, -33, -7, -33, -3, -20), True),
}


def _oracle_scores(cases, sc, colour):
    out = np.zeros(cases["goff"].size, dtype=np.int32)
    for t in range(out.size):
        out[t] = oracle.sw_vector(cases["genome"], cases["goff"][t], cases["glen"][t], cases["reads"][t],
                                  cases["rlen"][t], sc, cases["genome_ls"] if colour else None,
                                  cases["initbp"][t] if colour else -1)
    return out


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", sorted(SCORE_SETS))
def test_oracle_matches_reference_objects(name):
    sc, colour = SCORE_SETS[name]
    cases = make_vector_cases(seed=100 + len(name), n_tasks=1500, colour=colour)
    ref = oracle.RefSw(400, 200, sc, colour)
    got = _oracle_scores(cases, sc, colour)
    for t in range(got.size):
        want = ref.sw_vector(cases["genome"], cases["goff"][t], cases["glen"][t], cases["reads"][t],
                             cases["rlen"][t], cases["genome_ls"] if colour else None,
                             cases["initbp"][t] if colour else -1)
        assert got[t] == want, (name, t)


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_gapless_and_hash_match_reference():
    sc = LS_DEFAULT_SCORES
    cases = make_vector_cases(seed=7, n_tasks=800)
    ccases = make_vector_cases(seed=8, n_tasks=400, colour=True)
    ref = oracle.RefSw(400, 200, sc, False)
    rng = np.random.default_rng(5)
    L = oracle.ref_lib()
    for t in range(800):
        g_idx = int(cases["goff"][t] + rng.integers(0, cases["glen"][t]))
        r_idx = int(rng.integers(0, cases["rlen"][t]))
        glen_total = cases["genome"].size * 8
        a = oracle.sw_gapless(cases["genome"], glen_total, cases["reads"][t], cases["rlen"][t], g_idx, r_idx, sc)
        b = ref.sw_gapless(cases["genome"], glen_total, cases["reads"][t], cases["rlen"][t], g_idx, r_idx)
        assert a == b
        h1 = oracle.hash_genome_window(cases["genome"], cases["goff"][t], cases["glen"][t])
        h2 = L.ref_hash_genome_window(cases["genome"].ctypes.data, int(cases["goff"][t]), int(cases["glen"][t]))
        assert h1 == h2
    for t in range(400):
        g_idx = int(ccases["goff"][t] + rng.integers(0, ccases["glen"][t]))
        r_idx = int(rng.integers(0, 3))
        glen_total = ccases["genome"].size * 8
        a = oracle.sw_gapless(ccases["genome"], glen_total, ccases["reads"][t], ccases["rlen"][t], g_idx, r_idx, sc,
                              ccases["genome_ls"], ccases["initbp"][t])
        b = ref.sw_gapless(ccases["genome"], glen_total, ccases["reads"][t], ccases["rlen"][t], g_idx, r_idx,
                           ccases["genome_ls"], ccases["initbp"][t])
        assert a == b


@pytest.mark.parametrize("name", sorted(SCORE_SETS))
def test_oracle_matches_golden_vectors(name):
    path = os.path.join(GOLD, f"sw_vector_{name}.npz")
    gold = np.load(path)
    sc, colour = SCORE_SETS[name]
    cases = make_vector_cases(seed=int(gold["seed"]), n_tasks=int(gold["n_tasks"]), colour=colour)
    got = _oracle_scores(cases, sc, colour)
    assert np.array_equal(got, gold["scores"])
