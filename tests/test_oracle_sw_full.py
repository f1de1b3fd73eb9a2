"""Pins the oracle's sw_full_ls / sw_full_cs restatements against the reference's own objects
(oracle/_ref/libshrimp_ref.so) on random cases: score, coordinates, counts, crossovers and both
alignment strings; plus committed golden vectors."""
import os

import numpy as np
import pytest

import oracle
from fullcases import make_full_cases
from shrimp_b200.api import CS_DEFAULT_SCORES, LS_DEFAULT_SCORES, Scores

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CS_VEC = Scores(10, 10 - 20, -33, -7, -33, -3, -20)


def _trim(t):
    return t[:10] + (t[10].split(b"\0")[0], t[11].split(b"\0")[0])


def ls_case_result(c, local, anchor_width, fn_vec, fn_full):
    sc = LS_DEFAULT_SCORES
    v = fn_vec(c)
    thresh = int(c["rlen"] * sc.match * 0.5)
    if v < thresh:
        return None
    return _trim(fn_full(c, thresh, v))


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("local", [0, 1])
@pytest.mark.parametrize("anchor_width", [8, -1])
def test_sw_full_ls_matches_reference(local, anchor_width):
    sc = LS_DEFAULT_SCORES
    cases = make_full_cases(seed=31 + local + (anchor_width > 0), n=1500)
    ref = oracle.RefFull(400, 200, sc, anchor_width)
    refv = oracle.RefSw(400, 200, sc, False)
    n_checked = 0
    for c in cases:
        v = refv.sw_vector(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"])
        thresh = int(c["rlen"] * sc.match * 0.5)
        if v < thresh:
            continue
        a = _trim(oracle.sw_full_ls(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], thresh, v, c["revcmpl"],
                                    c["anchor"], anchor_width, local, sc))
        b = _trim(ref.sw_full_ls(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], thresh, v, c["revcmpl"],
                                 c["anchor"], local))
        if b[8] <= 0:
            continue   # the reference walks stale memory when nothing scores > 0; results are discarded upstream
        assert a == b, (c["anchor"], a, b)
        n_checked += 1
    assert n_checked > 500


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("local", [0, 1])
@pytest.mark.parametrize("taboo", [0, 3])
@pytest.mark.parametrize("perpos", [False, True])
def test_sw_full_cs_matches_reference(local, taboo, perpos):
    sc = CS_DEFAULT_SCORES
    cases = make_full_cases(seed=77 + local + taboo + perpos, n=800, colour=True, rlen_range=(25, 60))
    ref = oracle.RefFull(400, 200, sc, 8, colour=True, taboo=taboo)
    rng = np.random.default_rng(3)
    n_checked = 0
    for c in cases:
        thresh = int(c["rlen"] * sc.match * 0.4)
        xs = None
        if perpos:
            xs = np.ascontiguousarray(rng.integers(-40, 0, size=c["rlen"]).astype(np.int32))
        a = _trim(oracle.sw_full_cs(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], c["initbp"], thresh,
                                    c["revcmpl"], c["anchor"], 8, taboo, local, xs, sc))
        b = _trim(ref.sw_full_cs(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], c["initbp"], thresh,
                                 c["revcmpl"], c["anchor"], local, xs))
        assert a == b, (c["anchor"], a, b)
        n_checked += b[8] > 0
    assert n_checked > 200


def _golden_results(colour):
    sc = CS_DEFAULT_SCORES if colour else LS_DEFAULT_SCORES
    cases = make_full_cases(seed=500 + colour, n=400, colour=colour, rlen_range=(25, 60))
    out = []
    for c in cases:
        if colour:
            thresh = int(c["rlen"] * sc.match * 0.4)
            r = oracle.sw_full_cs(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], c["initbp"], thresh,
                                  c["revcmpl"], c["anchor"], 8, 0, 0, None, sc)
        else:
            v = oracle.sw_vector(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], sc)
            thresh = int(c["rlen"] * sc.match * 0.5)
            if v < thresh:
                out.append((0,) * 10 + (b"", b""))
                continue
            r = oracle.sw_full_ls(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], thresh, v, c["revcmpl"],
                                  c["anchor"], 8, 0, sc)
        out.append(_trim(r))
    return out


@pytest.mark.parametrize("colour", [False, True])
def test_sw_full_matches_golden_vectors(colour):
    gold = np.load(os.path.join(GOLD, f"sw_full_{'cs' if colour else 'ls'}.npz"))
    got = _golden_results(colour)
    ints = np.array([g[:10] for g in got], dtype=np.int64)
    keep = gold["ints"][:, 8] > 0
    assert np.array_equal(ints[keep], gold["ints"][keep])
    assert [g[10] for g, k in zip(got, keep) if k] == [x for x, k in zip(gold["db"].tolist(), keep) if k]
    assert [g[11] for g, k in zip(got, keep) if k] == [x for x, k in zip(gold["qr"].tolist(), keep) if k]
