"""GPU parity of the batched full Smith-Waterman entry (shrimp_gpu_sw_full_batch = sw_full_ls / sw_full_cs)
against (a) golden vectors produced by the reference's own objects and (b) the CPU oracle on seeded random
cases: global and local mode, anchor band and threshold band, indel taboo, reads long enough to reach every
ring-width class of sw_full_ring.cu and the global-scratch fall-back."""
import os

import numpy as np
import pytest

import oracle
from fullcases import make_full_cases
from shrimp_b200 import align
from shrimp_b200._lib import FullTaskC
from shrimp_b200.api import CS_DEFAULT_SCORES, LS_DEFAULT_SCORES

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _unpack(words, n):
    sh = (4 * np.arange(8, dtype=np.uint32))[None, :]
    return ((np.asarray(words)[:, None] >> sh) & 15).reshape(-1)[:n].astype(np.uint8)


def run_cases(ctx, cases, colour, sc, thresh_frac, anchor_width, local, taboo=0, vscores=None):
    """-> list of 12-tuples like oracle._sfr_tuple (None where the case is skipped)"""
    stride = max(c["read"].size for c in cases)
    reads = np.zeros((len(cases), stride), dtype=np.uint32)
    tasks = np.zeros(len(cases), dtype=FullTaskC)
    keep = []
    for t, c in enumerate(cases):
        reads[t, :c["read"].size] = c["read"]
        thresh = int(c["rlen"] * sc.match * thresh_frac)
        v = 0
        if not colour:
            v = vscores[t]
            if v < thresh:
                continue
        keep.append(t)
        tasks[t] = (c["goff"], c["glen"], t, c["rlen"], thresh, v, c["revcmpl"], c["anchor"][0], c["anchor"][1],
                    c["anchor"][2], c["anchor"][3], c["initbp"])
    keep = np.array(keep, dtype=np.int64)
    ctx.sw_setup(400, 200, sc, use_colours=colour, anchor_width=anchor_width, indel_taboo_len=taboo)
    res, edits = ctx.sw_full(cases[0]["genome"], reads, tasks[keep], local=local)
    gcodes = _unpack(cases[0]["genome"], cases[0]["genome"].size * 8)
    out = [None] * len(cases)
    for r, t in zip(res, keep):
        c = cases[int(t)]
        e = edits[int(r["edit_off"]): int(r["edit_off"]) + int(r["edit_len"])]
        rcodes = _unpack(c["read"], c["rlen"])
        if int(r["score"]) > 0 or int(r["edit_len"]) > 0:
            if colour:
                db, qr = align.align_strings_cs(e, gcodes, int(r["genome_start"]), rcodes, c["initbp"],
                                                int(r["read_start"]))
            else:
                db, qr = align.align_strings(e, gcodes, int(r["genome_start"]), rcodes, int(r["read_start"]))
        else:
            db, qr = b"", b""
        out[int(t)] = (int(r["read_start"]), int(r["rmapped"]), int(r["genome_start"]), int(r["gmapped"]),
                       int(r["matches"]), int(r["mismatches"]), int(r["insertions"]), int(r["deletions"]),
                       int(r["score"]), int(r["crossovers"]), db, qr)
    return out


def _vscores(cases, sc):
    return [oracle.sw_vector(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], sc) for c in cases]


@pytest.mark.parametrize("colour", [False, True])
def test_sw_full_matches_reference_golden(gpu_ctx, colour):
    sc = CS_DEFAULT_SCORES if colour else LS_DEFAULT_SCORES
    gold = np.load(os.path.join(GOLD, f"sw_full_{'cs' if colour else 'ls'}.npz"))
    cases = make_full_cases(seed=500 + colour, n=400, colour=colour, rlen_range=(25, 60))
    got = run_cases(gpu_ctx, cases, colour, sc, 0.4 if colour else 0.5, 8, False,
                    vscores=None if colour else _vscores(cases, sc))
    n = 0
    for t in range(len(cases)):
        if gold["ints"][t, 8] <= 0:
            continue
        g = got[t]
        assert g is not None
        assert list(g[:10]) == gold["ints"][t].tolist(), (t, g, gold["ints"][t])
        assert g[10] == gold["db"][t] and g[11] == gold["qr"][t], (t, g, gold["db"][t], gold["qr"][t])
        n += 1
    assert n > 150


def _trim(t):
    return t[:10] + (t[10].split(b"\0")[0], t[11].split(b"\0")[0])


@pytest.mark.parametrize("local", [0, 1])
@pytest.mark.parametrize("anchor_width", [8, -1])
@pytest.mark.parametrize("rlen_range", [(25, 80), (120, 200)])
def test_sw_full_ls_matches_oracle(gpu_ctx, local, anchor_width, rlen_range):
    sc = LS_DEFAULT_SCORES
    cases = make_full_cases(seed=131 + local + (anchor_width > 0) + rlen_range[0], n=600, rlen_range=rlen_range)
    vs = _vscores(cases, sc)
    got = run_cases(gpu_ctx, cases, False, sc, 0.5, anchor_width, bool(local), vscores=vs)
    n = 0
    for t, c in enumerate(cases):
        if got[t] is None:
            continue
        thresh = int(c["rlen"] * sc.match * 0.5)
        want = _trim(oracle.sw_full_ls(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], thresh, vs[t],
                                       c["revcmpl"], c["anchor"], anchor_width, local, sc))
        if want[8] <= 0:
            continue   # nothing scores > 0: the reference's result is discarded upstream
        assert got[t] == want, (t, c["anchor"], got[t], want)
        n += 1
    assert n > 200


@pytest.mark.parametrize("local", [0, 1])
@pytest.mark.parametrize("taboo", [0, 3])
@pytest.mark.parametrize("rlen_range", [(25, 60), (100, 180)])
def test_sw_full_cs_matches_oracle(gpu_ctx, local, taboo, rlen_range):
    sc = CS_DEFAULT_SCORES
    cases = make_full_cases(seed=177 + local + taboo + rlen_range[0], n=400, colour=True, rlen_range=rlen_range)
    got = run_cases(gpu_ctx, cases, True, sc, 0.4, 8, bool(local), taboo=taboo)
    n = 0
    for t, c in enumerate(cases):
        thresh = int(c["rlen"] * sc.match * 0.4)
        want = _trim(oracle.sw_full_cs(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], c["initbp"], thresh,
                                       c["revcmpl"], c["anchor"], 8, taboo, local, None, sc))
        assert got[t] == want, (t, c["anchor"], got[t], want)
        n += want[8] > 0
    assert n > 100


def test_sw_full_cs_threshold_band(gpu_ctx):
    """anchor_width < 0: the wide threshold band (sw-full-cs.c:285-303) -> wide ring classes / fall-back"""
    sc = CS_DEFAULT_SCORES
    cases = make_full_cases(seed=211, n=300, colour=True, rlen_range=(40, 150))
    got = run_cases(gpu_ctx, cases, True, sc, 0.4, -1, False)
    for t, c in enumerate(cases):
        thresh = int(c["rlen"] * sc.match * 0.4)
        want = _trim(oracle.sw_full_cs(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], c["initbp"], thresh,
                                       c["revcmpl"], c["anchor"], -1, 0, 0, None, sc))
        assert got[t] == want, (t, got[t], want)
