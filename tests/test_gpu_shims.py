"""The per-call symbols of integration/sw_shims.cpp -- sw_vector, sw_gapless, sw_full_ls, sw_full_cs with the
reference's C++ linkage and signatures -- called one window at a time through integration/_build/libshrimp_shims.so
(mangled names, ctypes) and checked against the CPU oracle, which is itself pinned on the reference objects
(tests/test_oracle_sw_*.py)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from fullcases import make_full_cases  # noqa: E402
from shrimp_b200.api import CS_DEFAULT_SCORES, LS_DEFAULT_SCORES  # noqa: E402
from swcases import make_vector_cases  # noqa: E402

SO = os.path.join(ROOT, "integration", "_build", "libshrimp_shims.so")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(SO), reason="integration/_build not built")]


class Sfr(C.Structure):   # struct sw_full_results, common/sw-full-common.h:13-48 (152 bytes on x86-64)
    _fields_ = [(n, C.c_int) for n in ("read_start", "rmapped", "genome_start", "gmapped", "matches", "mismatches",
                                       "insertions", "deletions", "score", "posterior_score", "pct_posterior_score")] + \
               [("dbalign", C.c_char_p), ("qralign", C.c_char_p), ("qual", C.c_char_p), ("posterior", C.c_double),
                ("mqv", C.c_int)] + [(n, C.c_double) for n in ("z0", "z1", "z2", "z3", "pr_top", "pr_missed", "ins_denom")] + \
               [("crossovers", C.c_int), ("dup", C.c_bool), ("in_use", C.c_bool)]


class Anchor(C.Structure):   # struct anchor, gmapper-definitions.h:67-75
    _fields_ = [("x", C.c_longlong), ("y", C.c_longlong), ("length", C.c_int), ("width", C.c_int), ("weight", C.c_int),
                ("cn", C.c_int), ("score", C.c_int)]


@pytest.fixture(scope="module")
def shims():
    assert C.sizeof(Sfr) == 152 and C.sizeof(Anchor) == 40
    L = C.CDLL(SO)
    vp, i = C.c_void_p, C.c_int
    f = {}
    f["vsetup"] = L["_Z15sw_vector_setupiiiiiiiiib"]
    f["vsetup"].argtypes = [i] * 9 + [C.c_bool]
    f["vector"] = L["_Z9sw_vectorPjiiS_iS_ib"]
    f["vector"].argtypes = [vp, i, i, vp, i, vp, i, C.c_bool]
    f["vstats"] = L["_Z15sw_vector_statsPmS_Pd"]
    f["vstats"].argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    f["gsetup"] = L["_Z16sw_gapless_setupiib"]
    f["gsetup"].argtypes = [i, i, C.c_bool]
    f["gapless"] = L["_Z10sw_gaplessPjiS_iiiS_ib"]
    f["gapless"].argtypes = [vp, i, vp, i, i, i, vp, i, C.c_bool]
    f["lsetup"] = L["_Z16sw_full_ls_setupiiiiiiiibi"]
    f["lsetup"].argtypes = [i] * 8 + [C.c_bool, i]
    f["full_ls"] = L["_Z10sw_full_lsPjiiS_iiiP15sw_full_resultsbP6anchorii"]
    f["full_ls"].argtypes = [vp, i, i, vp, i, i, i, C.POINTER(Sfr), C.c_bool, C.POINTER(Anchor), i, i]
    f["csetup"] = L["_Z16sw_full_cs_setupiiiiiiiiibii"]
    f["csetup"].argtypes = [i] * 9 + [C.c_bool, i, i]
    f["full_cs"] = L["_Z10sw_full_csPjiiS_iiiP15sw_full_resultsbbP6anchoriiPi"]
    f["full_cs"].argtypes = [vp, i, i, vp, i, i, i, C.POINTER(Sfr), C.c_bool, C.c_bool, C.POINTER(Anchor), i, i, vp]
    for k in ("vector", "gapless", "vsetup", "gsetup", "lsetup", "csetup"):
        f[k].restype = i
    for k in ("vstats", "full_ls", "full_cs"):
        f[k].restype = None
    return f


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("colour", [False, True])
def test_sw_vector_per_call(shims, colour):
    sc = CS_DEFAULT_SCORES if colour else LS_DEFAULT_SCORES
    cases = make_vector_cases(seed=31 + colour, n_tasks=300, rlen_range=(20, 110), colour=colour)
    vm = sc.match + sc.crossover if colour else sc.mismatch   # what f1_setup passes (gmapper.c:2935)
    assert shims["vsetup"](400, 200, sc.a_gap_open, sc.a_gap_ext, sc.b_gap_open, sc.b_gap_ext, sc.match, vm,
                           int(colour), True) == 0
    vsc = type(sc)(sc.match, vm, sc.a_gap_open, sc.a_gap_ext, sc.b_gap_open, sc.b_gap_ext, sc.crossover)
    cells = 0
    for t in range(300):
        rd = np.ascontiguousarray(cases["reads"][t])
        got = shims["vector"](_p(cases["genome"]), int(cases["goff"][t]), int(cases["glen"][t]), _p(rd),
                              int(cases["rlen"][t]), _p(cases["genome_ls"]), int(cases["initbp"][t]) if colour else -1,
                              False)
        want = oracle.sw_vector(cases["genome"], cases["goff"][t], cases["glen"][t], rd, cases["rlen"][t], vsc,
                                genome_ls=cases["genome_ls"], initbp=int(cases["initbp"][t]) if colour else -1)
        assert got == want, (t, got, want)
        cells += int(cases["glen"][t]) * int(cases["rlen"][t])
    inv, cl, secs = C.c_uint64(), C.c_uint64(), C.c_double()
    shims["vstats"](C.byref(inv), C.byref(cl), C.byref(secs))
    assert inv.value == 300 and cl.value == cells and secs.value > 0


@pytest.mark.parametrize("colour", [False, True])
def test_sw_gapless_per_call(shims, colour):
    sc = CS_DEFAULT_SCORES if colour else LS_DEFAULT_SCORES
    cases = make_vector_cases(seed=41 + colour, n_tasks=300, rlen_range=(18, 60), colour=colour)
    vm = sc.match + sc.crossover if colour else sc.mismatch
    shims["vsetup"](400, 200, sc.a_gap_open, sc.a_gap_ext, sc.b_gap_open, sc.b_gap_ext, sc.match, vm, int(colour), True)
    shims["gsetup"](sc.match, vm, True)
    rng = np.random.default_rng(5)
    glen = 20000
    for t in range(300):
        rd = np.ascontiguousarray(cases["reads"][t])
        rl = int(cases["rlen"][t])
        g_idx = int(cases["goff"][t]) + int(rng.integers(0, 10)) if t % 7 else int(rng.integers(0, 8))
        r_idx = int(rng.integers(0, rl))
        ib = int(cases["initbp"][t]) if colour else -1
        got = shims["gapless"](_p(cases["genome"]), glen, _p(rd), rl, g_idx, r_idx, _p(cases["genome_ls"]), ib, False)
        want = oracle.sw_gapless(cases["genome"], glen, rd, rl, g_idx, r_idx,
                                 type(sc)(sc.match, vm, sc.a_gap_open, sc.a_gap_ext, sc.b_gap_open, sc.b_gap_ext,
                                          sc.crossover), genome_ls=cases["genome_ls"], initbp=ib)
        assert got == want, (t, got, want)


def _tuple(s, colour):
    return (s.read_start, s.rmapped, s.genome_start, s.gmapped, s.matches, s.mismatches, s.insertions, s.deletions,
            s.score, s.crossovers if colour else 0, s.dbalign or b"", s.qralign or b"")


def _trim(t):
    return t[:10] + (t[10].split(b"\0")[0], t[11].split(b"\0")[0])


@pytest.mark.parametrize("local", [0, 1])
def test_sw_full_ls_per_call(shims, local):
    sc = LS_DEFAULT_SCORES
    cases = make_full_cases(seed=301 + local, n=200, rlen_range=(25, 80))
    shims["vsetup"](400, 200, sc.a_gap_open, sc.a_gap_ext, sc.b_gap_open, sc.b_gap_ext, sc.match, sc.mismatch, 0, True)
    shims["lsetup"](400, 200, sc.a_gap_open, sc.a_gap_ext, sc.b_gap_open, sc.b_gap_ext, sc.match, sc.mismatch, True, 8)
    n = 0
    for t, c in enumerate(cases):
        thresh = int(c["rlen"] * sc.match * 0.5)
        v = oracle.sw_vector(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], sc)
        if v < thresh:
            continue
        a = Anchor(c["anchor"][0], c["anchor"][1], c["anchor"][2], c["anchor"][3], 1, 0, 0)
        s = Sfr()
        shims["full_ls"](_p(c["genome"]), c["goff"], c["glen"], _p(c["read"]), c["rlen"], thresh, v, C.byref(s),
                         bool(c["revcmpl"]), C.byref(a), 1, local)
        want = _trim(oracle.sw_full_ls(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], thresh, v, c["revcmpl"],
                                       c["anchor"], 8, local, sc))
        if want[8] <= 0:
            continue
        assert _tuple(s, False) == want, (t, _tuple(s, False), want)
        n += 1
    assert n > 60


@pytest.mark.parametrize("xover", [False, True])
def test_sw_full_cs_per_call(shims, xover):
    sc = CS_DEFAULT_SCORES
    cases = make_full_cases(seed=311 + xover, n=200, colour=True, rlen_range=(25, 60))
    shims["vsetup"](400, 200, sc.a_gap_open, sc.a_gap_ext, sc.b_gap_open, sc.b_gap_ext, sc.match,
                    sc.match + sc.crossover, 1, True)
    shims["csetup"](400, 200, sc.a_gap_open, sc.a_gap_ext, sc.b_gap_open, sc.b_gap_ext, sc.match, sc.mismatch,
                    sc.crossover, True, 8, 0)
    rng = np.random.default_rng(9)
    n = 0
    for t, c in enumerate(cases):
        thresh = int(c["rlen"] * sc.match * 0.4)
        xs = rng.integers(2 * sc.crossover, 0, size=c["rlen"]).astype(np.int32) if xover else None
        a = Anchor(c["anchor"][0], c["anchor"][1], c["anchor"][2], c["anchor"][3], 1, 0, 0)
        s = Sfr()
        shims["full_cs"](_p(c["genome"]), c["goff"], c["glen"], _p(c["read"]), c["rlen"], c["initbp"], thresh,
                         C.byref(s), bool(c["revcmpl"]), False, C.byref(a), 1, 0, _p(xs))
        want = _trim(oracle.sw_full_cs(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], c["initbp"], thresh,
                                       c["revcmpl"], c["anchor"], 8, 0, 0, xs, sc))
        assert _tuple(s, True) == want, (t, _tuple(s, True), want)
        n += want[8] > 0
    assert n > 50
