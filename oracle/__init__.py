"""TEST INFRASTRUCTURE ONLY: ctypes loaders for the CPU oracle.

* ``liboracle.so``  -- oracle/shrimp_oracle.c, the C restatement of the reference hot path.
* ``_ref/libshrimp_ref.so`` -- the reference's own objects compiled in place from /root/reference
  (oracle/Makefile); present in the build container and shipped prebuilt to the GPU box.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import
this package.  The product (shrimp_b200/) must never do so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(HERE, "liboracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libshrimp_ref.so")
REF_GMAPPER = os.path.join(HERE, "_ref", "gmapper")


class OrcScores(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("match", "mismatch", "a_gap_open", "a_gap_ext", "b_gap_open",
                                       "b_gap_ext", "crossover")]


def build(ref: bool = True) -> None:
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    if ref and os.path.isdir("/root/reference/gmapper"):
        subprocess.run(["make", "-C", HERE, "-j8", "ref"], check=True, capture_output=True)


_orc = None
_ref = None


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def oracle_lib() -> C.CDLL:
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_LIB) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(
                os.path.join(HERE, "shrimp_oracle.c")):
            build(ref=False)
        _orc = C.CDLL(ORACLE_LIB)
        _orc.orc_hash_genome_window.restype = C.c_uint32
    return _orc


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def ref_lib() -> C.CDLL:
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_LIB)
        _ref.ref_hash_genome_window.restype = C.c_uint32
    return _ref


def scores_struct(s) -> OrcScores:
    return OrcScores(s.match, s.mismatch, s.a_gap_open, s.a_gap_ext, s.b_gap_open, s.b_gap_ext, s.crossover)


def sw_vector(genome, goff, glen, read, rlen, scores, genome_ls=None, initbp=-1) -> int:
    L = oracle_lib()
    sc = scores_struct(scores)
    return L.orc_sw_vector(_p(genome), int(goff), int(glen), _p(read), int(rlen), _p(genome_ls), int(initbp),
                           C.byref(sc))


def sw_gapless(genome, glen, read, rlen, g_idx, r_idx, scores, genome_ls=None, initbp=-1) -> int:
    L = oracle_lib()
    sc = scores_struct(scores)
    return L.orc_sw_gapless(_p(genome), int(glen), _p(read), int(rlen), int(g_idx), int(r_idx), _p(genome_ls),
                            int(initbp), C.byref(sc))


def hash_genome_window(genome, goff, glen) -> int:
    return int(oracle_lib().orc_hash_genome_window(_p(genome), int(goff), int(glen)))


class RefSw:
    """The reference's own sw_vector / sw_gapless through oracle/_ref/libshrimp_ref.so."""

    def __init__(self, dblen, qrlen, scores, use_colours=False):
        self.L = ref_lib()
        rc = self.L.ref_sw_vector_setup(dblen, qrlen, scores.a_gap_open, scores.a_gap_ext, scores.b_gap_open,
                                        scores.b_gap_ext, scores.match, scores.mismatch, int(use_colours))
        assert rc == 0
        self.L.ref_sw_gapless_setup(scores.match, scores.mismatch)

    def sw_vector(self, genome, goff, glen, read, rlen, genome_ls=None, initbp=-1) -> int:
        return self.L.ref_sw_vector(_p(genome), int(goff), int(glen), _p(read), int(rlen), _p(genome_ls), int(initbp))

    def sw_gapless(self, genome, glen, read, rlen, g_idx, r_idx, genome_ls=None, initbp=-1) -> int:
        return self.L.ref_sw_gapless(_p(genome), int(glen), _p(read), int(rlen), int(g_idx), int(r_idx),
                                     _p(genome_ls), int(initbp))


class OrcSfrC(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("read_start", "rmapped", "genome_start", "gmapped", "matches", "mismatches",
                                       "insertions", "deletions", "score", "crossovers")] + \
               [("dbalign", C.c_char * 640), ("qralign", C.c_char * 640), ("qual", C.c_char * 640)]


class RefSfrC(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("read_start", "rmapped", "genome_start", "gmapped", "matches", "mismatches",
                                       "insertions", "deletions", "score", "crossovers")] + \
               [("dbalign", C.c_char * 4096), ("qralign", C.c_char * 4096)]


def _sfr_tuple(s):
    return (s.read_start, s.rmapped, s.genome_start, s.gmapped, s.matches, s.mismatches, s.insertions, s.deletions,
            s.score, s.crossovers, bytes(s.dbalign), bytes(s.qralign))


def sw_full_ls(genome, goff, glen, read, rlen, thresh, maxscore, revcmpl, anchor, anchor_width, local, scores):
    L = oracle_lib()
    sc = scores_struct(scores)
    out = OrcSfrC()
    L.orc_sw_full_ls(_p(genome), int(goff), int(glen), _p(read), int(rlen), int(thresh), int(maxscore), int(revcmpl),
                     C.c_longlong(anchor[0]), C.c_longlong(anchor[1]), int(anchor[2]), int(anchor[3]),
                     int(anchor_width), int(local), C.byref(sc), C.byref(out))
    return _sfr_tuple(out)


def sw_full_cs(genome_ls, goff, glen, read, rlen, initbp, thresh, revcmpl, anchor, anchor_width, taboo, local,
               xover_scores, scores):
    L = oracle_lib()
    sc = scores_struct(scores)
    out = OrcSfrC()
    L.orc_sw_full_cs(_p(genome_ls), int(goff), int(glen), _p(read), int(rlen), int(initbp), int(thresh), int(revcmpl),
                     C.c_longlong(anchor[0]), C.c_longlong(anchor[1]), int(anchor[2]), int(anchor[3]),
                     int(anchor_width), int(taboo), int(local), _p(xover_scores), C.byref(sc), C.byref(out), None)
    return _sfr_tuple(out)


class RefFull:
    """The reference's sw_full_ls / sw_full_cs through oracle/_ref/libshrimp_ref.so."""

    def __init__(self, dblen, qrlen, scores, anchor_width, colour=False, taboo=0):
        self.L = ref_lib()
        if colour:
            self.L.ref_sw_full_cs_setup(dblen, qrlen, scores.a_gap_open, scores.a_gap_ext, scores.b_gap_open,
                                        scores.b_gap_ext, scores.match, scores.mismatch, scores.crossover,
                                        anchor_width, taboo)
        else:
            self.L.ref_sw_full_ls_setup(dblen, qrlen, scores.a_gap_open, scores.a_gap_ext, scores.b_gap_open,
                                        scores.b_gap_ext, scores.match, scores.mismatch, anchor_width)

    def sw_full_ls(self, genome, goff, glen, read, rlen, thresh, maxscore, revcmpl, anchor, local):
        out = RefSfrC()
        self.L.ref_sw_full_ls(_p(genome), int(goff), int(glen), _p(read), int(rlen), int(thresh), int(maxscore),
                              int(revcmpl), C.c_longlong(anchor[0]), C.c_longlong(anchor[1]), int(anchor[2]),
                              int(anchor[3]), int(local), C.byref(out))
        return _sfr_tuple(out)

    def sw_full_cs(self, genome_ls, goff, glen, read, rlen, initbp, thresh, revcmpl, anchor, local, xover_scores):
        out = RefSfrC()
        self.L.ref_sw_full_cs(_p(genome_ls), int(goff), int(glen), _p(read), int(rlen), int(initbp), int(thresh),
                              int(revcmpl), C.c_longlong(anchor[0]), C.c_longlong(anchor[1]), int(anchor[2]),
                              int(anchor[3]), int(local), _p(xover_scores), C.byref(out))
        return _sfr_tuple(out)
