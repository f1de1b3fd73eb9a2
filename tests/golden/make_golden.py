#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ from the REFERENCE itself.

Run in the build container only (needs /root/reference compiled into oracle/_ref by
`make -C oracle ref`).  The vectors are small and committed; the GPU box never needs the
reference sources.  Usage: python tests/golden/make_golden.py [what ...]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402


def golden_sw_vector():
    from swcases import make_vector_cases
    from test_oracle_sw_vector import SCORE_SETS
    for k, (name, (sc, colour)) in enumerate(sorted(SCORE_SETS.items())):
        seed, n = 1000 + k, 600
        cases = make_vector_cases(seed=seed, n_tasks=n, colour=colour)
        ref = oracle.RefSw(400, 200, sc, colour)
        scores = np.array([
            ref.sw_vector(cases["genome"], cases["goff"][t], cases["glen"][t], cases["reads"][t], cases["rlen"][t],
                          cases["genome_ls"] if colour else None, cases["initbp"][t] if colour else -1)
            for t in range(n)], dtype=np.int32)
        np.savez_compressed(os.path.join(HERE, f"sw_vector_{name}.npz"), seed=seed, n_tasks=n, scores=scores)
        print("wrote sw_vector", name, scores[:8])


def golden_mapping():
    """Runs the reference gmapper (and its DEBUG_HIT_LIST_* build) on the small synthetic configs and
    stores the hot-path SAM fields and the post-pass1 hit lists."""
    import re
    import subprocess
    import tempfile
    from mapcases import MAP_CASES, LsCase
    from oracle import pipeline as op
    pat = re.compile(r"\(cn:(\d+),st:(\d+),gen_st:(\d+),g_off:(-?\d+),w_len:(\d+),scores:\(wg=(-?\d+),vc=(-?\d+),"
                     r"fl=(-?\d+),poster=[^)]*\),matches:(\d+),pair_min:-?\d+,pair_max:-?\d+,"
                     r"anchor:\(x=(-?\d+),y=(-?\d+),ln=(\d+),wd=(\d+)\)\)")
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    for name, spec in sorted(MAP_CASES.items()):
        case = LsCase(name)
        with tempfile.TemporaryDirectory() as d:
            case.write_fasta(d)
            sam = os.path.join(d, "out.sam")
            with open(sam, "w") as f:
                subprocess.run([os.path.join(ref_dir, case.binary), *spec["args"], "reads.fa", "genome.fa"], cwd=d,
                               stdout=f, stderr=subprocess.DEVNULL, check=True)
            dbg = subprocess.run([os.path.join(ref_dir, "dbgbin", case.binary), *spec["args"], "reads.fa", "genome.fa"],
                                 cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, check=True, text=True).stderr
            recs = op.parse_sam(sam)
            seq_qual = op.parse_sam_seq_qual(sam) if name.endswith("_mq") else None
        name2idx = {n: i for i, n in enumerate(case.read_names)}
        cn2idx = {n: i for i, n in enumerate(case.contig_names)}
        sam_int = np.array([[name2idx[r[0]], r[1], cn2idx[r[2]], r[3], r[5], r[6]] for r in recs], dtype=np.int64)
        cigars = np.array([r[4] for r in recs])
        stage, mode, ridx = [], None, -1
        for line in dbg.splitlines():
            if line.startswith("Dumping hit list after pass1 for read:["):
                mode, ridx = "p1", name2idx[line.split("[")[1].split("]")[0]]
                continue
            if line.startswith("Dumping") or line.startswith("SW full"):
                mode = None
                continue
            if mode == "p1":
                m = pat.match(line)
                if m:
                    v = list(map(int, m.groups()))
                    st, ax, ay = v[1], v[9], v[10]
                    if v[2] == 1:
                        # colour-space pass 1 leaves strand-1 hits reversed (reverse_hit, mapping.c:254-263):
                        # undo it so the fixture is in positive-strand terms (anchor_reverse is an involution)
                        st = 1 - st
                        ax = -ax + (v[4] - 1) - (v[11] - 1) - (v[12] - 1)
                        ay = -ay + (int(case.read_len[ridx]) - 1) - (v[11] - 1) + (v[12] - 1)
                    stage.append((ridx, st, v[0], v[3], v[4], v[5], v[6], v[8], ax, ay, v[11], v[12]))
        extra = {}
        if seq_qual is not None:
            extra = dict(seq=np.array([a for a, _ in seq_qual]), qual=np.array([b for _, b in seq_qual]))
        np.savez_compressed(os.path.join(HERE, f"map_{name}.npz"), sam=sam_int, cigars=cigars,
                            stage=np.array(stage, dtype=np.int64), **extra)
        print("wrote map", name, sam_int.shape, len(stage))


def golden_sw_full():
    """sw_full_ls / sw_full_cs results of the REFERENCE objects on seeded random cases (global mode)."""
    from fullcases import make_full_cases
    from shrimp_b200.api import CS_DEFAULT_SCORES, LS_DEFAULT_SCORES
    for colour in (False, True):
        sc = CS_DEFAULT_SCORES if colour else LS_DEFAULT_SCORES
        cases = make_full_cases(seed=500 + colour, n=400, colour=colour, rlen_range=(25, 60))
        ref = oracle.RefFull(400, 200, sc, 8, colour=colour)
        refv = oracle.RefSw(400, 200, sc, False)
        ints, db, qr = [], [], []
        for c in cases:
            if colour:
                thresh = int(c["rlen"] * sc.match * 0.4)
                r = ref.sw_full_cs(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], c["initbp"], thresh,
                                   c["revcmpl"], c["anchor"], 0, None)
            else:
                v = refv.sw_vector(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"])
                thresh = int(c["rlen"] * sc.match * 0.5)
                if v < thresh:
                    ints.append((0,) * 10); db.append(b""); qr.append(b"")
                    continue
                r = ref.sw_full_ls(c["genome"], c["goff"], c["glen"], c["read"], c["rlen"], thresh, v, c["revcmpl"],
                                   c["anchor"], 0)
            ints.append(r[:10]); db.append(r[10].split(b"\0")[0]); qr.append(r[11].split(b"\0")[0])
        np.savez_compressed(os.path.join(HERE, f"sw_full_{'cs' if colour else 'ls'}.npz"),
                            ints=np.array(ints, dtype=np.int64), db=np.array(db), qr=np.array(qr))
        print("wrote sw_full", "cs" if colour else "ls", len(ints))


def golden_pairs():
    """Runs the reference gmapper on the small paired configs (default and --no-mapping-qualities) and stores every
    mapped SAM record's hot-path fields: pair, mate, flag, contig, pos, mate contig, mate pos, isize, AS, NM + CIGAR."""
    import subprocess
    import tempfile
    from mapcases import PAIR_CASES, PairCase
    from oracle import pipeline as op
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    for name, spec in sorted(PAIR_CASES.items()):
        case = PairCase(name)
        with tempfile.TemporaryDirectory() as d:
            case.write_fasta(d)
            sam = os.path.join(d, "out.sam")
            with open(sam, "w") as f:
                subprocess.run([os.path.join(ref_dir, case.binary), "-1", "m1.fa", "-2", "m2.fa", *spec["args"],
                                "genome.fa"], cwd=d, stdout=f, stderr=subprocess.DEVNULL, check=True)
            recs = op.parse_pair_sam(sam, {n: i for i, n in enumerate(case.contig_names)})
        ints = np.array([[r[0], r[1], r[2], r[3], r[4], r[6], r[7], r[8], r[9], r[10]] for r in recs], dtype=np.int64)
        cig = np.array([r[5] for r in recs])
        np.savez_compressed(os.path.join(HERE, f"pairs_{name}.npz"), recs=ints, cigars=cig)
        print("wrote pairs", name, ints.shape)


TARGETS = {"pairs": golden_pairs, "sw_vector": golden_sw_vector, "mapping": golden_mapping, "sw_full": golden_sw_full}

if __name__ == "__main__":
    assert oracle.have_ref(), "build oracle/_ref first: make -C oracle ref"
    for w in (sys.argv[1:] or sorted(TARGETS)):
        TARGETS[w]()
