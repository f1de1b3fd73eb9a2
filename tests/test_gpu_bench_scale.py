"""Bench-scale parity: the BASELINE.json workloads that bench.py times (C2 colour space vs 100 Mb; C3 pairs vs
300 Mb, the configuration that takes every read strand through the CTA-per-strand scan kernel, the heap-order
replay kernel and the slab levels) -- every mapped SAM record's hot-path fields equal the reference gmapper run on
this box's host cores.  The reference loads the projection that shrimp_gpu_projection_save wrote from HBM
(byte-identical to gmapper -S, tests/test_gpu_index.py), so its minutes-long serial index build stays out."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pipeline as op  # noqa: E402
from shrimp_b200 import align  # noqa: E402
from shrimp_b200.api import MapParams, auto_list_cutoff  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "gmapper-ls")),
                                 reason="prebuilt reference binaries (oracle/_ref) not present")]


def _reference_sam(bench, w, d, codes, ctx):
    bench.reference_setup(w, d, codes, ctx)
    rd = ["-1", "reads.fa.1", "-2", "reads.fa.2"] if w.paired else ["reads.fa"]
    sam = os.path.join(d, "ref.sam")
    with open(sam, "w") as f:
        subprocess.run([os.path.join(bench.REF_DIR, w.binary), "-N", str(os.cpu_count() or 1), *w.args, "-L", "proj", *rd],
                       cwd=d, stdout=f, stderr=subprocess.DEVNULL, check=True)
    return sam


def test_c2_sample_matches_reference_binary(tmp_path):
    import bench
    w = bench.WORKLOADS["c2nomq"]
    n = 60_000
    codes, initbp = w.reads(n, 77)
    ctx, scores, seeds, _ = bench.build_context(w, 0)
    try:
        ref = op.parse_sam(_reference_sam(bench, w, str(tmp_path), codes, ctx))
        params = MapParams(list_cutoff=auto_list_cutoff(w.genome_len, 12), compute_mapping_qualities=False)
        res = ctx.map_reads(params, scores, bench.pack_rows(codes), np.full(n, w.read_len, np.int32), initbp=initbp)
    finally:
        ctx.close()
    names = w.contig_names()
    assert len(res.hits) == len(ref) and len(ref) > n // 2
    bad = 0
    for h, rrec in zip(res.hits, ref):
        e = res.edits[int(h["edit_off"]): int(h["edit_off"]) + int(h["edit_len"])]
        f = align.sam_fields(h, e, w.read_len, w.contig_len, True)
        mine = (f"r{int(h['read_idx'])}", f[0], names[f[1]], f[2], f[3], f[4], f[5])
        if mine != rrec:
            bad += 1
            if bad < 5:
                print("DIFF", mine, rrec)
    assert bad == 0


def test_c2_mapping_qualities_sample_matches_reference_binary(tmp_path):
    """gmapper-cs with its default options: post_sw (device) rescoring -- AS, NM, position, CIGAR and the corrected
    base calls (SEQ) of every record equal the reference's"""
    import bench
    w = bench.WORKLOADS["c2"]
    n = 40_000
    codes, initbp = w.reads(n, 79)
    ctx, scores, seeds, _ = bench.build_context(w, 0)
    try:
        sam_path = _reference_sam(bench, w, str(tmp_path), codes, ctx)
        ref = op.parse_sam(sam_path)
        ref_sq = op.parse_sam_seq_qual(sam_path)
        params = MapParams(list_cutoff=auto_list_cutoff(w.genome_len, 12), compute_mapping_qualities=True)
        res = ctx.map_reads(params, scores, bench.pack_rows(codes), np.full(n, w.read_len, np.int32), initbp=initbp)
    finally:
        ctx.close()
    names = w.contig_names()
    assert len(res.hits) == len(ref) and len(ref) > n // 2
    bad = 0
    for h, rrec, (rseq, rqual) in zip(res.hits, ref, ref_sq):
        e0, el, rm = int(h["edit_off"]), int(h["edit_len"]), int(h["rmapped"])
        e = res.edits[e0:e0 + el]
        f = align.sam_fields(h, e, w.read_len, w.contig_len, True)
        seq, _ = align.post_sw_seq_qual(e, res.edits[e0 + el:e0 + el + rm], int(h["gen_st"]) == 1, False)
        mine = (f"r{int(h['read_idx'])}", f[0], names[f[1]], f[2], f[3], f[4], f[5])
        if mine != rrec or seq != rseq or rqual != "*":
            bad += 1
            if bad < 5:
                print("DIFF", mine, rrec, seq, rseq)
    assert bad == 0


def test_c3_sample_matches_reference_binary(tmp_path):
    import bench
    w = bench.WORKLOADS["c3"]
    if w.genome_len != 300_000_000:
        w.resize(300)
    n_pairs = 12_000
    codes, _ = w.reads(2 * n_pairs, 78)
    ctx, scores, seeds, _ = bench.build_context(w, 0)
    try:
        idx = {nm: i for i, nm in enumerate(w.contig_names())}
        ref = op.parse_pair_sam(_reference_sam(bench, w, str(tmp_path), codes, ctx), idx)
        params = MapParams(list_cutoff=auto_list_cutoff(w.genome_len, 12), match_mode=4)
        res = ctx.map_pairs(params, scores, bench.pack_rows(codes), np.full(2 * n_pairs, w.read_len, np.int32))
    finally:
        ctx.close()
    assert res.stats["scan_big_strands"] == 4 * n_pairs      # the dense-regime kernel served every strand
    assert res.stats["heap_replays"] > 0                     # ... and the replay kernel ran
    lens = [w.contig_len] * w.n_contigs
    read_len = np.full(2 * n_pairs, w.read_len, np.int32)

    def fields(h, rl, gl):
        e = res.edits[int(h["edit_off"]): int(h["edit_off"]) + int(h["edit_len"])]
        return align.sam_fields(h, e, rl, gl, False)

    pair_hits = [(res.hits[int(p["hit_idx"][0])], res.hits[int(p["hit_idx"][1])]) for p in res.pairs]
    mine = op.pair_sam_records(pair_hits, res.pairs["pair_idx"], res.hits[res.n_paired_hits:], lens, read_len, fields,
                               n_pairs)
    assert len(mine) == len(ref) and len(ref) > n_pairs
    bad = [k for k, (a, b) in enumerate(zip(mine, ref)) if tuple(a) != tuple(b)]
    assert not bad, (len(bad), [(mine[k], ref[k]) for k in bad[:3]])
