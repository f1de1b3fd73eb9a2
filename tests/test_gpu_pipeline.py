"""GPU parity of the chunk pipeline (shrimp_gpu_map_reads): post-pass1 hit lists and the reported
alignments vs (a) golden vectors from the reference gmapper and (b) the CPU oracle on the same inputs."""
import os

import numpy as np
import pytest

from mapcases import GOLD, MAP_CASES, LsCase, stage_tuple_array
from shrimp_b200 import align
from shrimp_b200.api import MapParams, _pack_codes, auto_list_cutoff

pytestmark = pytest.mark.gpu


def run_gpu(ctx, case, **over):
    ctx.sw_setup(1500, 1000, case.scores, use_colours=case.colour, anchor_width=case.anchor_width)
    ctx.load_genome([_pack_codes(c.astype(np.uint32)) for c in case.contig_codes], [c.size for c in case.contig_codes],
                    colour_space=case.colour)
    ctx.build_index(case.seeds, hflag=case.hflag)
    params = MapParams(list_cutoff=case.list_cutoff, **over)
    return ctx.map_reads(params, case.scores, case.packed, case.read_len, initbp=case.initbp, want_stage=True,
                         crossover_scores=case.crossover_scores,
                         quals=case.quals if params.compute_mapping_qualities else None)


def sam_arrays(case, res):
    rows, cig = [], []
    for h in res.hits:
        e = res.edits[int(h["edit_off"]): int(h["edit_off"]) + int(h["edit_len"])]
        f = align.sam_fields(h, e, int(case.read_len[h["read_idx"]]), int(case.contig_codes[h["cn"]].size), case.colour)
        rows.append([int(h["read_idx"]), f[0], f[1], f[2], f[4], f[5]])
        cig.append(f[3])
    return np.array(rows, dtype=np.int64).reshape(-1, 6), np.array(cig)


@pytest.mark.parametrize("name", sorted(MAP_CASES))
def test_pipeline_matches_reference_golden(gpu_ctx, name):
    gold = np.load(os.path.join(GOLD, f"map_{name}.npz"))
    case = LsCase(name)
    res = run_gpu(gpu_ctx, case, **MAP_CASES[name]["opts"])
    got_stage = stage_tuple_array(res.stage)
    assert got_stage.shape == gold["stage"].shape
    bad = np.nonzero((got_stage != gold["stage"]).any(axis=1))[0]
    assert bad.size == 0, (bad[:5], got_stage[bad[:5]], gold["stage"][bad[:5]])
    sam, cig = sam_arrays(case, res)
    assert sam.shape == gold["sam"].shape
    bad = np.nonzero((sam != gold["sam"]).any(axis=1))[0]
    assert bad.size == 0, (bad[:5], sam[bad[:5]], gold["sam"][bad[:5]])
    assert np.array_equal(cig, gold["cigars"])
    if "seq" in gold:   # colour space with mapping qualities: post_sw's corrected base calls and base qualities
        sq = []
        for h in res.hits:
            e0, el, rm = int(h["edit_off"]), int(h["edit_len"]), int(h["rmapped"])
            sq.append(align.post_sw_seq_qual(res.edits[e0:e0 + el], res.edits[e0 + el:e0 + el + rm],
                                             int(h["gen_st"]) == 1, case.quals is not None))
        assert [a for a, _ in sq] == gold["seq"].tolist()
        assert [b for _, b in sq] == gold["qual"].tolist()


@pytest.mark.parametrize("name,env", [("c1_small", {"SHRIMP_BUCKET_HEADS": "0"}), ("c2_small", {"SHRIMP_BUCKET_HEADS": "0"}),
                                      ("c1_repeat", {"SHRIMP_BUCKET_HEADS": "0"}),
                                      ("c2_small_mq", {"SHRIMP_POST_SW_HALF": "1"}),
                                      ("c2_small_fastq_mq", {"SHRIMP_POST_SW_HALF": "1"})])
def test_alternative_paths_match_reference_golden(gpu_ctx, name, env, monkeypatch):
    """the paths the defaults no longer take on these small cases: k-mer lookups through the CSR bounds instead of
    the bucket heads of a sparse projection (genome.cuh), and the half-warp post_sw kernel instead of the quad one"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    test_pipeline_matches_reference_golden(gpu_ctx, name)


def test_repeats_reach_the_cta_scan_kernel(gpu_ctx):
    """read strands with more candidates than a warp's slab go through scan_big_kernel (and still match, see the
    golden test of c1_repeat)"""
    case = LsCase("c1_repeat")
    res = run_gpu(gpu_ctx, case)
    assert res.stats["scan_big_strands"] > 0


@pytest.mark.parametrize("name", ["c1_small", "c2_small", "c5_small"])
def test_cta_scan_kernel_alone_matches_reference_golden(gpu_ctx, name, monkeypatch):
    """every read strand forced through the CTA-per-strand scan kernel (the path long index lists take at
    hg18 scale): same hit lists and alignments as the reference"""
    monkeypatch.setenv("SHRIMP_SCAN_FORCE_BIG", "1")
    gold = np.load(os.path.join(GOLD, f"map_{name}.npz"))
    case = LsCase(name)
    res = run_gpu(gpu_ctx, case, **MAP_CASES[name]["opts"])
    assert res.stats["scan_big_strands"] == 2 * len(case.read_len)
    got_stage = stage_tuple_array(res.stage)
    assert got_stage.shape == gold["stage"].shape and np.array_equal(got_stage, gold["stage"])
    sam, cig = sam_arrays(case, res)
    assert np.array_equal(sam, gold["sam"]) and np.array_equal(cig, gold["cigars"])


@pytest.mark.parametrize("name,env", [
    ("c1_small", {"SHRIMP_SCAN_BM_LOG2": "5"}),                                   # 4 partitions of 32 regions
    ("c1_repeat", {"SHRIMP_SCAN_BM_LOG2": "6"}),                                  # 3 partitions of 64 regions
    ("c2_small", {"SHRIMP_SCAN_BM_LOG2": "5", "SHRIMP_SCAN_CTA_THREADS": "64"}),
    ("c1_repeat", {"SHRIMP_SCAN_CTA_CAP": "32"}),                                  # slab of 32: global-slab pass
    ("c1_repeat", {"SHRIMP_SCAN_CTA_CAP": "32", "SHRIMP_SCAN_BM_LOG2": "5"}),
    ("c5_small", {"SHRIMP_SCAN_CTA_CAP": "64"}),                                   # no region filter
    # the staged-window passes (the mate-pair modes still run them): several staging windows
    ("c1_repeat", {"SHRIMP_SCAN_WALK": "0", "SHRIMP_SCAN_WIN": "64", "SHRIMP_SCAN_BM_LOG2": "6"}),
    ("c5_small", {"SHRIMP_SCAN_WALK": "0", "SHRIMP_SCAN_WIN": "128", "SHRIMP_SCAN_LANES_LOG2": "5"}),
    ("c2_small", {"SHRIMP_SCAN_WALK": "0", "SHRIMP_SCAN_WIN": "96", "SHRIMP_SCAN_LANES_LOG2": "0"}),
    ("c1_repeat", {"SHRIMP_SCAN_HASHED": "1", "SHRIMP_SCAN_BM_LOG2": "7"}),       # hashed bitmaps of 128 bits
    ("c2_small", {"SHRIMP_SCAN_WALK": "0", "SHRIMP_SCAN_HASHED": "1", "SHRIMP_SCAN_WIN": "64"}),
    # the cursor walk with one lane and with a whole warp per list, tiles of 32 regions
    ("c1_repeat", {"SHRIMP_SCAN_LANES_LOG2": "0", "SHRIMP_SCAN_BM_LOG2": "5"}),
    ("c5_small", {"SHRIMP_SCAN_LANES_LOG2": "5"}),
    ("c2_small", {"SHRIMP_SCAN_LANES_LOG2": "4", "SHRIMP_SCAN_BM_LOG2": "6"}),
])
def test_cta_scan_partitions_and_global_slabs(gpu_ctx, name, env, monkeypatch):
    """the CTA scan kernel with the genome cut into several bitmap partitions (how a 3 Gb genome runs), and with
    a candidate slab so small that strands fall through to the global-slab launch: same results"""
    monkeypatch.setenv("SHRIMP_SCAN_FORCE_BIG", "1")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    gold = np.load(os.path.join(GOLD, f"map_{name}.npz"))
    case = LsCase(name)
    res = run_gpu(gpu_ctx, case, **MAP_CASES[name]["opts"])
    assert res.stats["scan_big_strands"] == 2 * len(case.read_len)
    if "SHRIMP_SCAN_CTA_CAP" in env:
        assert res.stats["scan_global_strands"] > 0
    got_stage = stage_tuple_array(res.stage)
    assert got_stage.shape == gold["stage"].shape and np.array_equal(got_stage, gold["stage"])
    sam, cig = sam_arrays(case, res)
    assert np.array_equal(sam, gold["sam"]) and np.array_equal(cig, gold["cigars"])


def test_pipeline_alignment_strings_match_oracle(gpu_ctx):
    """dbalign/qralign rebuilt from the edit script equal the oracle's pretty_print strings"""
    from oracle import pipeline as op
    from test_oracle_pipeline import run_oracle
    case = LsCase("c1_small")
    res = run_gpu(gpu_ctx, case)
    g, hits, nper, stage, stats = run_oracle(case)
    assert len(hits) == len(res.hits)
    cmpl = np.array([3, 2, 1, 0, 0, 10, 9, 7, 8, 6, 5, 14, 13, 12, 11, 15], dtype=np.uint8)
    for a, b in zip(res.hits[:400], hits[:400]):
        codes = case.contig_codes[int(a["cn"])]
        if int(a["gen_st"]) == 1:
            codes = cmpl[codes[::-1]]
        e = res.edits[int(a["edit_off"]): int(a["edit_off"]) + int(a["edit_len"])]
        rcodes = np.frombuffer(case.reads[int(a["read_idx"])][1].tobytes(), dtype=np.uint8)
        from shrimp_b200.api import _LS_CODE
        db, qr = align.align_strings(e, codes, int(a["genome_start"]), _LS_CODE[rcodes], int(a["read_start"]))
        assert db == bytes(b["sfr"]["dbalign"]).split(b"\0")[0]
        assert qr == bytes(b["sfr"]["qralign"]).split(b"\0")[0]
        for k in ("read_start", "rmapped", "genome_start", "gmapped", "mismatches", "insertions", "deletions"):
            assert int(a[k]) == int(b["sfr"][k])
        assert int(a["sfr_matches"]) == int(b["sfr"]["matches"])
        assert int(a["sw_score"]) == int(b["sw_score"]) and int(a["score_full"]) == int(b["score_full"])
    assert res.stats["vector_calls"] == stats["vector_calls"]
    assert res.stats["vector_cells"] == stats["vector_cells"]
    assert res.stats["full_cells"] == stats["full_cells"]


def test_second_context_shares_the_index(gpu_ctx):
    """one context per host thread, shared genome + projection (gmapper's -N threads): same results from both,
    also when they map concurrently"""
    import threading
    import shrimp_b200
    case = LsCase("c1_small")
    want = run_gpu(gpu_ctx, case)
    with shrimp_b200.GpuContext(0) as ctx2:
        ctx2.sw_setup(1500, 1000, case.scores, anchor_width=case.anchor_width)
        ctx2.share_genome_from(gpu_ctx)
        params = MapParams(list_cutoff=case.list_cutoff)
        out = [None, None]

        def go(i, c):
            out[i] = c.map_reads(params, case.scores, case.packed, case.read_len)

        ths = [threading.Thread(target=go, args=(0, gpu_ctx)), threading.Thread(target=go, args=(1, ctx2))]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    for res in out:
        assert np.array_equal(res.n_hits_per_read, want.n_hits_per_read)
        for k in ("cn", "gen_st", "genome_start", "score_full", "edit_len"):
            assert np.array_equal(res.hits[k], want.hits[k])
        assert np.array_equal(res.edits, want.edits)


def test_pipeline_empty_and_tiny_reads(gpu_ctx):
    case = LsCase("c1_small")
    gpu_ctx.sw_setup(1400, 1000, case.scores)
    gpu_ctx.load_genome([_pack_codes(c.astype(np.uint32)) for c in case.contig_codes], [c.size for c in case.contig_codes])
    gpu_ctx.build_index(case.seeds)
    params = MapParams(list_cutoff=1000)
    res = gpu_ctx.map_reads(params, case.scores, case.packed[:0], case.read_len[:0])
    assert len(res.hits) == 0
    # reads shorter than every seed map nowhere (gmapper.c:503-507 skips them)
    rl = case.read_len[:8].copy()
    rl[:] = 10
    res = gpu_ctx.map_reads(params, case.scores, case.packed[:8], rl)
    assert len(res.hits) == 0 and (res.n_hits_per_read == 0).all()


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "gmapper-ls")),
                    reason="prebuilt reference binary (oracle/_ref) not present")
def test_c1_full_size_matches_reference_binary(gpu_ctx, tmp_path):
    """BASELINE.json configs[0] at full size (100 k x 50 bp vs 10 Mb): every SAM record's hot-path fields
    equal the reference gmapper run on this box's CPUs (oracle/_ref/gmapper-ls -N <cores>).  This size has
    the rare equal-position anchors that need the exact heap replay (SURVEY hard part 3a)."""
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    import bench
    from oracle import pipeline as op
    from shrimp_b200 import seeds as S
    from shrimp_b200.api import LS_DEFAULT_SCORES
    n = 100_000
    genome, reads, _, _ = bench.make_workload(n)
    d = str(tmp_path)
    bench.write_fasta(os.path.join(d, "genome.fa"), ["contig0"], [genome])
    bench.write_fasta(os.path.join(d, "reads.fa"), [f"r{i}" for i in range(n)], reads)
    with open(os.path.join(d, "ref.sam"), "w") as f:
        subprocess.run([os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "gmapper-ls"), "-N",
                        str(os.cpu_count() or 1), "reads.fa", "genome.fa"], cwd=d, stdout=f,
                       stderr=subprocess.DEVNULL, check=True)
    ref = op.parse_sam(os.path.join(d, "ref.sam"))
    gpu_ctx.sw_setup(1400, 1000, LS_DEFAULT_SCORES)
    gpu_ctx.load_genome([_pack_codes(genome.astype(np.uint32))], [genome.size])
    gpu_ctx.build_index(S.load_default_seeds())
    params = MapParams(list_cutoff=auto_list_cutoff(genome.size, 12))
    res = gpu_ctx.map_reads(params, LS_DEFAULT_SCORES, bench.pack_rows(reads), np.full(n, 50, np.int32))
    assert res.stats["heap_replays"] > 0
    assert len(res.hits) == len(ref)
    bad = 0
    for h, rrec in zip(res.hits, ref):
        e = res.edits[int(h["edit_off"]): int(h["edit_off"]) + int(h["edit_len"])]
        f = align.sam_fields(h, e, 50, genome.size)
        mine = (f"r{int(h['read_idx'])}", f[0], "contig0", f[2], f[3], f[4], f[5])
        if mine != rrec:
            bad += 1
            if bad < 5:
                print("DIFF", mine, rrec)
    assert bad == 0
