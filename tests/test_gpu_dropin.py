"""The linked drop-in (integration/): the reference's UNCHANGED gmapper.c / genome.c / output.c / fasta.c objects
linked with the shims + libshrimp_b200.so, run FASTA -> SAM, and diffed byte for byte (minus the @PG line, which
carries the command line) against the reference binary on the same files and options.  This pins everything the
SAM body shows: MAPQ, Z0..Z6, mate fields, TLEN, SEQ/QUAL, XX:Z, CM:i, AS, NM.

Both binaries are prebuilt here (oracle/_ref, integration/_build) and travel to the GPU box."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mapcases import MAP_CASES, PAIR_CASES, LsCase, PairCase  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
NEW = os.path.join(ROOT, "integration", "_build")

needs_bins = pytest.mark.skipif(
    not (os.path.exists(os.path.join(REF, "gmapper-ls")) and os.path.exists(os.path.join(NEW, "gmapper-ls"))),
    reason="prebuilt reference / drop-in binaries not present")


def run_sam(bindir, binary, args, cwd, threads, extra=()):
    cmd = [os.path.join(bindir, binary), "-N", str(threads), *extra, *args]
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=1200)
    assert r.returncode == 0, (cmd, r.stderr.decode(errors="replace")[-3000:])
    body = [ln for ln in r.stdout.split(b"\n") if not ln.startswith(b"@PG")]
    return body, r.stderr.decode(errors="replace")


def assert_same_sam(ref, new):
    assert len(ref) == len(new), (len(ref), len(new))
    bad = [i for i, (a, b) in enumerate(zip(ref, new)) if a != b]
    assert not bad, (len(bad), [(ref[i], new[i]) for i in bad[:3]])


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MAP_CASES))
def test_unpaired_sam_identical(name, tmp_path):
    case = LsCase(name)
    case.write_fasta(str(tmp_path))
    args = [*MAP_CASES[name]["args"], "reads.fa", "genome.fa"]
    ref, _ = run_sam(REF, case.binary, args, str(tmp_path), 4)
    # -K 700: several chunks per thread, so the look-ahead batches start in the middle of the file too
    new, err = run_sam(NEW, case.binary, args, str(tmp_path), 2, ["-K", "700"])
    assert_same_sam(ref, new)
    assert sum(1 for ln in new if ln and not ln.startswith(b"@")) > 100


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(PAIR_CASES))
def test_paired_sam_identical(name, tmp_path):
    case = PairCase(name)
    case.write_fasta(str(tmp_path))
    args = [*PAIR_CASES[name]["args"], "-1", "m1.fa", "-2", "m2.fa", "genome.fa"]
    ref, _ = run_sam(REF, case.binary, args, str(tmp_path), 4)
    new, err = run_sam(NEW, case.binary, args, str(tmp_path), 2, ["-K", "300"])
    assert_same_sam(ref, new)
    assert sum(1 for ln in new if ln and not ln.startswith(b"@")) > 100


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("extra", [["--sam-unaligned"], ["--single-best-mapping"], ["--strata"], ["-o", "3"],
                                   ["--extra-sam-fields"], ["--all-contigs"], ["--trim-front", "3", "--trim-end", "2"],
                                   ["-U", "--local"], ["--local"], ["-n", "1"], ["-h", "40%", "-o", "20", "--strata"]])
def test_output_options_c1(extra, tmp_path):
    """options that only touch the unchanged output code or the loop of gmapper.c the look-ahead has to predict"""
    case = LsCase("c1_small")
    case.write_fasta(str(tmp_path))
    args = [*extra, "reads.fa", "genome.fa"]
    ref, _ = run_sam(REF, case.binary, args, str(tmp_path), 4)
    new, _ = run_sam(NEW, case.binary, args, str(tmp_path), 2, ["-K", "700"])
    assert_same_sam(ref, new)


@needs_bins
@pytest.mark.gpu
def test_lookahead_with_dropped_reads_and_many_threads(tmp_path):
    """one read in seven is dropped by the loop of gmapper.c (too long / low average quality) while four threads take
    chunks of 100: another thread's drops cut a look-ahead batch short, and the next batch starts in the middle of the
    chunk -- its look-ahead must stop at the end of re_buffer[] (it ran past it until round 2: a crash in one run of
    five).  Five runs, all equal to the reference."""
    import numpy as np
    case = LsCase("c1_small")
    case.write_fasta(str(tmp_path))
    rng = np.random.default_rng(77)
    reads = _mixed_reads(case, rng, 900, 22, 400, 0.01, False)
    with open(os.path.join(str(tmp_path), "mixed.fq"), "wb") as f:
        for name, s, q in reads:
            f.write(b"@" + name.encode() + b"\n" + s + b"\n+\n" + q + b"\n")
    args = ["-Q", "--qv-offset", "33", "--longest-read", "380", "mixed.fq", "genome.fa"]
    ref, _ = run_sam(REF, case.binary, args, str(tmp_path), 4)
    for threads, chunk in ((4, "100"), (2, "250"), (4, "100"), (3, "64"), (2, "250")):
        new, _ = run_sam(NEW, case.binary, args, str(tmp_path), threads, ["-K", chunk])
        assert_same_sam(ref, new)


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_small", "c2_small_fastq_mq"])
def test_dropin_with_the_references_own_reader_and_formatter(name, tmp_path):
    """gmapper-b200-refio: the same link line without integration/fast_io.cpp -- the reference's fasta.o, util.o and
    output.o as they are (the strict reading of "input/output and SAM emission unchanged"); bench.py reports it next to
    the default binary"""
    refio = os.path.join(NEW, "refio")
    if not os.path.exists(os.path.join(refio, "gmapper-ls")):
        pytest.skip("integration/_build/refio not built")
    case = LsCase(name)
    case.write_fasta(str(tmp_path))
    args = [*MAP_CASES[name]["args"], "reads.fa", "genome.fa"]
    ref, _ = run_sam(REF, case.binary, args, str(tmp_path), 4)
    new, _ = run_sam(refio, case.binary, args, str(tmp_path), 2, ["-K", "700"])
    assert_same_sam(ref, new)


@needs_bins
@pytest.mark.gpu
def test_long_reads_and_contig_ends(tmp_path):
    """reads of 500 to 1,000 bases (--longest-read is 1,000 by default, gmapper-defaults.h:72; the full SW leaves the
    256-column ring classes for the global-scratch kernel) and reads cut from the first and last 60 bases of every
    contig, both strands, some of them hanging over the end by a few random bases"""
    import numpy as np
    case = LsCase("c1_small")
    case.write_fasta(str(tmp_path))
    rng = np.random.default_rng(99)
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    recs = []
    for i in range(120):   # long reads, 1 % substitutions, a few short indels
        cn = int(rng.integers(0, len(case.contigs)))
        g = case.contigs[cn][1]
        rl = int(rng.integers(500, 1001))
        pos = int(rng.integers(0, g.size - rl - 1))
        frag = g[pos:pos + rl].copy()
        sub = rng.random(rl) < 0.01
        frag[sub] = acgt[rng.integers(0, 4, size=int(sub.sum()))]
        if i % 3 == 0:
            cut = int(rng.integers(50, rl - 50))
            frag = np.concatenate([frag[:cut], frag[cut + int(rng.integers(1, 4)):]])
        if rng.random() < 0.5:
            frag = comp[frag][::-1].copy()
        recs.append((b"long%d" % i, bytes(frag)))
    for cn, (_, g) in enumerate(case.contigs):   # contig ends
        for k in range(12):
            rl = int(rng.integers(40, 61))
            over = int(rng.integers(0, 6))
            junk = acgt[rng.integers(0, 4, size=over)]
            frag = np.concatenate([junk, g[:rl - over]]) if k % 2 == 0 else np.concatenate([g[g.size - (rl - over):], junk])
            if k % 4 >= 2:
                frag = comp[frag][::-1].copy()
            recs.append((b"end%d_%d" % (cn, k), bytes(frag)))
    with open(os.path.join(str(tmp_path), "long.fa"), "wb") as f:
        for name, s_ in recs:
            f.write(b">" + name + b"\n" + s_ + b"\n")
    args = ["long.fa", "genome.fa"]
    ref, _ = run_sam(REF, case.binary, args, str(tmp_path), 4)
    new, _ = run_sam(NEW, case.binary, args, str(tmp_path), 2, ["-K", "60"])
    assert_same_sam(ref, new)
    assert sum(1 for ln in new if ln and not ln.startswith(b"@")) > 100


@needs_bins
@pytest.mark.gpu
def test_long_colour_space_reads(tmp_path):
    """colour-space reads of 300 to 900 colours with qualities (per-position crossover scores, post_sw over hundreds of
    columns: one warp of eight alignments per CTA, the global-scratch full SW for the wide bands)"""
    import numpy as np
    import gen_synth
    case = LsCase("c2_small")
    case.write_fasta(str(tmp_path))
    rng = np.random.default_rng(123)
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    with open(os.path.join(str(tmp_path), "long.fq"), "wb") as f:
        for i in range(90):
            cn = int(rng.integers(0, len(case.contigs)))
            g = case.contigs[cn][1]
            rl = int(rng.integers(300, 901))
            pos = int(rng.integers(0, g.size - rl - 1))
            frag = g[pos:pos + rl].copy()
            if rng.random() < 0.5:
                frag = comp[frag][::-1].copy()
            s_ = gen_synth.letters_to_colour_read(frag, rng, 0.01)
            q = bytes((33 + rng.integers(5, 41, size=rl)).astype(np.uint8))
            f.write(b"@lc%d\n" % i + s_ + b"\n+\n" + q + b"\n")
    args = ["-Q", "--qv-offset", "33", "long.fq", "genome.fa"]
    ref, _ = run_sam(REF, case.binary, args, str(tmp_path), 4)
    new, _ = run_sam(NEW, case.binary, args, str(tmp_path), 2, ["-K", "40"])
    assert_same_sam(ref, new)
    assert sum(1 for ln in new if ln and not ln.startswith(b"@")) > 50


@needs_bins
@pytest.mark.gpu
def test_long_mates(tmp_path):
    """pairs whose mates are 250 to 450 bases long (inserts of 700 to 1,100, -I 0,1200), some mates with a short
    indel, some pairs with one mate that maps nowhere (half-paired output)"""
    import numpy as np
    case = PairCase("c3_small")
    case.write_fasta(str(tmp_path))
    rng = np.random.default_rng(321)
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    with open(os.path.join(str(tmp_path), "l1.fa"), "wb") as f1, open(os.path.join(str(tmp_path), "l2.fa"), "wb") as f2:
        for i in range(150):
            cn = int(rng.integers(0, len(case.contigs)))
            g = case.contigs[cn][1]
            ins = int(rng.integers(700, 1101))
            l1, l2 = int(rng.integers(250, 451)), int(rng.integers(250, 451))
            pos = int(rng.integers(0, g.size - ins - 1))
            a = g[pos:pos + l1].copy()
            b = comp[g[pos + ins - l2:pos + ins]][::-1].copy()
            for m in (a, b):
                sub = rng.random(m.size) < 0.01
                m[sub] = acgt[rng.integers(0, 4, size=int(sub.sum()))]
            if i % 4 == 0:
                cut = int(rng.integers(40, l1 - 40))
                a = np.concatenate([a[:cut], a[cut + 2:]])
            if i % 9 == 0:
                b = acgt[rng.integers(0, 4, size=l2)]
            if rng.random() < 0.5:   # the pair read from the other strand
                a, b = b, a
            f1.write(b">lp%d/1\n" % i + bytes(a) + b"\n")
            f2.write(b">lp%d/2\n" % i + bytes(b) + b"\n")
    args = ["-p", "opp-in", "-I", "0,1200", "-1", "l1.fa", "-2", "l2.fa", "genome.fa"]
    ref, _ = run_sam(REF, case.binary, args, str(tmp_path), 4)
    new, _ = run_sam(NEW, case.binary, args, str(tmp_path), 2, ["-K", "60"])
    assert_same_sam(ref, new)
    assert sum(1 for ln in new if ln and not ln.startswith(b"@")) > 150


STAGE_LISTS = {
    # a strict first set that stops at one alignment of 95 % of the maximum score, then a one-seed-match set with
    # lower thresholds for the reads that are left (every set that finds alignments prints them, mapping.c:1824-1833)
    "two_sets": ["--unpaired-options", "0;1/1,1,1/1,0,2,55%/1,60%,90%,2,0,30/68%,0,0,10/1,95%",
                 "--unpaired-options", "0;0/1,1,0/1,0,1,55%/1,50%,90%,1,0,30/55%,0,0,10/0"],
    "three_sets": ["--unpaired-options", "0;1/1,1,1/1,0,2,60%/1,70%,90%,2,0,20/75%,1,0,5/2,90%",
                   "--unpaired-options", "0;1/1,1,1/1,0,2,55%/1,60%,80%,2,0,30/65%,0,0,10/1,80%",
                   "--unpaired-options", "0;0/1,1,0/1,0,1,50%/1,45%,90%,1,0,30/50%,0,0,10/0"],
    # one user-given set that is not the default one
    "one_custom_set": ["--unpaired-options", "0;0/1,1,0/1,0,1,50%/1,55%,80%,1,0,25/60%,1,0,4/0"],
}


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("which", sorted(STAGE_LISTS))
@pytest.mark.parametrize("name", ["c1_small", "c2_small_mq"])
def test_unpaired_option_lists(name, which, tmp_path):
    """--unpaired-options (gmapper.c:2199-2218, handle_read's loop over the option sets mapping.c:1790-1841)"""
    case = LsCase(name)
    case.write_fasta(str(tmp_path))
    args = [*STAGE_LISTS[which], "reads.fa", "genome.fa"]
    ref, _ = run_sam(REF, case.binary, args, str(tmp_path), 4)
    new, _ = run_sam(NEW, case.binary, args, str(tmp_path), 2, ["-K", "700"])
    assert_same_sam(ref, new)
    assert sum(1 for ln in new if ln and not ln.startswith(b"@")) > 100


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["opp-out", "col-fw", "col-bw"])
def test_pair_modes(mode, tmp_path):
    """the pair orientations that reverse a mate before mapping (pair_reverse, gmapper-defaults.h:184-191)"""
    case = PairCase("c3_small")
    case.write_fasta(str(tmp_path))
    args = ["-p", mode, "-I", "0,1000", "-1", "m1.fa", "-2", "m2.fa", "genome.fa"]
    ref, _ = run_sam(REF, case.binary, args, str(tmp_path), 4)
    new, _ = run_sam(NEW, case.binary, args, str(tmp_path), 2, ["-K", "300"])
    assert_same_sam(ref, new)


def test_dropin_exports_the_reference_symbols():
    """nm: the mangled names mapping.o / sw-vector.o / sw-gapless.o / sw-full-ls.o / sw-full-cs.o export in the
    reference are all defined by the shims"""
    objs = [os.path.join(NEW, "mapping_shim.o"), os.path.join(NEW, "sw_shims.o")]
    if not all(os.path.exists(o) for o in objs):
        pytest.skip("integration/_build not built")
    out = subprocess.run(["nm", "--defined-only", *objs], stdout=subprocess.PIPE, check=True).stdout.decode()
    have = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    want = ["_Z15sw_vector_setupiiiiiiiiib", "_Z9sw_vectorPjiiS_iS_ib", "_Z15sw_vector_statsPmS_Pd",
            "_Z17sw_vector_cleanupv", "_Z16sw_gapless_setupiib", "_Z10sw_gaplessPjiS_iiiS_ib",
            "_Z16sw_gapless_statsPmS_S_", "_Z16sw_full_ls_setupiiiiiiiibi",
            "_Z10sw_full_lsPjiiS_iiiP15sw_full_resultsbP6anchorii", "_Z16sw_full_ls_statsPmS_Pd",
            "_Z18sw_full_ls_cleanupv", "_Z16sw_full_cs_setupiiiiiiiiibii",
            "_Z10sw_full_csPjiiS_iiiP15sw_full_resultsbbP6anchoriiPi", "_Z16sw_full_cs_statsPmS_Pd",
            "_Z18sw_full_cs_cleanupv", "_Z11handle_readP10read_entryP22read_mapping_options_ti",
            "_Z15handle_readpairP10pair_entryP26readpair_mapping_options_ti", "_Z15get_insert_sizeP8read_hitS0_"]
    missing = [w for w in want if w not in have]
    assert not missing, missing
    # and they are the reference's names: the reference objects define the very same symbols
    ref_objs = [os.path.join(REF, "obj", f"{n}.o") for n in ("gmapper_mapping", "common_sw-vector", "common_sw-gapless",
                                                             "common_sw-full-ls", "common_sw-full-cs")]
    if all(os.path.exists(o) for o in ref_objs):
        rout = subprocess.run(["nm", "--defined-only", *ref_objs], stdout=subprocess.PIPE, check=True).stdout.decode()
        rhave = {ln.split()[-1] for ln in rout.splitlines() if " T " in ln}
        assert set(want) <= rhave, sorted(set(want) - rhave)
        assert rhave <= have, sorted(rhave - have)


# ---- bench-size genomes (BASELINE.json configs) through the drop-in -------------------------------------------------
def _bench_case(key, n_reads, tmp_path, extra=(), genome_mb=300):
    """reads.fa of a bench.py workload + the projection held in HBM saved in the -S format (byte-identical to
    `gmapper -S`, tests/test_gpu_index.py), loaded by both binaries with -L so that the reference's serial index build
    stays out of the test"""
    sys.path.insert(0, ROOT)
    import copy

    import bench
    w = copy.copy(bench.WORKLOADS[key])
    w._genome = None
    if key == "c3":
        w.resize(genome_mb)
    codes, _ = w.reads(n_reads, 31)
    ctx = bench.build_context(w, 0)[0]
    try:
        bench.reference_setup(w, str(tmp_path), codes, ctx)
    finally:
        ctx.close()
    rd = ["-1", "reads.fa.1", "-2", "reads.fa.2"] if w.paired else ["reads.fa"]
    args = [*w.load_args(), *extra, "-L", "proj", *rd]
    ref, _ = run_sam(REF, w.binary, args, str(tmp_path), os.cpu_count() or 4)
    new, _ = run_sam(NEW, w.binary, args, str(tmp_path), 3, ["-K", str(max(1000, (n_reads // 5) & ~1))])
    assert_same_sam(ref, new)
    return sum(1 for ln in new if ln and not ln.startswith(b"@"))


@needs_bins
@pytest.mark.gpu
def test_c1_full_size(tmp_path):
    """BASELINE.json configs[0] at its full size: 100 k reads of 50 bp against the 10 Mb genome"""
    assert _bench_case("c1", 100_000, tmp_path) > 90_000


@needs_bins
@pytest.mark.gpu
def test_c4_full_size_default_and_mirna(tmp_path):
    """configs[3]: 22 bp reads against 2,000 contigs of 22 bp, default options and -M mirna"""
    d1, d2 = tmp_path / "a", tmp_path / "b"
    d1.mkdir()
    d2.mkdir()
    assert _bench_case("c4", 60_000, d1) > 50_000
    assert _bench_case("c4mirna", 60_000, d2) > 40_000


@needs_bins
@pytest.mark.gpu
def test_c5_sensitive_full_genome(tmp_path):
    """configs[4]: the overly sensitive option set against the 10 Mb genome (about a thousand windows per read)"""
    assert _bench_case("c5", 3_000, tmp_path) > 2_000


@needs_bins
@pytest.mark.gpu
def test_c3_pairs_300mb_sample(tmp_path):
    assert _bench_case("c3", 8_000, tmp_path) > 6_000


@needs_bins
@pytest.mark.gpu
def test_c3_pairs_hg18_size_prefix(tmp_path):
    """configs[2] at its full genome size: 5,000 pairs of 2 x 100 bp against the 3 Gb genome in 24 contigs (a 36 GB
    projection, saved from HBM and loaded by both binaries with -L): the drop-in's SAM equals the reference's.  Needs
    host memory for the reference's copy of the projection and disk for the files."""
    import shutil

    import psutil
    if psutil.virtual_memory().available < 120e9 or shutil.disk_usage(str(tmp_path)).free < 60e9:
        pytest.skip("needs 120 GB of host memory and 60 GB of disk")
    try:
        assert _bench_case("c3", 10_000, tmp_path, genome_mb=3000) > 9_000
    finally:
        for f in os.listdir(str(tmp_path)):
            if f.startswith("proj."):
                os.remove(os.path.join(str(tmp_path), f))


def _mixed_reads(case, rng, n, lo, hi, n_frac, colour):
    """reads of mixed lengths cut from the case's contigs, 2 % substitutions, some N; -> (names, strings, quals)"""
    import numpy as np
    out = []
    for i in range(n):
        cn = int(rng.integers(0, len(case.contigs)))
        g = case.contigs[cn][1]
        rl = int(rng.integers(lo, hi + 1))
        pos = int(rng.integers(0, g.size - rl - 1))
        frag = g[pos:pos + rl].copy()
        sub = rng.random(rl) < 0.02
        frag[sub] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(sub.sum()))]
        if rng.random() < 0.5:
            comp = np.zeros(256, dtype=np.uint8)
            for a, b in zip(b"ACGTN", b"TGCAN"):
                comp[a] = b
            frag = comp[frag][::-1].copy()
        frag[rng.random(rl) < n_frac] = ord("N")
        if colour:
            import gen_synth
            s = gen_synth.letters_to_colour_read(frag, rng, 0.02)
            q = bytes((33 + rng.integers(2, 41, size=rl)).astype(np.uint8))
        else:
            s = bytes(frag)
            # some reads of low average quality: the loop of gmapper.c drops them (min_avg_qv 10)
            top = 8 if rng.random() < 0.1 else 41
            q = bytes((33 + rng.integers(2, top, size=rl)).astype(np.uint8))
        out.append((f"m{i}", s, q))
    return out


@needs_bins
@pytest.mark.gpu
@pytest.mark.parametrize("colour", [False, True])
def test_mixed_lengths_fastq_and_n(colour, tmp_path):
    """one chunk holding reads of 22 to 400 bases (letter space; 25 to 120 colours in colour space), reads with N,
    FASTQ qualities (letter-space QUAL column, the average-quality filter of gmapper.c:496-527, per-position crossover
    scores in colour space) -- and a read that is too long for --longest-read"""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    case = LsCase("c2_small" if colour else "c1_small")
    case.write_fasta(str(tmp_path))
    rng = np.random.default_rng(77)
    reads = _mixed_reads(case, rng, 900, 25 if colour else 22, 120 if colour else 400, 0.01, colour)
    with open(os.path.join(str(tmp_path), "mixed.fq"), "wb") as f:
        for name, s, q in reads:
            f.write(b"@" + name.encode() + b"\n" + s + b"\n+\n" + q + b"\n")
    args = ["-Q", "--qv-offset", "33", "--longest-read", "380", "mixed.fq", "genome.fa"]
    ref, _ = run_sam(REF, case.binary, args, str(tmp_path), 4)
    new, _ = run_sam(NEW, case.binary, args, str(tmp_path), 2, ["-K", "250"])
    assert_same_sam(ref, new)
    assert sum(1 for ln in new if ln and not ln.startswith(b"@")) > 500
