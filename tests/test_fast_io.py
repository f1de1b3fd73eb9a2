"""SURVEY section 8 f2 on the CPU: integration/fast_io.cpp (FASTA/FASTQ reader + SAM record formatter) linked into the
REFERENCE -- `integration/_build/gmapper-ref-fastio` is the reference's own objects, mapping.o and the SW kernels
included, with only fasta_get_next_read_with_range and hit_output replaced -- against the plain reference binary
(oracle/_ref): the complete SAM text, @PG line aside, must be byte-identical.  No GPU involved, so every output
option, FASTQ flavour and malformed-input message can be exercised here; the drop-in proper (the same two functions
over the device path) is diffed on the GPU box by tests/test_gpu_dropin.py."""
import gzip
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mapcases import MAP_CASES, PAIR_CASES, LsCase, PairCase  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
NEW = os.path.join(ROOT, "integration", "_build", "fastio")

pytestmark = pytest.mark.skipif(
    not (os.path.exists(os.path.join(REF, "gmapper-ls")) and os.path.exists(os.path.join(NEW, "gmapper-ls"))),
    reason="prebuilt reference / gmapper-ref-fastio binaries not present")


def run(bindir, binary, args, cwd, threads=4, stdin=None, ok=(0,)):
    cmd = [os.path.join(bindir, binary), "-N", str(threads), *args]
    # the optional read-ahead thread of fast_io.cpp is switched on, and starts after 50 entries instead of 4,096 so
    # that these small files use it (test_without_read_ahead covers the default)
    env = dict(os.environ, SHRIMP_B200_READ_AHEAD="1",
               SHRIMP_B200_READ_AHEAD_AFTER=os.environ.get("SHRIMP_B200_READ_AHEAD_AFTER", "50"))
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=1200, stdin=stdin, env=env)
    assert r.returncode in ok, (cmd, r.returncode, r.stderr.decode(errors="replace")[-2000:])
    body = [ln for ln in r.stdout.split(b"\n") if not ln.startswith(b"@PG")]
    return body, r.stderr.decode(errors="replace"), r.returncode


def same(ref, new):
    assert len(ref) == len(new), (len(ref), len(new))
    bad = [i for i, (a, b) in enumerate(zip(ref, new)) if a != b]
    assert not bad, (len(bad), [(ref[i], new[i]) for i in bad[:3]])


def subset(case, n):
    case.reads = case.reads[:n]
    if getattr(case, "quals", None) is not None:
        case.quals = case.quals[:n]
    return case


@pytest.mark.parametrize("name", ["c1_small", "c2_small_mq", "c2_small_fastq_mq", "c2_small_fastq_local", "c4_small_mirna"])
def test_unpaired(name, tmp_path):
    case = subset(LsCase(name), 1500)
    case.write_fasta(str(tmp_path))
    args = [*MAP_CASES[name]["args"], "reads.fa", "genome.fa"]
    ref, _, _ = run(REF, case.binary, args, str(tmp_path))
    new, _, _ = run(NEW, case.binary, args, str(tmp_path), 3, )
    same(ref, new)
    assert sum(1 for ln in new if ln and not ln.startswith(b"@")) > 100


@pytest.mark.parametrize("name", ["c3_small", "c3_small_nomq", "c2p_small_mq", "c3_small_nohp"])
def test_paired(name, tmp_path):
    case = PairCase(name)
    case.write_fasta(str(tmp_path))
    # the first 600 pairs
    for f in ("m1.fa", "m2.fa"):
        lines = open(os.path.join(str(tmp_path), f), "rb").read().split(b"\n")
        open(os.path.join(str(tmp_path), f), "wb").write(b"\n".join(lines[:1200]) + b"\n")
    args = [*PAIR_CASES[name]["args"], "-1", "m1.fa", "-2", "m2.fa", "genome.fa"]
    ref, _, _ = run(REF, case.binary, args, str(tmp_path))
    new, _, _ = run(NEW, case.binary, args, str(tmp_path), 3, )
    same(ref, new)
    assert sum(1 for ln in new if ln and not ln.startswith(b"@")) > 100


OUTPUT_OPTIONS = [["--sam-unaligned"], ["--single-best-mapping"], ["--strata"], ["-o", "3"], ["--extra-sam-fields"],
                  ["--all-contigs"], ["--trim-front", "3", "--trim-end", "2"], ["--read-group", "grp1,sampleA"],
                  ["-P"], ["-R"], ["--sam-unaligned", "--no-mapping-qualities"]]


@pytest.mark.parametrize("extra", OUTPUT_OPTIONS, ids=lambda e: "_".join(e).replace("-", ""))
def test_output_options_letter_space(extra, tmp_path):
    case = subset(LsCase("c1_small"), 1200)
    case.write_fasta(str(tmp_path))
    # some reads that map nowhere, for --sam-unaligned
    rng = np.random.default_rng(5)
    with open(os.path.join(str(tmp_path), "reads.fa"), "ab") as f:
        for i in range(40):
            f.write(b">junk%d\n" % i + bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=50)) + b"\n")
    args = [*extra, "reads.fa", "genome.fa"]
    if extra in (["-P"], ["-R"]):
        args = ["--shrimp-format", *args]   # the old output format: goes to the reference's own hit_output
    ref, _, _ = run(REF, case.binary, args, str(tmp_path))
    new, _, _ = run(NEW, case.binary, args, str(tmp_path), 3)
    same(ref, new)


@pytest.mark.parametrize("extra", [["--sam-unaligned"], ["--bfast"], ["--ignore-qvs"],
                                   ["--no-mapping-qualities", "--sam-unaligned"]],
                         ids=lambda e: "_".join(e).replace("-", ""))
def test_output_options_colour_fastq(extra, tmp_path):
    case = subset(LsCase("c2_small_fastq_mq"), 800)
    case.write_fasta(str(tmp_path))
    args = ["-Q", *extra, "reads.fa", "genome.fa"]
    ref, _, _ = run(REF, case.binary, args, str(tmp_path))
    new, _, _ = run(NEW, case.binary, args, str(tmp_path), 3)
    same(ref, new)


@pytest.mark.parametrize("extra", [["--sam-unaligned", "--sam-r2"], ["--no-half-paired", "--sam-unaligned"],
                                   ["--single-best-mapping"], ["--no-improper-mappings"], ["--extra-sam-fields"]],
                         ids=lambda e: "_".join(e).replace("-", ""))
def test_output_options_pairs(extra, tmp_path):
    case = PairCase("c3_small")
    case.write_fasta(str(tmp_path))
    for f in ("m1.fa", "m2.fa"):
        lines = open(os.path.join(str(tmp_path), f), "rb").read().split(b"\n")
        open(os.path.join(str(tmp_path), f), "wb").write(b"\n".join(lines[:800]) + b"\n")
    args = ["-p", "opp-in", "-I", "0,1000", *extra, "-1", "m1.fa", "-2", "m2.fa", "genome.fa"]
    ref, _, _ = run(REF, case.binary, args, str(tmp_path))
    new, _, _ = run(NEW, case.binary, args, str(tmp_path), 3)
    same(ref, new)


def _letter_fastq(case, rng, n, lower=False, iupac=False, offset=33):
    recs = []
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    for i in range(n):
        cn = int(rng.integers(0, len(case.contigs)))
        g = case.contigs[cn][1]
        rl = int(rng.integers(30, 120))
        pos = int(rng.integers(0, g.size - rl - 1))
        frag = g[pos:pos + rl].copy()
        if rng.random() < 0.5:
            frag = comp[frag][::-1].copy()
        if iupac and rng.random() < 0.3:
            frag[int(rng.integers(0, rl))] = int(rng.choice(np.frombuffer(b"RYSWKMBDHVN", dtype=np.uint8)))
        s = bytes(frag)
        if lower and rng.random() < 0.3:
            s = s.lower()
        q = bytes((offset + rng.integers(2, 41, size=rl)).astype(np.uint8))
        recs.append((b"q%d" % i, s, q))
    return recs


def test_letter_space_fastq_flavours(tmp_path):
    """FASTQ in letter space: QUAL column forward and reversed, lower-case and IUPAC letters in reads (SEQ rule of
    output.c:314-336, :503-531), PHRED+64 input (--qv-offset 64: QUAL is rebased to 33), names with blanks and
    tab-separated comments, a name pair that shares a prefix ending in '/', Windows-free plain lines, a gzip copy"""
    case = LsCase("c1_small")
    case.write_fasta(str(tmp_path))
    rng = np.random.default_rng(11)
    recs = _letter_fastq(case, rng, 700, lower=True, iupac=True)
    with open(os.path.join(str(tmp_path), "r.fq"), "wb") as f:
        f.write(b"# a header comment\n#another\n")
        for k, (name, s, q) in enumerate(recs):
            nm = name + (b" extra words" if k % 3 == 0 else b"") + (b"\tcomment field" if k % 5 == 0 else b"")
            f.write(b"@" + nm + b"\n" + s + b"\n+" + (name if k % 2 else b"") + b"\n" + q + b"\n")
    args = ["-Q", "--qv-offset", "33", "--sam-unaligned", "r.fq", "genome.fa"]
    ref, _, _ = run(REF, "gmapper-ls", args, str(tmp_path))
    new, _, _ = run(NEW, "gmapper-ls", args, str(tmp_path), 3)
    same(ref, new)
    assert sum(1 for ln in new if ln and not ln.startswith(b"@")) > 400
    # the same file gzip-compressed, and through stdin
    with open(os.path.join(str(tmp_path), "r.fq"), "rb") as f, gzip.open(os.path.join(str(tmp_path), "r.fq.gz"), "wb") as g:
        g.write(f.read())
    new_gz, _, _ = run(NEW, "gmapper-ls", ["-Q", "--qv-offset", "33", "--sam-unaligned", "r.fq.gz", "genome.fa"],
                       str(tmp_path), 3)
    same(ref, new_gz)
    with open(os.path.join(str(tmp_path), "r.fq"), "rb") as f:
        new_in, _, _ = run(NEW, "gmapper-ls", ["-Q", "--qv-offset", "33", "--sam-unaligned", "-", "genome.fa"],
                           str(tmp_path), 3, stdin=f)
    same(ref, new_in)
    # PHRED+64
    recs64 = _letter_fastq(case, rng, 300, offset=64)
    with open(os.path.join(str(tmp_path), "r64.fq"), "wb") as f:
        for name, s, q in recs64:
            f.write(b"@" + name + b"\n" + s + b"\n+\n" + q + b"\n")
    args = ["-Q", "--qv-offset", "64", "r64.fq", "genome.fa"]
    ref, _, _ = run(REF, "gmapper-ls", args, str(tmp_path))
    new, _, _ = run(NEW, "gmapper-ls", args, str(tmp_path), 3)
    same(ref, new)


def test_multi_line_fasta_and_fastq(tmp_path):
    """sequences and qualities folded over several lines, comment lines between entries, no newline at the end of
    the file, and a genome whose contigs are folded at 60 columns (the genome goes through the same reader)"""
    case = LsCase("c1_small")
    case.write_fasta(str(tmp_path))
    rng = np.random.default_rng(12)
    recs = _letter_fastq(case, rng, 300)

    def fold(b, w):
        return b"\n".join(b[i:i + w] for i in range(0, len(b), w))
    with open(os.path.join(str(tmp_path), "f.fa"), "wb") as f:
        for k, (name, s, _) in enumerate(recs):
            if k % 7 == 0:
                f.write(b"#comment between entries\n")
            f.write(b">" + name + b"\n" + fold(s, 17) + (b"\n" if k + 1 < len(recs) else b""))
    with open(os.path.join(str(tmp_path), "f.fq"), "wb") as f:
        for name, s, q in recs:
            f.write(b"@" + name + b"\n" + fold(s, 23) + b"\n+\n" + fold(q, 23) + b"\n")
    for args in (["f.fa", "genome.fa"], ["-Q", "--qv-offset", "33", "f.fq", "genome.fa"]):
        ref, _, _ = run(REF, "gmapper-ls", args, str(tmp_path))
        new, _, _ = run(NEW, "gmapper-ls", args, str(tmp_path), 3)
        same(ref, new)
        assert sum(1 for ln in new if ln and not ln.startswith(b"@")) > 200


@pytest.mark.parametrize("kind", ["fastq_as_fasta", "short_qual", "empty_seq", "truncated", "blank_tail"])
def test_malformed_input_same_outcome(kind, tmp_path):
    """what the reference does with broken input -- stop reading at the bad entry with a message, or exit -- happens
    here too: same SAM up to that point, same exit status"""
    case = LsCase("c1_small")
    case.write_fasta(str(tmp_path))
    rng = np.random.default_rng(13)
    recs = _letter_fastq(case, rng, 60)
    p = os.path.join(str(tmp_path), "bad.in")
    fq = lambda n, s, q: b"@" + n + b"\n" + s + b"\n+\n" + q + b"\n"   # noqa: E731
    args = ["-Q", "--qv-offset", "33", "bad.in", "genome.fa"]
    with open(p, "wb") as f:
        if kind == "fastq_as_fasta":
            for n, s, q in recs:
                f.write(fq(n, s, q))
            args = ["--no-autodetect-input", "bad.in", "genome.fa"]
        elif kind == "short_qual":   # (entry 58: behind the start of the read-ahead thread, which run() switches on)
            for k, (n, s, q) in enumerate(recs):
                f.write(fq(n, s, q[:-3] if k == 58 else q))
        elif kind == "empty_seq":
            for k, (n, s, q) in enumerate(recs):
                f.write(b">" + n + b"\n" + (b"" if k == 30 else s + b"\n"))
            args = ["bad.in", "genome.fa"]
        elif kind == "truncated":
            blob = b"".join(fq(n, s, q) for n, s, q in recs)
            f.write(blob[:len(blob) - 40])
        elif kind == "blank_tail":
            for n, s, q in recs:
                f.write(b">" + n + b"\n" + s + b"\n")
            f.write(b"\n\n")
            args = ["bad.in", "genome.fa"]
    ref, _, rc_ref = run(REF, "gmapper-ls", args, str(tmp_path), ok=(0, 1))
    new, _, rc_new = run(NEW, "gmapper-ls", args, str(tmp_path), 3, ok=(0, 1))
    assert rc_ref == rc_new
    same(ref, new)


@pytest.mark.parametrize("block", ["300", "4096", "70000"])
@pytest.mark.parametrize("ahead", ["0", "1"])
def test_entries_across_block_ends(block, ahead, tmp_path, monkeypatch):
    """the reader's block made so small that most entries (300 bytes: every one) straddle a block end: the in-place
    path hands those to the piece-wise path, which carries the tail over; FASTQ and folded FASTA"""
    monkeypatch.setenv("SHRIMP_B200_READER_BLOCK", block)
    monkeypatch.setenv("SHRIMP_B200_READ_AHEAD_AFTER", "50" if ahead == "1" else "1000000000")
    case = LsCase("c1_small")
    case.write_fasta(str(tmp_path))
    rng = np.random.default_rng(21)
    recs = _letter_fastq(case, rng, 500)
    with open(os.path.join(str(tmp_path), "b.fq"), "wb") as f:
        for name, s, q in recs:
            f.write(b"@" + name + b"\n" + s + b"\n+\n" + q + b"\n")
    with open(os.path.join(str(tmp_path), "b.fa"), "wb") as f:
        for k, (name, s, _) in enumerate(recs):
            f.write(b">" + name + b"\n" + (b"\n".join(s[i:i + 40] for i in range(0, len(s), 40)) if k % 3 == 0 else s) + b"\n")
    for fn in ("b.fq", "b.fa"):   # and no newline at the end of either file: the last piece is what is left of the block
        blob = open(os.path.join(str(tmp_path), fn), "rb").read()
        open(os.path.join(str(tmp_path), fn), "wb").write(blob.rstrip(b"\n"))
    for args in (["-Q", "--qv-offset", "33", "b.fq", "genome.fa"], ["b.fa", "genome.fa"]):
        ref, _, _ = run(REF, "gmapper-ls", args, str(tmp_path))
        new, _, _ = run(NEW, "gmapper-ls", args, str(tmp_path), 3)
        same(ref, new)
        assert sum(1 for ln in new if ln and not ln.startswith(b"@")) > 300


def test_without_read_ahead(tmp_path, monkeypatch):
    """the reader called in place (no read-ahead thread), chunks of 100 reads"""
    monkeypatch.setenv("SHRIMP_B200_READ_AHEAD_AFTER", "1000000000")
    case = subset(LsCase("c2_small_fastq_mq"), 900)
    case.write_fasta(str(tmp_path))
    args = ["-Q", "-K", "100", "reads.fa", "genome.fa"]
    ref, _, _ = run(REF, case.binary, args, str(tmp_path))
    new, _, _ = run(NEW, case.binary, args, str(tmp_path), 3)
    same(ref, new)


def test_fast_io_replaces_exactly_the_named_reference_symbols():
    b = os.path.join(ROOT, "integration", "_build")
    out = subprocess.run(["nm", "--defined-only", os.path.join(b, "fast_io.o")], stdout=subprocess.PIPE,
                         check=True).stdout.decode()
    have = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert have == {"_Z10hit_outputP10read_entryP8read_hitS2_bPiib", "_Z11fasta_closeP8_fasta_t",
                    "_Z30fasta_get_next_read_with_rangeP8_fasta_tP10read_entry",
                    "_Z26fasta_sequence_to_bitfieldP8_fasta_tPc", "_Z26reverse_complement_read_csPjaajb"}, have
    weak = subprocess.run(["nm", os.path.join(b, "weak", "gmapper_output.o")], stdout=subprocess.PIPE,
                          check=True).stdout.decode()
    assert " W _Z10hit_outputP10read_entryP8read_hitS2_bPiib" in weak and " T shrimp_ref_hit_output" in weak
