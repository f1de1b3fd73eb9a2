"""GPU parity: genome arrays and the HBM-resident projection vs the oracle (= the reference's -S content)."""
import numpy as np
import pytest

from mapcases import LsCase
from oracle import pipeline as op
from shrimp_b200 import seeds as S
from shrimp_b200.api import _pack_codes

pytestmark = pytest.mark.gpu


def _global_pack(contig_codes):
    return _pack_codes(np.concatenate(contig_codes).astype(np.uint32))


def _unpack(words, n):
    sh = (4 * np.arange(8, dtype=np.uint32))[None, :]
    return ((words[:, None] >> sh) & 15).reshape(-1)[:n]


def _oracle_array(ptrs, lens):
    """concatenate the oracle's per-contig packed arrays into global nibble coordinates"""
    import ctypes as C
    out = []
    for c, n in enumerate(lens):
        w = np.ctypeslib.as_array(ptrs[c], shape=((int(n) + 7) // 8,))
        out.append(_unpack(w.copy(), int(n)))
    return np.concatenate(out)


@pytest.mark.parametrize("colour", [False, True])
def test_genome_arrays_and_index_match_oracle(gpu_ctx, colour):
    import ctypes as C
    rng = np.random.default_rng(21)
    # ragged contigs incl. N runs, lengths not multiples of 8, one contig shorter than the seeds
    lens = [4001, 13, 2500, 777, 9999]
    codes = []
    for n in lens:
        c = rng.integers(0, 4, size=n).astype(np.uint8)
        if n > 100:
            for _ in range(3):
                p = int(rng.integers(0, n - 30))
                c[p:p + int(rng.integers(1, 25))] = 15
        codes.append(c)
    seeds = S.load_default_seeds() + [S.add_spaced_seed("1101")]
    g = op.Genome(codes, colour)
    ix = op.Index(g, seeds)
    gpu_ctx.load_genome([_pack_codes(c.astype(np.uint32)) for c in codes], lens, colour_space=colour)
    gpu_ctx.build_index(seeds)
    tot = sum(lens)

    class G(C.Structure):
        _fields_ = [("num_contigs", C.c_int), ("colour_space", C.c_int), ("contig_offsets", C.POINTER(C.c_uint32)),
                    ("genome_len", C.POINTER(C.c_uint32)), ("ls", C.POINTER(C.POINTER(C.c_uint32))),
                    ("ls_rc", C.POINTER(C.POINTER(C.c_uint32))), ("cs", C.POINTER(C.POINTER(C.c_uint32))),
                    ("cs_rc", C.POINTER(C.POINTER(C.c_uint32))), ("total_len", C.c_uint64)]
    gs = C.cast(g.h, C.POINTER(G)).contents
    assert np.array_equal(_unpack(gpu_ctx.genome_export(0), tot), _oracle_array(gs.ls, lens))
    assert np.array_equal(_unpack(gpu_ctx.genome_export(1), tot), _oracle_array(gs.ls_rc, lens))
    if colour:
        assert np.array_equal(_unpack(gpu_ctx.genome_export(2), tot), _oracle_array(gs.cs, lens))
        assert np.array_equal(_unpack(gpu_ctx.genome_export(3), tot), _oracle_array(gs.cs_rc, lens))
    for sn in range(len(seeds)):
        lens_g, pos_g = gpu_ctx.export_index(sn)
        assert np.array_equal(lens_g, ix.bucket_lens(sn))
        assert np.array_equal(pos_g, ix.positions(sn))


def test_index_c1_small_matches_oracle(gpu_ctx):
    case = LsCase("c1_small")
    g = op.Genome(case.contig_codes, False)
    ix = op.Index(g, case.seeds)
    gpu_ctx.load_genome([_pack_codes(c.astype(np.uint32)) for c in case.contig_codes],
                        [c.size for c in case.contig_codes])
    gpu_ctx.build_index(case.seeds)
    for sn in range(len(case.seeds)):
        lens_g, pos_g = gpu_ctx.export_index(sn)
        assert np.array_equal(lens_g, ix.bucket_lens(sn))
        assert np.array_equal(pos_g, ix.positions(sn))


def test_hashed_seeds_match_oracle(gpu_ctx):
    """-H / mirna seeds (gmapper.h:323-336): 5 seeds of span 20 hashed into 4^12 buckets"""
    rng = np.random.default_rng(5)
    codes = [rng.integers(0, 4, size=3000).astype(np.uint8)]
    seeds = S.load_default_mirna_seeds()
    g = op.Genome(codes, False)
    ix = op.Index(g, seeds, hflag=True)
    gpu_ctx.load_genome([_pack_codes(codes[0].astype(np.uint32))], [3000])
    gpu_ctx.build_index(seeds, hflag=True)
    for sn in range(len(seeds)):
        lens_g, pos_g = gpu_ctx.export_index(sn)
        assert np.array_equal(lens_g, ix.bucket_lens(sn))
        assert np.array_equal(pos_g, ix.positions(sn))


import os  # noqa: E402

_REF = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref")


@pytest.mark.skipif(not os.path.exists(os.path.join(_REF, "gmapper-ls")), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["c1_small", "c2_small"])
def test_projection_save_is_byte_identical_to_gmapper_S(gpu_ctx, tmp_path, name):
    """shrimp_gpu_projection_save vs `gmapper -S` (save_genome_map, genome.c:185-272) on the same genome:
    every file equal after gunzip, and `gmapper -L` on our files gives the SAM of a from-FASTA run."""
    import gzip
    import subprocess
    case = LsCase(name)
    d = str(tmp_path)
    case.write_fasta(d)
    binary = os.path.join(_REF, case.binary)
    subprocess.run([binary, "-S", "ref", "genome.fa"], cwd=d, check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    gpu_ctx.load_genome([_pack_codes(c.astype(np.uint32)) for c in case.contig_codes],
                        [c.size for c in case.contig_codes], colour_space=case.colour)
    gpu_ctx.build_index(case.seeds)
    gpu_ctx.save_projection(os.path.join(d, "mine"), case.contig_names)
    for suffix in ["genome"] + [f"seed.{sn}" for sn in range(len(case.seeds))]:
        want = gzip.open(os.path.join(d, f"ref.{suffix}"), "rb").read()
        got = open(os.path.join(d, f"mine.{suffix}"), "rb").read()
        assert got == want, suffix
    args = ["--no-mapping-qualities"] if case.colour else []
    a = subprocess.run([binary, *args, "-L", "mine", "reads.fa"], cwd=d, check=True, capture_output=True).stdout
    b = subprocess.run([binary, *args, "reads.fa", "genome.fa"], cwd=d, check=True, capture_output=True).stdout
    strip = lambda s: [l for l in s.splitlines() if not l.startswith(b"@PG")]  # noqa: E731
    assert strip(a) == strip(b)
