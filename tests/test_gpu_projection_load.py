"""shrimp_gpu_projection_load (the load half): the -S files -- written by this library or by the reference
`gmapper -S` (gzip) -- go straight into HBM; the resident genome arrays, the CSR of every seed and the mapping results
equal those of a context that built its projection from the genome."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import shrimp_b200  # noqa: E402
from mapcases import MAP_CASES, LsCase  # noqa: E402
from shrimp_b200.api import MapParams, _pack_codes  # noqa: E402

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "oracle", "_ref")


def _built(case):
    ctx = shrimp_b200.GpuContext(0)
    ctx.sw_setup(1500, 1000, case.scores, use_colours=case.colour, anchor_width=case.anchor_width)
    ctx.load_genome([_pack_codes(c.astype(np.uint32)) for c in case.contig_codes], [c.size for c in case.contig_codes],
                    colour_space=case.colour)
    ctx.build_index(case.seeds, hflag=case.hflag)
    return ctx


@pytest.mark.parametrize("name,source", [("c1_small", "library"), ("c2_small", "library"), ("c4_small_mirna", "library"),
                                         ("c1_small", "gmapper"), ("c2_small", "gmapper")])
def test_projection_load_equals_build(name, source, tmp_path):
    case = LsCase(name)
    a = _built(case)
    prefix = os.path.join(str(tmp_path), "proj")
    if source == "library":
        a.save_projection(prefix, case.contig_names)
    else:
        if not os.path.exists(os.path.join(REF, case.binary)):
            pytest.skip("reference binary not present")
        case.write_fasta(str(tmp_path))
        subprocess.run([os.path.join(REF, case.binary), "-S", "proj", "genome.fa"], cwd=str(tmp_path), check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    b = shrimp_b200.GpuContext(0)
    try:
        b.sw_setup(1500, 1000, case.scores, use_colours=case.colour, anchor_width=case.anchor_width)
        names = b.load_projection(prefix)
        assert names == case.contig_names
        b.genome_len = a.genome_len
        b.total_len = a.total_len
        b.colour_space = a.colour_space
        for which in range(4 if case.colour else 2):
            assert np.array_equal(a.genome_export(which), b.genome_export(which)), which
        b.seeds = a.seeds
        for sn in range(len(case.seeds)):
            la, pa = a.export_index(sn)
            lb, pb = b.export_index(sn)
            assert np.array_equal(la, lb) and np.array_equal(pa, pb), sn
        opts = dict(MAP_CASES[name]["opts"])
        params = MapParams(list_cutoff=case.list_cutoff, **opts)
        ra = a.map_reads(params, case.scores, case.packed, case.read_len, initbp=case.initbp)
        rb = b.map_reads(params, case.scores, case.packed, case.read_len, initbp=case.initbp)
        assert len(ra.hits) > 100 and len(ra.hits) == len(rb.hits) and ra.edits.tobytes() == rb.edits.tobytes()
        for f in ra.hits.dtype.names:
            if f != "hit_slot":   # the slot a hit got in the chunk's hit buffer depends on the order of the reservations
                assert np.array_equal(ra.hits[f], rb.hits[f]), f
    finally:
        b.close()
        a.close()
