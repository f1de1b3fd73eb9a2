import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mapcases import MAP_CASES, LsCase
from oracle import pipeline as op
import shrimp_b200
from shrimp_b200 import align
from shrimp_b200.api import MapParams, _pack_codes
name = sys.argv[1] if len(sys.argv) > 1 else "c2_small_fastq_mq"
case = LsCase(name)
ctx = shrimp_b200.GpuContext(0)
ctx.sw_setup(1500, 1000, case.scores, use_colours=case.colour, anchor_width=case.anchor_width)
ctx.load_genome([_pack_codes(c.astype(np.uint32)) for c in case.contig_codes], [c.size for c in case.contig_codes], colour_space=case.colour)
ctx.build_index(case.seeds, hflag=case.hflag)
params = MapParams(list_cutoff=case.list_cutoff, **MAP_CASES[name]["opts"])
res = ctx.map_reads(params, case.scores, case.packed, case.read_len, initbp=case.initbp, crossover_scores=case.crossover_scores, quals=case.quals)
g = op.Genome(case.contig_codes, case.colour); ix = op.Index(g, case.seeds, hflag=case.hflag)
opts = op.MapOptions(scores=case.scores, colour_space=case.colour, list_cutoff=case.list_cutoff, anchor_width=case.anchor_width, **MAP_CASES[name]["opts"])
hits, nper, _, _ = op.map_reads(g, ix, opts, case.packed, case.read_len, initbp=case.initbp, crossover_scores=case.crossover_scores, quals=case.quals)
print(len(res.hits), len(hits))
nbad = 0
for a, b in zip(res.hits, hits):
    e0, el, rm = int(a["edit_off"]), int(a["edit_len"]), int(a["rmapped"])
    seq, q = align.post_sw_seq_qual(res.edits[e0:e0 + el], res.edits[e0 + el:e0 + el + rm], False, True)
    oq = bytes(b["sfr"]["qralign"]).split(b"\0")[0]
    oseq = bytes(c for c in oq if c != ord("-")).upper().decode()
    oqual = bytes(b["sfr"]["qual"]).split(b"\0")[0].decode()
    same = (int(a["mismatches"]) == int(b["sfr"]["mismatches"]) and seq == oseq and q == oqual)
    rel = abs(float(a["posterior"]) - float(b["posterior"])) / max(float(b["posterior"]), 1e-300)
    if not same:
        nbad += 1
        if nbad <= 6:
            print("read", int(a["read_idx"]), "post", float(a["posterior"]), float(b["posterior"]), "rel", rel, "mm", int(a["mismatches"]), int(b["sfr"]["mismatches"]), "rs", int(a["read_start"]))
            print("  gpu seq ", seq, q); print("  orc seq ", oseq, oqual)
            print("  db      ", bytes(b["sfr"]["dbalign"]).split(b"\0")[0].decode()); print("  read", bytes(case.reads[int(a["read_idx"])][1]), bytes(case.quals[int(a["read_idx"])]))
print("differing hits", nbad, "of", len(hits))
rels = [abs(float(a["posterior"]) - float(b["posterior"])) / max(float(b["posterior"]), 1e-300) for a, b in zip(res.hits, hits)]
print("max rel posterior diff", max(rels), "exact equal", sum(1 for r in rels if r == 0))
