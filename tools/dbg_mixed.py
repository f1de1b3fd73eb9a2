import os, sys, tempfile
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "tools")
from test_gpu_dropin import _mixed_reads, run_sam, REF, NEW
from mapcases import LsCase
case = LsCase("c2_small")
d = tempfile.mkdtemp()
case.write_fasta(d)
rng = np.random.default_rng(77)
reads = _mixed_reads(case, rng, 900, 25, 120, 0.01, True)
with open(os.path.join(d, "mixed.fq"), "wb") as f:
    for name, s, q in reads:
        f.write(b"@" + name.encode() + b"\n" + s + b"\n+\n" + q + b"\n")
for extra in ([], ["--no-mapping-qualities"], ["--ignore-qvs"]):
    args = ["-Q", "--qv-offset", "33", "--longest-read", "380", *extra, "mixed.fq", "genome.fa"]
    ref, _ = run_sam(REF, case.binary, args, d, 4)
    new, _ = run_sam(NEW, case.binary, args, d, 2, ["-K", "250"])
    bad = [i for i, (a, b) in enumerate(zip(ref, new)) if a != b]
    print("EXTRA", extra, "lines", len(ref), len(new), "bad", len(bad))
    lens = {}
    for i in bad:
        fa = ref[i].split(b"\t"); fb = new[i].split(b"\t")
        cols = [k for k, (x, y) in enumerate(zip(fa, fb)) if x != y]
        key = tuple(cols)
        lens.setdefault(key, []).append(i)
    for key, idx in list(lens.items())[:6]:
        i = idx[0]
        fa = ref[i].split(b"\t"); fb = new[i].split(b"\t")
        print("  cols", key, "count", len(idx), "read", fa[0], "len CS", len(fa[-3]) if len(fa) > 3 else 0)
        for k in key[:4]:
            print("     ref", fa[k][:160]); print("     new", fb[k][:160])
