"""Stress run of the linked drop-in: the mixed-length letter-space FASTQ case of tests/test_gpu_dropin.py (one read in
seven dropped by the loop of gmapper.c) N times with two threads and small chunks, crash back-traces switched on
(SHRIMP_B200_BACKTRACE=1).  This is the run that showed the look-ahead bound bug of round 2 (13 crashes in 60 runs).

    python tools/dropin_stress.py 60
"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, ROOT)
import test_gpu_dropin as t
from mapcases import LsCase
d = "/tmp/mixed"; os.makedirs(d, exist_ok=True)
case = LsCase("c1_small"); case.write_fasta(d)
rng = np.random.default_rng(77)
reads = t._mixed_reads(case, rng, 900, 22, 400, 0.01, False)
with open(os.path.join(d, "mixed.fq"), "wb") as f:
    for name, s, q in reads:
        f.write(b"@" + name.encode() + b"\n" + s + b"\n+\n" + q + b"\n")
args = ["-Q", "--qv-offset", "33", "--longest-read", "380", "mixed.fq", "genome.fa"]
B = os.path.join(ROOT, "integration", "_build")
import collections
out = collections.Counter()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    r = subprocess.run([os.path.join(B, "gmapper-ls"), "-N", "2", "-K", "250", *args], cwd=d, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, env=dict(os.environ, SHRIMP_B200_BACKTRACE="1"))
    out[(r.returncode, len(r.stdout))] += 1
    if r.returncode != 0:
        print(r.stderr.decode(errors="replace")[-3500:], flush=True)
print(dict(out))
