import os, subprocess, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from mapcases import PairCase
c = PairCase("c3_small")
d = "/tmp/pm"
c.write_fasta(d)
for mode in ["opp-out", "col-fw", "col-bw"]:
    for t in ["1", "2"]:
        r = subprocess.run([os.path.abspath("integration/_build/gmapper-ls"), "-N", t, "-K", "300", "-p", mode, "-I", "0,1000", "-1", "m1.fa", "-2", "m2.fa", "genome.fa"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        print("MODE", mode, "threads", t, "rc", r.returncode)
        print("\n".join(r.stderr.decode(errors="replace").splitlines()[-6:]))
