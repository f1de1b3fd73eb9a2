import os, subprocess, sys, tempfile
sys.path.insert(0, ".")
import bench
w = bench.WORKLOADS["c2"]
n = 2_000_000
codes, _ = w.reads(n, 4242)
ctx = bench.build_context(w, 0)[0]
with tempfile.TemporaryDirectory() as d:
    bench.reference_setup(w, d, codes, ctx)
    ctx.close()
    for th, ck in ((1, 250000), (8, 125000)):
        env = dict(os.environ, SHRIMP_TIMING="1", SHRIMP_B200_VERBOSE="1")
        r = subprocess.run([os.path.abspath("integration/_build/gmapper-cs"), "-N", str(th), "-K", str(ck), "-L", "proj", "reads.fa"],
                           cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env)
        print("== threads", th, "chunk", ck)
        for ln in r.stderr.splitlines():
            if ln.startswith("[") or "Mapping Time" in ln:
                print(ln[:300])
