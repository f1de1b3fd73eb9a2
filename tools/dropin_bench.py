#!/usr/bin/env python
"""FASTA -> SAM through the linked drop-in (integration/_build/gmapper-{ls,cs}: the reference's unchanged gmapper.c /
output.c / fasta.c objects + the shims + libshrimp_b200.so) on a bench.py workload, next to the reference binary on
the same files: reads/s from each binary's own "Read Mapping Time", and a byte-for-byte diff of the SAM bodies.

    python tools/dropin_bench.py --workload c2 --reads 2000000 --threads 8 --chunk 250000 --diff 100000
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NEW_DIR = os.path.join(ROOT, "integration", "_build")


def run_binary(bindir, w, workdir, threads, reads_fa, extra, sam_path=None, env=None):
    rd = ["-1", reads_fa + ".1", "-2", reads_fa + ".2"] if w.paired else [reads_fa]
    cmd = [os.path.join(bindir, w.binary), "-N", str(threads), *extra, *w.load_args(), "-L", "proj", *rd]
    t0 = time.time()
    out = open(sam_path, "wb") if sam_path else subprocess.DEVNULL
    r = subprocess.run(cmd, cwd=workdir, stdout=out, stderr=subprocess.PIPE, text=True, env=env)
    wall = time.time() - t0
    if sam_path:
        out.close()
    if r.returncode != 0:
        raise RuntimeError(f"{cmd} failed: {r.stderr[-1500:]}")
    m = re.search(r"Read Mapping Time:\s+([0-9.]+) seconds", r.stderr)
    return float(m.group(1)), wall, r.stderr


def run_split(bindir, w, d, codes, procs, threads, chunk, env):
    """the read set cut into `procs` files, one gmapper process each, all started together: aggregate reads/s =
    all reads / the longest "Read Mapping Time" (the processes load the same projection and start mapping together)"""
    n = codes.shape[0]
    per = ((n // procs) + 1) & ~1
    jobs = []
    for p in range(procs):
        part = codes[p * per:(p + 1) * per]
        if part.shape[0] == 0:
            continue
        w.write_reads_fasta(os.path.join(d, f"part{p}.fa"), part)
        rd = ["-1", f"part{p}.fa.1", "-2", f"part{p}.fa.2"] if w.paired else [f"part{p}.fa"]
        cmd = [os.path.join(bindir, w.binary), "-N", str(threads), "-K", str(chunk), *w.load_args(), "-L", "proj", *rd]
        jobs.append((part.shape[0], cmd))
    t0 = time.time()
    running = [(nr, subprocess.Popen(cmd, cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env))
               for nr, cmd in jobs]
    secs = []
    for nr, pr in running:
        err = pr.communicate()[1]
        if pr.returncode != 0:
            raise RuntimeError("split run failed: " + err[-800:])
        secs.append(float(re.search(r"Read Mapping Time:\s+([0-9.]+) seconds", err).group(1)))
    wall = time.time() - t0
    return {"procs": procs, "threads_each": threads, "chunk": chunk, "map_s_max": max(secs), "wall_s": wall,
            "reads_per_s": n / max(secs)}


def sam_body(path):
    with open(path, "rb") as f:
        return [ln for ln in f.read().split(b"\n") if not ln.startswith(b"@PG")]


def main():
    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--threads", default="8")
    ap.add_argument("--chunk", default="250000")
    ap.add_argument("--diff", type=int, default=0, help="also diff the SAM of the first N reads against the reference")
    ap.add_argument("--genome-mb", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--variants", default="fast", help="comma list: fast = the drop-in with fast_io.cpp (f2), refio = the "
                    "same link line with the reference's own FASTA reader and SAM formatter")
    ap.add_argument("--procs", default="", help="also run the read file split over P concurrent processes (comma list), "
                    "the reference's own way of scaling past its serial parser (SPLITTING_AND_MERGING)")
    a = ap.parse_args()
    w = bench.WORKLOADS[a.workload]
    if a.genome_mb or w.key == "c3":
        w.resize(a.genome_mb or 300)
    ncores = os.cpu_count() or 1
    codes, _ = w.reads(a.reads, 4242)
    ctx = bench.build_context(w, 0)[0]
    out = {"workload": w.key, "reads": a.reads, "cores": ncores, "runs": []}
    env = dict(os.environ, SHRIMP_B200_GPUS=str(a.gpus), SHRIMP_B200_VERBOSE="1")
    with tempfile.TemporaryDirectory() as d:
        bench.reference_setup(w, d, codes, ctx)
        ctx.close()
        for var, th, ck in [(v, int(t), int(c)) for v in a.variants.split(",") for t in a.threads.split(",")
                            for c in a.chunk.split(",")]:
            if True:
                bindir = NEW_DIR if var == "fast" else os.path.join(NEW_DIR, var)
                s, wall, err = run_binary(bindir, w, d, th, "reads.fa", ["-K", str(ck)], env=env)
                out["runs"].append({"variant": var, "threads": th, "chunk": ck, "map_s": s, "wall_s": wall,
                                    "reads_per_s": a.reads / s})
                print(json.dumps(out["runs"][-1]), flush=True)
                for ln in [x for x in err.splitlines() if x.startswith("[gmapper-b200] thread")][:3]:
                    print("   ", ln, flush=True)
                for ln in [x for x in err.splitlines() if x.startswith("[gmapper-b200] reader") or "Wait Time" in x or
                           "Read Load Time" in x or "Fasta Lib Time" in x]:
                    print("   ", ln.strip(), flush=True)
        for P in [int(x) for x in a.procs.split(",") if x]:
            out["runs"].append(run_split(NEW_DIR, w, d, codes, P, max(1, ncores // P), int(a.chunk.split(",")[0]), env))
            print(json.dumps(out["runs"][-1]), flush=True)
        if a.diff:
            n = a.diff & ~1
            w.write_reads_fasta(os.path.join(d, "sub.fa"), codes[:n])
            s_ref, _, _ = run_binary(bench.REF_DIR, w, d, ncores, "sub.fa", [], sam_path=os.path.join(d, "ref.sam"))
            s_new, _, err = run_binary(NEW_DIR, w, d, 4, "sub.fa", ["-K", str(max(1000, n // 7) & ~1)],
                                       sam_path=os.path.join(d, "new.sam"), env=env)
            ref, new = sam_body(os.path.join(d, "ref.sam")), sam_body(os.path.join(d, "new.sam"))
            bad = [i for i, (x, y) in enumerate(zip(ref, new)) if x != y]
            out["diff"] = {"reads": n, "ref_lines": len(ref), "new_lines": len(new), "differing_lines": len(bad),
                           "reference_reads_per_s": n / s_ref, "reference_threads": ncores,
                           "dropin_reads_per_s": n / s_new}
            for i in bad[:3]:
                print("DIFF", ref[i][:300], new[i][:300], sep="\n", flush=True)
            if len(ref) != len(new):
                print("LINES", len(ref), len(new), flush=True)
            tail = [ln for ln in err.splitlines() if "Time" in ln or "Invocations" in ln or "Cells" in ln]
            out["dropin_stats_tail"] = tail[:12]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
