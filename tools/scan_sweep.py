"""Sweep the seed-scan launch knobs (test hooks read by chunk_scan through getenv) on one resident batch.

    python tools/scan_sweep.py --workload c3 --reads 100000 [--genome-mb 300] KEY=v1,v2 KEY2=...

Builds the index once, maps the same resident batch under every combination and prints the seed_scan stage time.
"""
import argparse
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--reads", type=int, default=100000)
    ap.add_argument("--genome-mb", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("knobs", nargs="*")
    a = ap.parse_args()
    import torch
    from shrimp_b200.api import MapParams, auto_list_cutoff
    w = bench.WORKLOADS[a.workload]
    if a.genome_mb or w.key == "c3":
        w.resize(a.genome_mb or 300)
    codes, initbp_np = w.reads(a.reads, 2)
    packed = torch.from_numpy(bench.pack_rows(codes)).pin_memory().numpy()
    read_len = torch.full((a.reads,), w.read_len, dtype=torch.int32).pin_memory().numpy()
    initbp = torch.from_numpy(initbp_np).pin_memory().numpy() if initbp_np is not None else None
    ctx, scores, seeds, index_s = bench.build_context(w, 0)
    print("index built in %.1f s" % index_s, flush=True)
    params = MapParams(list_cutoff=auto_list_cutoff(w.genome_len, 12), compute_mapping_qualities="--no-mapping-qualities" not in w.args,
                       match_mode=4 if w.paired else 2)
    if w.paired:
        ctx.map_pairs(params, scores, packed, read_len, reuse_buffers=True)
        run = lambda: ctx.map_pairs_resident(params, scores)
    else:
        ctx.map_reads(params, scores, packed, read_len, initbp=initbp, reuse_buffers=True)
        run = lambda: ctx.map_resident(params, scores)
    keys = [k.split("=")[0] for k in a.knobs]
    vals = [k.split("=")[1].split(",") for k in a.knobs]
    for combo in itertools.product(*vals) if keys else [()]:
        for k, v in zip(keys, combo):
            if v == "-":
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        run()
        ctx.stage_times_reset()
        for _ in range(a.reps):
            ctx.flush_l2()
            st = run()
        t = ctx.stage_times()
        print(dict(zip(keys, combo)), "scan %.2f ms/step" % (t["seed_scan"][0] / a.reps),
              {k: round(v[0] / a.reps, 2) for k, v in t.items() if k != "seed_scan" and v[0] > 0}, flush=True)


if __name__ == "__main__":
    main()
