"""How much of the device-resident step overlaps across contexts (one stream each, shared index)?

    python tools/resident_overlap.py --workload c2 --reads 1000000 --contexts 1,2,3 --steps 4

Every context keeps its own batch of --reads reads resident and maps it --steps times from its own host thread;
prints reads/s over the wall clock of all threads (device synchronised on both sides)."""
import argparse
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--reads", type=int, default=1000000)
    ap.add_argument("--contexts", default="1,2,3")
    ap.add_argument("--steps", type=int, default=4)
    a = ap.parse_args()
    import torch
    import shrimp_b200
    from shrimp_b200.api import MapParams, auto_list_cutoff
    w = bench.WORKLOADS[a.workload]
    ctx, scores, seeds, _ = bench.build_context(w, 0)
    params = MapParams(list_cutoff=auto_list_cutoff(w.genome_len, 12),
                       compute_mapping_qualities="--no-mapping-qualities" not in w.args, match_mode=4 if w.paired else 2)
    read_len = torch.full((a.reads,), w.read_len, dtype=torch.int32).pin_memory().numpy()
    ctxs = [ctx]
    n_max = max(int(x) for x in a.contexts.split(","))
    for i in range(1, n_max):
        cx = shrimp_b200.GpuContext(0)
        cx.sw_setup(1400, 1000, scores, use_colours=w.colour)
        cx.share_genome_from(ctx)
        ctxs.append(cx)
    runs = []
    for i, cx in enumerate(ctxs):
        codes, initbp_np = w.reads(a.reads, 2 + i)
        packed = torch.from_numpy(bench.pack_rows(codes)).pin_memory().numpy()
        initbp = torch.from_numpy(initbp_np).pin_memory().numpy() if initbp_np is not None else None
        if w.paired:
            cx.map_pairs(params, scores, packed, read_len, reuse_buffers=True)
            runs.append(lambda cx=cx: cx.map_pairs_resident(params, scores))
        else:
            cx.map_reads(params, scores, packed, read_len, initbp=initbp, reuse_buffers=True)
            runs.append(lambda cx=cx: cx.map_resident(params, scores))
        runs[-1]()
    for n in [int(x) for x in a.contexts.split(",")]:
        def work(fn):
            for _ in range(a.steps):
                fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ths = [threading.Thread(target=work, args=(runs[i],)) for i in range(n)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("contexts %d: %.2f M reads/s (%.1f ms per 1 M reads)" % (n, n * a.reads * a.steps / dt / 1e6,
                                                                       dt * 1e3 / (n * a.steps) * 1e6 / a.reads), flush=True)


if __name__ == "__main__":
    main()
