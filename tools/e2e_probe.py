"""Where an end-to-end step goes: wall time of shrimp_gpu_map_reads on one host thread vs the device stage times."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch
from shrimp_b200.api import MapParams, auto_list_cutoff

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
if w.key == "c3":
    w.resize(300)
n = int(sys.argv[2]) if len(sys.argv) > 2 else w.default_reads
codes, initbp_np = w.reads(n, 2)
packed = torch.from_numpy(bench.pack_rows(codes)).pin_memory().numpy()
read_len = torch.full((n,), w.read_len, dtype=torch.int32).pin_memory().numpy()
initbp = torch.from_numpy(initbp_np).pin_memory().numpy() if initbp_np is not None else None
ctx, scores, seeds, _ = bench.build_context(w, 0)
params = MapParams(list_cutoff=auto_list_cutoff(w.genome_len, 12), compute_mapping_qualities="--no-mapping-qualities" not in w.args,
                   match_mode=4 if w.paired else 2)
f = (lambda: ctx.map_pairs(params, scores, packed, read_len, reuse_buffers=True)) if w.paired else \
    (lambda: ctx.map_reads(params, scores, packed, read_len, initbp=initbp, reuse_buffers=True))
f(); f()
for _ in range(3):
    ctx.stage_times_reset()
    t0 = time.perf_counter()
    f()
    wall = (time.perf_counter() - t0) * 1e3
    st = ctx.stage_times()
    dev = sum(v[0] for v in st.values())
    print("wall %.1f ms, device stages %.1f ms, rest (H2D, D2H, host pass 2, python) %.1f ms" % (wall, dev, wall - dev),
          {k: round(v[0], 1) for k, v in st.items() if v[0] > 0.05}, "omp threads", os.environ.get("OMP_NUM_THREADS"))
