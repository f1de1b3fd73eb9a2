"""Scratch: sw_vector kernel throughput on uniform batches (C1 50x70, C5 75x112, C3 100x140)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import shrimp_b200
from shrimp_b200.api import _pack_codes

ctx = shrimp_b200.GpuContext(0)
print("dpx peak G thread-instr/s:", ctx.dpx_peak())
rng = np.random.default_rng(0)
G = 1 << 22
g = rng.integers(0, 4, size=G).astype(np.uint32)
gp = _pack_codes(g)
for rl, gl, n in [(50, 70, 2_000_000), (75, 112, 1_000_000), (100, 140, 1_000_000), (36, 50, 2_000_000)]:
    stride = (rl + 7) // 8
    nr = 100_000
    reads = rng.integers(0, 2**32, size=(nr, stride), dtype=np.uint64).astype(np.uint32) & 0x33333333
    off = rng.integers(0, G - gl, size=n).astype(np.uint32)
    ridx = (np.arange(n) % nr).astype(np.int32)
    ctx.sw_setup(400, 200, shrimp_b200.LS_DEFAULT_SCORES)
    for rep in range(3):
        ctx.stage_times_reset()
        t0 = time.time()
        sc = ctx.sw_vector(gp, off, np.full(n, gl, np.int32), reads, ridx, np.full(n, rl, np.int32))
        t1 = time.time()
        ms = ctx.stage_times()["sw_vector"][0]
    cells = n * rl * gl
    print(f"{rl}x{gl} n={n}: kernel {ms:.3f} ms -> {cells/ms/1e6:.1f} GCUPS; e2e {(t1-t0)*1e3:.1f} ms -> {cells/(t1-t0)/1e9:.1f} GCUPS")
