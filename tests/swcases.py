"""Seeded random sw_vector / sw_full cases shared by the oracle and GPU parity tests."""
from __future__ import annotations

import numpy as np

from shrimp_b200.api import _pack_codes


def random_codes(rng, n, n_frac=0.0, alphabet=4):
    c = rng.integers(0, alphabet, size=n).astype(np.uint32)
    if n_frac > 0:
        c[rng.random(n) < n_frac] = 15
    return c


def mutate(rng, codes, sub=0.05, indel=0.02):
    out = []
    for c in codes:
        r = rng.random()
        if r < indel / 2:
            continue
        if r < indel:
            out.append(int(rng.integers(0, 4)))
        if rng.random() < sub:
            out.append(int((c + rng.integers(1, 4)) % 4) if c < 4 else int(c))
        else:
            out.append(int(c))
    return np.array(out, dtype=np.uint32)


def make_vector_cases(seed, n_tasks, rlen_range=(20, 120), glen_factor=(1.0, 1.6), n_frac=0.01,
                      genome_len=20000, colour=False):
    """Returns dict with a packed genome (and letter genome for colour), packed reads [n, stride], task arrays."""
    rng = np.random.default_rng(seed)
    g_ls = random_codes(rng, genome_len, n_frac)
    if colour:
        prev = np.concatenate([[3], g_ls[:-1]])
        g = np.where((g_ls > 3) | (prev > 3), 15, g_ls ^ prev).astype(np.uint32)
    else:
        g = g_ls
    max_r = rlen_range[1]
    stride = (max_r + 7) // 8
    reads = np.zeros((n_tasks, stride), dtype=np.uint32)
    goff = np.zeros(n_tasks, dtype=np.uint32)
    glen = np.zeros(n_tasks, dtype=np.int32)
    rlen = np.zeros(n_tasks, dtype=np.int32)
    initbp = np.zeros(n_tasks, dtype=np.int8)
    for t in range(n_tasks):
        rl = int(rng.integers(rlen_range[0], rlen_range[1] + 1))
        gl = int(rl * rng.uniform(*glen_factor))
        gl = max(1, min(gl, genome_len - 1))
        off = int(rng.integers(0, genome_len - gl))
        kind = rng.random()
        if kind < 0.7:
            # implant: read derived from inside the window
            s = off + int(rng.integers(0, max(1, gl - rl + 1)))
            src = g[s:s + rl + 8]
            rd = mutate(rng, src, sub=rng.uniform(0, 0.12), indel=rng.uniform(0, 0.06))[:rl]
            if rd.size < rl:
                rd = np.concatenate([rd, random_codes(rng, rl - rd.size)])
        else:
            rd = random_codes(rng, rl, n_frac)
        if rng.random() < 0.05:
            rd = rd.copy()
            rd[int(rng.integers(0, rl))] = 15
        reads[t, :] = _pack_codes(rd.astype(np.uint32), stride)
        goff[t], glen[t], rlen[t] = off, gl, rl
        initbp[t] = int(rng.integers(0, 4))
    return dict(genome=_pack_codes(g), genome_ls=_pack_codes(g_ls) if colour else None, reads=reads,
                goff=goff, glen=glen, rlen=rlen, read_idx=np.arange(n_tasks, dtype=np.int32), initbp=initbp)
