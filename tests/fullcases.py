"""Seeded random full-SW cases (letter and colour space) shared by oracle and GPU tests."""
from __future__ import annotations

import numpy as np

from shrimp_b200.api import _pack_codes
from swcases import mutate, random_codes


def make_full_cases(seed, n, colour=False, rlen_range=(25, 80), n_frac=0.01):
    """list of dicts: genome (packed letters), goff, glen, read (packed), rlen, initbp, anchor (x,y,len,width),
    vscore placeholder.  The read derives from the window so that alignments are non-trivial."""
    rng = np.random.default_rng(seed)
    G = 6000
    g = random_codes(rng, G, n_frac)
    gp = _pack_codes(g)
    out = []
    for _ in range(n):
        rl = int(rng.integers(rlen_range[0], rlen_range[1] + 1))
        gl = int(rl * rng.uniform(1.2, 1.5))
        off = int(rng.integers(0, G - gl - 8))
        s = int(rng.integers(0, gl - rl + 1))
        letters = mutate(rng, g[off + s: off + s + rl + 10], sub=rng.uniform(0, 0.08), indel=rng.uniform(0, 0.05))[:rl]
        if letters.size < rl:
            letters = np.concatenate([letters, random_codes(rng, rl - letters.size)])
        if rng.random() < 0.1:
            letters = letters.copy()
            letters[int(rng.integers(0, rl))] = 15
        initbp = int(rng.integers(0, 4))
        if colour:
            prev = np.concatenate([[initbp], letters[:-1]])
            rd = np.where((letters > 3) | (prev > 3), 15, letters ^ prev).astype(np.uint32)
            err = rng.random(rl) < rng.uniform(0, 0.06)
            rd[err & (rd < 4)] = (rd[err & (rd < 4)] + rng.integers(1, 4, size=int((err & (rd < 4)).sum()))) % 4
            if rng.random() < 0.1:
                rd[int(rng.integers(0, rl))] = 15
        else:
            rd = letters.astype(np.uint32)
        # a plausible anchor: a diagonal segment near the true diagonal, sometimes joined (width > 1)
        alen = int(rng.integers(8, max(9, rl - 4)))
        ay = int(rng.integers(0, max(1, rl - alen)))
        ax = s + ay + int(rng.integers(-2, 3))
        aw = 1 if rng.random() < 0.7 else int(rng.integers(2, 7))
        out.append(dict(genome=gp, goff=off, glen=gl, read=_pack_codes(rd, (rlen_range[1] + 7) // 8), rlen=rl,
                        initbp=initbp, anchor=(ax, ay, alen, aw), revcmpl=int(rng.integers(0, 2))))
    return out
