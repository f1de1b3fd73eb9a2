#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ from the REFERENCE itself.

Run in the build container only (needs /root/reference compiled into oracle/_ref by
`make -C oracle ref`).  The vectors are small and committed; the GPU box never needs the
reference sources.  Usage: python tests/golden/make_golden.py [what ...]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402


def golden_sw_vector():
    from swcases import make_vector_cases
    from test_oracle_sw_vector import SCORE_SETS
    for k, (name, (sc, colour)) in enumerate(sorted(SCORE_SETS.items())):
        seed, n = 1000 + k, 600
        cases = make_vector_cases(seed=seed, n_tasks=n, colour=colour)
        ref = oracle.RefSw(400, 200, sc, colour)
        scores = np.array([
            ref.sw_vector(cases["genome"], cases["goff"][t], cases["glen"][t], cases["reads"][t], cases["rlen"][t],
                          cases["genome_ls"] if colour else None, cases["initbp"][t] if colour else -1)
            for t in range(n)], dtype=np.int32)
        np.savez_compressed(os.path.join(HERE, f"sw_vector_{name}.npz"), seed=seed, n_tasks=n, scores=scores)
        print("wrote sw_vector", name, scores[:8])


TARGETS = {"sw_vector": golden_sw_vector}

if __name__ == "__main__":
    assert oracle.have_ref(), "build oracle/_ref first: make -C oracle ref"
    for w in (sys.argv[1:] or sorted(TARGETS)):
        TARGETS[w]()
