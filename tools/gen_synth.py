#!/usr/bin/env python
"""Seeded synthetic genomes and simulated reads for the BASELINE.json configs.

Test/bench infrastructure only.  Everything is
generated from numpy's PCG64 with fixed seeds so the GPU box and this
container produce identical inputs without shipping data files.
"""
from __future__ import annotations

import argparse
import os
import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"):
    COMP[a] = b
# colour of a letter pair = XOR of the 2-bit letter codes (A=0,C=1,G=2,T=3)
CODE = np.full(256, 4, dtype=np.uint8)
for i, c in enumerate(b"ACGT"):
    CODE[c] = i


def make_genome(total_len: int, n_contigs: int, seed: int, n_frac: float = 0.0,
                repeat_unit: int = 0, repeat_copies: int = 0, repeat_div: float = 0.02):
    """Return list of (name, uint8 ascii array)."""
    rng = np.random.default_rng(seed)
    lens = [total_len // n_contigs] * n_contigs
    lens[-1] += total_len - sum(lens)
    contigs = []
    for ci, ln in enumerate(lens):
        seq = BASES[rng.integers(0, 4, size=ln, dtype=np.uint8)]
        if repeat_unit and repeat_copies:
            unit = BASES[rng.integers(0, 4, size=repeat_unit, dtype=np.uint8)]
            for k in range(repeat_copies):
                pos = int(rng.integers(0, max(1, ln - repeat_unit)))
                cp = unit.copy()
                mut = rng.random(repeat_unit) < repeat_div
                cp[mut] = BASES[rng.integers(0, 4, size=int(mut.sum()), dtype=np.uint8)]
                seq[pos:pos + repeat_unit] = cp[: max(0, min(repeat_unit, ln - pos))]
        if n_frac > 0:
            n_runs = max(1, int(ln * n_frac / 500))
            for _ in range(n_runs):
                pos = int(rng.integers(0, max(1, ln - 500)))
                seq[pos:pos + 500] = ord("N")
        contigs.append((f"contig{ci}", seq))
    return contigs


def _mutate(rng, frag: np.ndarray, sub: float, indel_p: float, max_indel: int) -> np.ndarray:
    frag = frag.copy()
    if sub > 0:
        m = rng.random(frag.size) < sub
        if m.any():
            # substitute with a different base
            idx = np.nonzero(m)[0]
            old = CODE[frag[idx]] & 3
            frag[idx] = BASES[(old + rng.integers(1, 4, size=idx.size)) % 4]
    if indel_p > 0 and rng.random() < indel_p and frag.size > 2 * max_indel + 4:
        k = int(rng.integers(1, max_indel + 1))
        pos = int(rng.integers(max_indel, frag.size - max_indel))
        if rng.random() < 0.5:  # deletion from the read
            frag = np.concatenate([frag[:pos], frag[pos + k:]])
        else:
            ins = BASES[rng.integers(0, 4, size=k, dtype=np.uint8)]
            frag = np.concatenate([frag[:pos], ins, frag[pos:]])
    return frag


def revcomp(a: np.ndarray) -> np.ndarray:
    return COMP[a[::-1]]


def simulate_reads(contigs, n_reads: int, read_len: int, seed: int, sub: float = 0.02,
                   indel_p: float = 0.0, max_indel: int = 5, n_base_p: float = 0.0):
    """Letter-space reads: list of (name, ascii uint8 array)."""
    rng = np.random.default_rng(seed)
    lens = np.array([c[1].size for c in contigs], dtype=np.int64)
    probs = lens / lens.sum()
    out = []
    extra = max_indel + 2 if indel_p > 0 else 0
    for i in range(n_reads):
        cn = int(rng.choice(len(contigs), p=probs)) if len(contigs) > 1 else 0
        g = contigs[cn][1]
        span = read_len + extra
        if g.size <= span:
            pos, frag = 0, g.copy()
        else:
            pos = int(rng.integers(0, g.size - span))
            frag = g[pos:pos + span]
        frag = _mutate(rng, frag, sub, indel_p, max_indel)[:read_len]
        st = int(rng.integers(0, 2))
        if st:
            frag = revcomp(frag)
        if n_base_p > 0:
            m = rng.random(frag.size) < n_base_p
            frag = frag.copy()
            frag[m] = ord("N")
        out.append((f"r{i}_{cn}_{pos}_{'-' if st else '+'}", frag))
    return out


def letters_to_colour_read(frag: np.ndarray, rng=None, col_err: float = 0.0) -> bytes:
    """SOLiD read string: primer base 'T' then colours 0-3 ('.' for unknown)."""
    codes = CODE[frag]
    prev = np.concatenate([[3], codes[:-1]])  # primer T
    col = np.where((codes > 3) | (prev > 3), 4, codes ^ prev).astype(np.uint8)
    if rng is not None and col_err > 0:
        m = (rng.random(col.size) < col_err) & (col < 4)
        idx = np.nonzero(m)[0]
        col[idx] = (col[idx] + rng.integers(1, 4, size=idx.size)) % 4
    s = bytes(b"0123."[c] for c in col)
    return b"T" + s


def simulate_cs_reads(contigs, n_reads: int, read_len: int, seed: int, snp_p: float = 0.3,
                      col_err: float = 0.03):
    rng = np.random.default_rng(seed)
    lens = np.array([c[1].size for c in contigs], dtype=np.int64)
    probs = lens / lens.sum()
    out = []
    for i in range(n_reads):
        cn = int(rng.choice(len(contigs), p=probs)) if len(contigs) > 1 else 0
        g = contigs[cn][1]
        pos = int(rng.integers(0, g.size - read_len))
        frag = g[pos:pos + read_len].copy()
        if rng.random() < snp_p:
            p = int(rng.integers(0, read_len))
            frag[p] = BASES[((CODE[frag[p]] & 3) + int(rng.integers(1, 4))) % 4]
        st = int(rng.integers(0, 2))
        if st:
            frag = revcomp(frag)
        out.append((f"r{i}_{cn}_{pos}_{'-' if st else '+'}", letters_to_colour_read(frag, rng, col_err)))
    return out


def simulate_pairs(contigs, n_pairs: int, read_len: int, seed: int, ins_mean: float = 300.0, ins_sd: float = 30.0,
                   sub: float = 0.02, indel_p: float = 0.0, max_indel: int = 3, junk_p: float = 0.05,
                   far_p: float = 0.05):
    """Opp-in pairs: mate 1 = start of the fragment, mate 2 = reverse complement of its end; the fragment is
    flipped half of the time.  junk_p: one mate replaced by random sequence; far_p: mates taken 5 kb apart
    (no proper pairing inside the default 0..1000 insert range).  Returns (mates1, mates2)."""
    rng = np.random.default_rng(seed)
    lens = np.array([c[1].size for c in contigs], dtype=np.int64)
    probs = lens / lens.sum()
    m1, m2 = [], []
    for i in range(n_pairs):
        cn = int(rng.choice(len(contigs), p=probs)) if len(contigs) > 1 else 0
        g = contigs[cn][1]
        ins = max(read_len + 5, int(rng.normal(ins_mean, ins_sd)))
        far = rng.random() < far_p
        span = ins + (5000 if far else 0)
        pos = int(rng.integers(0, max(1, g.size - span - 10)))
        a = g[pos:pos + read_len + 8]
        b_end = pos + span
        b = g[b_end - read_len - 8:b_end]
        a = _mutate(rng, a, sub, indel_p, max_indel)[:read_len]
        b = _mutate(rng, b, sub, indel_p, max_indel)[-read_len:]
        r1, r2 = a, revcomp(b)
        if rng.random() < 0.5:
            r1, r2 = revcomp(b), a
        j = rng.random()
        if j < junk_p:
            which = int(rng.integers(0, 2))
            junk = BASES[rng.integers(0, 4, size=read_len, dtype=np.uint8)]
            if which == 0:
                r1 = junk
            else:
                r2 = junk
        m1.append((f"p{i}/1", np.ascontiguousarray(r1)))
        m2.append((f"p{i}/2", np.ascontiguousarray(r2)))
    return m1, m2


def write_fasta(path: str, records, width: int = 0):
    with open(path, "wb") as f:
        for name, seq in records:
            f.write(b">" + name.encode() + b"\n")
            b = seq.tobytes() if isinstance(seq, np.ndarray) else bytes(seq)
            if width:
                for k in range(0, len(b), width):
                    f.write(b[k:k + width] + b"\n")
            else:
                f.write(b + b"\n")


CONFIGS = {
    # name: (genome kwargs, read kwargs)
    "c1": dict(genome=dict(total_len=10_000_000, n_contigs=1, seed=1),
               reads=dict(n_reads=100_000, read_len=50, seed=2, sub=0.02)),
    "c1_small": dict(genome=dict(total_len=200_000, n_contigs=3, seed=11),
                     reads=dict(n_reads=2_000, read_len=50, seed=12, sub=0.02)),
    "c1_repeat": dict(genome=dict(total_len=300_000, n_contigs=2, seed=13, repeat_unit=3000,
                                  repeat_copies=40, repeat_div=0.03),
                      reads=dict(n_reads=2_000, read_len=50, seed=14, sub=0.02)),
    "c2_small": dict(genome=dict(total_len=200_000, n_contigs=2, seed=17, n_frac=0.002),
                     reads=dict(n_reads=1_500, read_len=36, seed=18, snp_p=0.3, col_err=0.03), colour=True),
    "c3_small": dict(genome=dict(total_len=300_000, n_contigs=3, seed=19, n_frac=0.002, repeat_unit=1500,
                                 repeat_copies=30, repeat_div=0.02),
                     reads=dict(n_pairs=600, read_len=100, seed=20, sub=0.02, indel_p=0.1), paired=True),
    # BASELINE.json configs[3]: short reads vs a miRNA-like database of many 22 bp contigs
    "c4_small": dict(genome=dict(total_len=22 * 300, n_contigs=300, seed=21),
                     reads=dict(n_reads=800, read_len=22, seed=22, sub=0.03)),
    "c2p_small": dict(genome=dict(total_len=300_000, n_contigs=3, seed=23, repeat_unit=1200, repeat_copies=20,
                                  repeat_div=0.02),
                      reads=dict(n_pairs=500, read_len=40, seed=24, sub=0.01, ins_mean=250.0, ins_sd=25.0),
                      paired=True, colour=True),
    "c5_small": dict(genome=dict(total_len=200_000, n_contigs=1, seed=15),
                     reads=dict(n_reads=500, read_len=75, seed=16, sub=0.04, indel_p=0.5, max_indel=5)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=sorted(CONFIGS))
    ap.add_argument("outdir")
    ap.add_argument("--n-reads", type=int, default=None)
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    os.makedirs(a.outdir, exist_ok=True)
    contigs = make_genome(**cfg["genome"])
    rk = dict(cfg["reads"])
    if a.n_reads:
        rk["n_reads"] = a.n_reads
    rk.pop("colour", None)
    reads = (simulate_cs_reads if cfg.get("colour") else simulate_reads)(contigs, **rk)
    write_fasta(os.path.join(a.outdir, "genome.fa"), contigs, width=80)
    write_fasta(os.path.join(a.outdir, "reads.fa"), reads)


if __name__ == "__main__":
    main()
