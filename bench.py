#!/usr/bin/env python
"""bench.py -- the gmapper hot path on B200: reads/s mapped (and vector-SW GCUPS) on synthetic inputs.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA)
    python bench.py --impl reference --gpus N --steps K ...  # the reference gmapper on the host cores

A step is one pass of the whole hot path (seed scan -> sw_vector -> pass-1 replay -> sw_full_{ls,cs})
over one batch of simulated reads against the HBM-resident index.  Default workload: BASELINE.json
configs[1] (colour space, 36-colour reads vs an iid 100 Mb genome); --workload c1 = configs[0].
Under torchrun every rank maps its own batch (weak scaling, no collective on the data path).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

# the host stages of a step (OpenMP inside the library) share a rank's few cores with the other contexts of the rank:
# idle OpenMP workers must sleep, not spin (libgomp reads this when it is loaded, i.e. before torch is imported)
_USER_WAIT_POLICY = os.environ.get("OMP_WAIT_POLICY")
os.environ.setdefault("OMP_WAIT_POLICY", "passive")


def child_env():
    """environment of the gmapper binaries this script starts (the reference and the drop-in): the caller's own, without
    the wait policy set above for this process"""
    env = dict(os.environ)
    if _USER_WAIT_POLICY is None:
        env.pop("OMP_WAIT_POLICY", None)
    return env


import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

CMPL = np.array([3, 2, 1, 0], dtype=np.uint8)
# algorithmic integer operations per band cell of the full SW: 3 states / 7 candidates in
# letter space, 12 states / 88 candidates in colour space, ~2 operations (add + compare-select) per candidate
FULL_INSTR_PER_CELL = {"ls": 30.0, "cs": 160.0}


class Workload:
    """Seeded synthetic genome + simulated reads of one BASELINE.json config."""

    def __init__(self, key, desc, colour, read_len, contig_len, n_contigs, seed_genome, default_reads, binary, args,
                 paired=False, n_frac=0.0, opts=None, seeds_weight=0, mirna=False, anchor_width=8, gap_open=None,
                 list_cutoff=None, kind=None, cpu_sample=200_000):
        self.key, self.desc, self.colour, self.read_len = key, desc, colour, read_len
        self.contig_len, self.n_contigs, self.seed_genome = contig_len, n_contigs, seed_genome
        self.genome_len = contig_len * n_contigs
        self.default_reads, self.binary, self.args = default_reads, binary, args
        self.paired, self.n_frac = paired, n_frac
        # option set of the gmapper command line `args` as MapParams overrides (tests/mapcases.py has the same table)
        self.opts, self.seeds_weight, self.mirna = dict(opts or {}), seeds_weight, mirna
        self.anchor_width, self.gap_open, self.fixed_cutoff = anchor_width, gap_open, list_cutoff
        self.kind = kind or key
        self.cpu_sample = cpu_sample
        self._genome = None

    def load_args(self):
        """the options of a run that loads the projection with -L: the seeds come with the file (gmapper.c:2372)"""
        out, skip = [], False
        for x in self.args:
            if skip:
                skip = False
            elif x == "-s":
                skip = True
            else:
                out.append(x)
        return out

    def seeds(self):
        from shrimp_b200 import seeds as S
        return S.load_default_mirna_seeds() if self.mirna else S.load_default_seeds(self.seeds_weight)

    def scores(self):
        import shrimp_b200
        from shrimp_b200.api import Scores
        sc = shrimp_b200.CS_DEFAULT_SCORES if self.colour else shrimp_b200.LS_DEFAULT_SCORES
        if self.gap_open is not None:
            sc = Scores(sc.match, sc.mismatch, self.gap_open, sc.a_gap_ext, self.gap_open, sc.b_gap_ext, sc.crossover)
        return sc

    def map_params(self):
        from shrimp_b200.api import MapParams, auto_list_cutoff
        cutoff = self.fixed_cutoff if self.fixed_cutoff is not None else auto_list_cutoff(
            self.genome_len, 12 if self.mirna else max(s.weight for s in self.seeds()))
        kw = dict(list_cutoff=cutoff, compute_mapping_qualities="--no-mapping-qualities" not in self.args,
                  match_mode=4 if self.paired else 2)
        kw.update(self.opts)
        return MapParams(**kw)

    def resize(self, genome_mb: int):
        self.contig_len = genome_mb * 1_000_000 // self.n_contigs
        self.genome_len = self.contig_len * self.n_contigs
        self.desc += f" [genome scaled to {genome_mb} Mb]"
        self._genome = None

    def genome(self):
        if self._genome is None:
            rng = np.random.default_rng(self.seed_genome)
            g = rng.integers(0, 4, size=self.genome_len, dtype=np.uint8)
            if self.n_frac > 0:     # runs of N (code 15), 1 kb each
                n_runs = int(self.genome_len * self.n_frac / 1000)
                for p0 in rng.integers(0, self.genome_len - 1000, size=n_runs):
                    g[p0:p0 + 1000] = 15
            self._genome = g
        return self._genome

    def packed_contigs(self):
        """4 bits per base, 8 per uint32 (util.h:41-42), one array per contig"""
        out = []
        for c in self.contigs():
            n = c.size
            pad = (-n) % 8
            cc = np.concatenate([c, np.zeros(pad, dtype=np.uint8)]) if pad else c
            out.append(np.ascontiguousarray(cc[0::2] | (cc[1::2] << 4)).view(np.uint32))
        return out

    def contig_names(self):
        return [f"contig{i}" for i in range(self.n_contigs)]

    def contigs(self):
        g = self.genome()
        return [g[i * self.contig_len:(i + 1) * self.contig_len] for i in range(self.n_contigs)]

    def pairs(self, n_pairs, seed):
        """opp-in pairs, insert N(300, 30), 2 % substitutions; mates interleaved: rows 2k, 2k+1"""
        g, rl = self.genome(), self.read_len
        rr = np.random.default_rng(seed)
        ins = np.maximum(rl + 5, rr.normal(300.0, 30.0, size=n_pairs).astype(np.int64))
        cn = rr.integers(0, self.n_contigs, size=n_pairs)
        pos = cn * self.contig_len + rr.integers(0, self.contig_len - 400 - rl, size=n_pairs)
        idx = np.arange(rl)[None, :]
        a = g[pos[:, None] + idx].copy()
        b = g[(pos + ins - rl)[:, None] + idx].copy()
        for m in (a, b):
            sub = (rr.random(m.shape) < 0.02) & (m < 4)
            m[sub] = (m[sub] + rr.integers(1, 4, size=int(sub.sum()))) % 4
        cm = np.arange(16, dtype=np.uint8)
        cm[:4] = CMPL
        brc = cm[b][:, ::-1]
        flip = rr.random(n_pairs) < 0.5
        r1 = np.where(flip[:, None], brc, a)
        r2 = np.where(flip[:, None], a, brc)
        out = np.empty((2 * n_pairs, rl), dtype=np.uint8)
        out[0::2] = r1
        out[1::2] = r2
        return out, None

    def reads(self, n_reads, seed):
        """-> (codes [n, read_len] uint8 (letters, or colours in colour space), initbp or None)"""
        if self.paired:
            return self.pairs(n_reads // 2, seed)
        g, rl = self.genome(), self.read_len
        rr = np.random.default_rng(seed)
        cn = rr.integers(0, self.n_contigs, size=n_reads)
        room = self.contig_len - rl - (6 if self.kind == "c5" else 0)
        pos = cn * self.contig_len + (rr.integers(0, room, size=n_reads) if room > 0 else 0)
        if self.kind == "c5":    # C5: 4 % substitutions, half of the reads with one indel of 1..5 bases
            frag = g[pos[:, None] + np.arange(rl + 6)[None, :]].copy()
            sub = rr.random(frag.shape) < 0.04
            frag[sub] = (frag[sub] + rr.integers(1, 4, size=int(sub.sum()))) % 4
            out = frag[:, :rl].copy()
            for i in np.nonzero(rr.random(n_reads) < 0.5)[0]:
                k, at = int(rr.integers(1, 6)), int(rr.integers(5, rl - 10))
                if rr.random() < 0.5:    # the read skips k genome bases
                    out[i, at:] = frag[i, at + k:at + k + rl - at]
                else:                    # k extra bases in the read
                    out[i, at:at + k] = rr.integers(0, 4, size=k)
                    out[i, at + k:] = frag[i, at:rl - k]
            frag = out
        else:
            frag = g[pos[:, None] + np.arange(rl)[None, :]].copy()
        if self.kind == "c5":
            pass
        elif self.kind == "c4":  # C4: one substitution in half of the reads
            m = np.nonzero(rr.random(n_reads) < 0.5)[0]
            p = rr.integers(0, rl, size=m.size)
            frag[m, p] = (frag[m, p] + rr.integers(1, 4, size=m.size)) % 4
        elif not self.colour:      # C1: 2 % substitutions
            sub = rr.random(frag.shape) < 0.02
            frag[sub] = (frag[sub] + rr.integers(1, 4, size=int(sub.sum()))) % 4
        else:                    # C2: one SNP in 30 % of the reads
            snp = np.nonzero(rr.random(n_reads) < 0.3)[0]
            p = rr.integers(0, rl, size=snp.size)
            frag[snp, p] = (frag[snp, p] + rr.integers(1, 4, size=snp.size)) % 4
        rc = rr.random(n_reads) < 0.5
        frag[rc] = CMPL[frag[rc]][:, ::-1]
        if not self.colour:
            return frag, None
        prev = np.concatenate([np.full((n_reads, 1), 3, dtype=np.uint8), frag[:, :-1]], axis=1)  # primer base T
        col = frag ^ prev
        err = rr.random(col.shape) < 0.03    # 3 % colour errors
        col[err] = (col[err] + rr.integers(1, 4, size=int(err.sum()))) % 4
        return col, np.full(n_reads, 3, dtype=np.int8)

    def write_reads_fasta(self, path, codes):
        lut = np.frombuffer(b"0123............" if self.colour else b"ACGTNNNNNNNNNNNN", dtype=np.uint8)
        body = lut[codes]
        if self.paired:     # path.1 / path.2, mates p<k>/1 and p<k>/2
            with open(path + ".1", "wb") as f1, open(path + ".2", "wb") as f2:
                for k in range(body.shape[0] // 2):
                    f1.write(b">p%d/1\n" % k + body[2 * k].tobytes() + b"\n")
                    f2.write(b">p%d/2\n" % k + body[2 * k + 1].tobytes() + b"\n")
            return
        with open(path, "wb") as f:
            for i in range(body.shape[0]):
                f.write(b">r%d\n" % i + (b"T" if self.colour else b"") + body[i].tobytes() + b"\n")

    def write_genome_fasta(self, path):
        lut = np.frombuffer(b"ACGTNNNNNNNNNNNN", dtype=np.uint8)
        with open(path, "wb") as f:
            for nm, c in zip(self.contig_names(), self.contigs()):
                f.write(b">" + nm.encode() + b"\n" + lut[c].tobytes() + b"\n")


WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on that fits one GPU, with gmapper-cs's default
    # options (mapping qualities on: post_sw rescoring of every alignment); c2nomq = the same with
    # --no-mapping-qualities
    "c2": Workload("c2", "C2 colour-space: 36-colour SOLiD reads (SNP in 30%, 3% colour errors) vs iid 100 Mb genome "
                   "in 10 contigs, 3 default seeds w12, sw_full_cs crossovers, mapping qualities (post_sw) on as by "
                   "default; one step = one batch of the 10 M-read job", True, 36, 10_000_000, 10, 3, 1_000_000,
                   "gmapper-cs", []),
    "c2nomq": Workload("c2nomq", "C2 colour-space without mapping qualities: 36-colour SOLiD reads (SNP in 30%, 3% "
                       "colour errors) vs iid 100 Mb genome in 10 contigs, 3 default seeds w12, sw_full_cs crossovers, "
                       "--no-mapping-qualities; one step = one batch of the 10 M-read job", True, 36, 10_000_000, 10, 3,
                       1_000_000, "gmapper-cs", ["--no-mapping-qualities"]),
    # BASELINE.json configs[2]: paired-end letter space; the genome is scaled by --genome-mb (default 300 Mb; 3000 =
    # the hg18-sized configuration, every read strand then goes through the CTA-per-strand scan kernel)
    "c3": Workload("c3", "C3 paired-end letter-space: 2x100bp opp-in pairs (insert N(300,30), 2% subs) vs iid genome "
                   "in 24 contigs with 1% N runs, 3 default seeds w12, -p opp-in -I 0,1000; value counts reads "
                   "(2 per pair)", False, 100, 12_500_000, 24, 5, 400_000, "gmapper-ls",
                   ["-p", "opp-in", "-I", "0,1000"], paired=True, n_frac=0.01),
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    "c1": Workload("c1", "C1 letter-space: 100k x 50bp reads (2% subs) vs iid 10 Mb genome, 3 default seeds w12",
                   False, 50, 10_000_000, 1, 1, 100_000, "gmapper-ls", []),
    # BASELINE.json configs[3]: 22 bp reads vs a miRNA-like database of 2,000 x 22 bp contigs (short targets, many
    # windows), with gmapper-ls's default options and with -M mirna (gmapper.c:1498-1515: hashed 5-seed set, gapless
    # pass 1, gap opens -255, no window cache, one seed match, window 100 %, local full SW, no mapping qualities)
    "c4": Workload("c4", "C4 miRNA: 22bp reads (one substitution in half of them) vs 2,000 x 22bp contigs, default "
                   "options", False, 22, 22, 2000, 7, 1_000_000, "gmapper-ls", [], kind="c4"),
    "c4mirna": Workload("c4mirna", "C4 miRNA with -M mirna: 22bp reads vs 2,000 x 22bp contigs, hashed 5-seed set, "
                        "gapless pass 1, local full SW", False, 22, 22, 2000, 7, 1_000_000, "gmapper-ls", ["-M", "mirna"],
                        opts=dict(match_mode=1, window_len=100.0, gapless=True, hash_filter_calls=False, Gflag=False,
                                  compute_mapping_qualities=False), mirna=True, anchor_width=0, gap_open=-255, kind="c4"),
    # BASELINE.json configs[4]: the "overly sensitive" mode (README:481-534, letter-space subset): four weight-11
    # seeds, one seed match, wide windows, threshold band (-a -1), no window cache (-Z), no index trimming (-V):
    # about a thousand sw_vector calls per read in the reference -- the DPX kernel's real test
    "c5": Workload("c5", "C5 overly sensitive: 75bp reads (4% subs, one indel of 1-5 bp in half of them) vs iid 10 Mb "
                   "genome, -s w11 -n 1 -w 150% -r 50% -l 40% -Z -h 60% -a -1 -V", False, 75, 10_000_000, 1, 1, 200_000,
                   "gmapper-ls", ["-s", "w11", "-n", "1", "-w", "150%", "-r", "50%", "-l", "40%", "-Z", "-h", "60%", "-a",
                                  "-1", "-V"],
                   opts=dict(match_mode=1, window_len=150.0, window_gen_threshold=50.0, window_overlap=40.0,
                             hash_filter_calls=False, sw_full_threshold=60.0), seeds_weight=11, anchor_width=-1,
                   list_cutoff=0xFFFFFFFF, kind="c5", cpu_sample=4_000),
}


def make_workload(n_reads: int, seed_genome: int = 1, seed_reads: int = 2):
    """C1 inputs as (genome codes, read codes, None, None) -- kept for tests/test_gpu_pipeline.py"""
    w = WORKLOADS["c1"]
    codes, _ = w.reads(n_reads, seed_reads)
    return w.genome(), codes, None, None


def pack_rows(codes: np.ndarray) -> np.ndarray:
    n, rl = codes.shape
    stride = (rl + 7) // 8
    buf = np.zeros((n, stride * 8), dtype=np.uint32)
    buf[:, :rl] = codes
    sh = (4 * np.arange(8, dtype=np.uint32))[None, None, :]
    return np.bitwise_or.reduce(buf.reshape(n, stride, 8) << sh, axis=2).astype(np.uint32)


def write_fasta(path, names, codes):
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    with open(path, "wb") as f:
        for nm, row in zip(names, codes):
            f.write(b">" + nm.encode() + b"\n" + lut[row].tobytes() + b"\n")


# -------------------------------------------------------------------------------------------------
# clocks
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.samples = []
        self.stop = threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.idx)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=3)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[1]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][2]), "reasons": sorted(reasons),
                "samples": len(self.samples)}


# -------------------------------------------------------------------------------------------------
# the reference's own CPU implementation (oracle/_ref/gmapper-{ls,cs}, compiled from /root/reference)
# -------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def run_reference(w: Workload, workdir: str, n_threads: int, reads_fa: str, prefix: str):
    """returns (seconds of 'Read Mapping Time', vector GCUPS aggregate, wall seconds)"""
    rd = ["-1", reads_fa + ".1", "-2", reads_fa + ".2"] if w.paired else [reads_fa]
    cmd = [os.path.join(REF_DIR, w.binary), "-N", str(n_threads), *w.load_args(), "-L", prefix, *rd]
    t0 = time.time()
    r = subprocess.run(cmd, cwd=workdir, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=child_env())
    wall = time.time() - t0
    if r.returncode != 0:
        raise RuntimeError("reference gmapper failed: " + r.stderr[-500:])
    m = re.search(r"Read Mapping Time:\s+([0-9.]+) seconds", r.stderr)
    map_s = float(m.group(1))
    cells = re.search(r"Vector Smith-Waterman:.*?Cells Computed:\s+([0-9.]+) million", r.stderr, re.S)
    vsec = re.search(r"Vector Smith-Waterman:\s+Run-time:\s+([0-9.]+) seconds", r.stderr)
    gcups = None
    if cells and vsec and float(vsec.group(1)) > 0:
        gcups = float(cells.group(1)) * 1e6 / (float(vsec.group(1)) / n_threads) / 1e9
    return map_s, gcups, wall


def reference_setup(w: Workload, workdir: str, reads_codes, ctx=None):
    """reads.fa + the projection files `proj.*` that the timed runs load with -L (the reference's
    "Read Mapping Time" excludes loading).  With a GPU context the projection held in HBM is saved in the
    -S format (shrimp_gpu_projection_save; tests/test_gpu_index.py checks the files byte for byte against
    `gmapper -S`); without one the reference projects the genome itself (minutes at 100 Mb)."""
    w.write_reads_fasta(os.path.join(workdir, "reads.fa"), reads_codes)
    if ctx is not None:
        ctx.save_projection(os.path.join(workdir, "proj"), w.contig_names())
        return "projection saved from HBM in the -S format"
    w.write_genome_fasta(os.path.join(workdir, "genome.fa"))
    r = subprocess.run([os.path.join(REF_DIR, w.binary), *(w.args if w.paired else []), "-S", "proj", "genome.fa"], cwd=workdir,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=child_env())
    if r.returncode != 0:
        raise RuntimeError("reference gmapper -S failed: " + r.stderr[-500:])
    return "projection built by gmapper -S"


def run_oracle_port(w: Workload, n_sample: int):
    """CPU baseline when the compiled reference (oracle/_ref) is not on the box: the C restatement of the path
    (oracle/, one host thread) on a small sample of the workload.  Returns (reads/s, sample description)."""
    from oracle import pipeline as op
    from shrimp_b200 import seeds as S
    import shrimp_b200
    n = min(n_sample, 20_000) & ~1
    codes, initbp = w.reads(n, 1000)
    scores = w.scores()
    g = op.Genome(w.contigs(), w.colour)
    ix = op.Index(g, w.seeds())
    opts = op.MapOptions(scores=scores, colour_space=w.colour, list_cutoff=op.auto_list_cutoff(w.genome_len, 12),
                         compute_mapping_qualities="--no-mapping-qualities" not in w.args,
                         match_mode=4 if w.paired else 2)
    rl = np.full(n, w.read_len, dtype=np.int32)
    t0 = time.time()
    if w.paired:
        op.map_pairs(g, ix, opts, pack_rows(codes), rl)
    else:
        op.map_reads(g, ix, opts, pack_rows(codes), rl, initbp=initbp)
    dt = time.time() - t0
    return n / dt, f"{n} reads of the same workload through the C oracle port (oracle/shrimp_oracle.c), 1 host thread"


def bench_config(w: Workload, n_reads: int, world: int):
    """the `config` of the JSON line, the same for both arms"""
    return {"workload": w.desc, "reads_per_gpu_per_step": n_reads, "read_len": w.read_len, "genome_len": w.genome_len,
            "parallelism": f"read-sharded x{world}, replicated index",
            "l2": "L2 flushed (256 MB write) before every timed step; index > L2"}


def build_context(w: Workload, device: int):
    import shrimp_b200
    from shrimp_b200 import seeds as S
    ctx = shrimp_b200.GpuContext(device)
    scores = w.scores()
    seeds = w.seeds()
    # dblen/qrlen as gmapper sets them up (longest_read_len 1000, window 140 % / 150 %)
    ctx.sw_setup(1500, 1000, scores, use_colours=w.colour, anchor_width=w.anchor_width)
    t0 = time.time()
    ctx.load_genome(w.packed_contigs(), [w.contig_len] * w.n_contigs, colour_space=w.colour)
    ctx.build_index(seeds, hflag=w.mirna)
    return ctx, scores, seeds, time.time() - t0


def measure(a, w, n_reads, rank, world, local_rank, ncores, compact=False):
    """One workload on this rank's GPU: device-resident value, end to end through the C ABI, per-stage rooflines, the
    reference on the host cores (rank 0, one GPU) and the drop-in binary FASTA -> SAM.  Returns the JSON line (rank 0).
    compact: a short leg for the `workloads` block of the default run."""
    ref_bin = os.path.join(REF_DIR, w.binary)
    import torch
    import torch.distributed as dist

    from shrimp_b200.api import MapParams, auto_list_cutoff

    torch.cuda.set_device(local_rank)

    codes, initbp_np = w.reads(n_reads, 2 + rank)
    # pinned host buffers for the end-to-end leg
    packed = torch.from_numpy(pack_rows(codes)).pin_memory().numpy()
    read_len = torch.full((n_reads,), w.read_len, dtype=torch.int32).pin_memory().numpy()
    initbp = torch.from_numpy(initbp_np).pin_memory().numpy() if initbp_np is not None else None

    ctx, scores, seeds, index_s = build_context(w, local_rank)
    # torchrun exports OMP_NUM_THREADS=1: give every rank its share of the host cores for the host stages
    from shrimp_b200.api import set_host_threads
    set_host_threads(a.host_threads or max(1, ncores // max(1, world)))
    params = w.map_params()

    def map_host():
        if w.paired:
            return ctx.map_pairs(params, scores, packed, read_len, reuse_buffers=True)
        return ctx.map_reads(params, scores, packed, read_len, initbp=initbp, reuse_buffers=True)

    def map_dev():
        if w.paired:
            return ctx.map_pairs_resident(params, scores)
        return ctx.map_resident(params, scores)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # first call uploads the batch and leaves it resident; also the e2e path's warm-up
    res = map_host()
    if w.paired:
        n_mapped = int(2 * (res.n_pairs_per_pair > 0).sum() +
                       ((res.n_unpaired_per_read.reshape(-1, 2) > 0) & (res.n_pairs_per_pair[:, None] == 0)).sum())
    else:
        n_mapped = int((res.n_hits_per_read > 0).sum())
    for _ in range(a.warmup):
        map_dev()

    # ---- device-resident timed region: exactly K steps -------------------------------------------
    dpx_peak = ctx.dpx_peak()
    fp64_peak_live = ctx.fp64_peak()
    ctx.stage_times_reset()
    with ClockSampler(local_rank) as clk:
        barrier()
        launches0 = ctx.launch_count()
        step_ms = []
        for _ in range(a.steps):
            ctx.flush_l2()
            ctx.event_record(0)
            st = map_dev()
            ctx.event_record(1)
            step_ms.append(ctx.event_elapsed_ms())
        barrier()
        launches = ctx.launch_count() - launches0
    clocks = clk.summary()
    stage = ctx.stage_times()
    total_ms = float(sum(step_ms))
    from shrimp_b200 import shard
    total_ms_max = shard.max_over_ranks(total_ms, world, device="cuda")   # the slowest rank's timed region
    value = world * n_reads * a.steps / (total_ms_max * 1e-3)

    # ---- end to end through the public API: pinned host buffers in, hits out -----------------------
    # A few host threads per GPU (--e2e-threads), one context each (own stream and chunk buffers, shared index), the way
    # gmapper's -N threads share the projection: the host half of a step (read_pass2 on the host cores, D2H)
    # overlaps the device half of the other thread's step.  Every step still uploads its reads and
    # downloads its records; K steps in total.
    import shrimp_b200
    n_host = max(1, min(a.e2e_threads or (4 if world == 1 else 3), a.steps))
    workers = [map_host]
    extra_ctx = []
    for i in range(1, n_host):
        cx = shrimp_b200.GpuContext(local_rank)
        cx.sw_setup(1400, 1000, scores, use_colours=w.colour)
        cx.share_genome_from(ctx)
        codes_i, initbp_i_np = w.reads(n_reads, 1001 + i + 10 * rank)
        packed_i = torch.from_numpy(pack_rows(codes_i)).pin_memory().numpy()
        initbp_i = torch.from_numpy(initbp_i_np).pin_memory().numpy() if initbp_i_np is not None else None

        def map_host_i(cx=cx, packed_i=packed_i, initbp_i=initbp_i):
            if w.paired:
                return cx.map_pairs(params, scores, packed_i, read_len, reuse_buffers=True)
            return cx.map_reads(params, scores, packed_i, read_len, initbp=initbp_i, reuse_buffers=True)

        map_host_i()   # warm-up of this context's buffers
        workers.append(map_host_i)
        extra_ctx.append(cx)
    todo = list(range(a.steps))
    lock = threading.Lock()

    def work(fn):
        while True:
            with lock:
                if not todo:
                    return
                todo.pop()
            fn()

    barrier()
    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(fn,)) for fn in workers]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d, d2h = ctx.last_transfer_bytes()
    e2e_val = world * n_reads * a.steps / shard.max_over_ranks(e2e_s, world, device="cuda")

    if rank != 0:
        for cx in extra_ctx:
            cx.close()
        ctx.close()
        return None

    # ---- roofline of the dominant kernels ----------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    vec_ms = stage["sw_vector"][0] / a.steps
    scan_ms = stage["seed_scan"][0] / a.steps
    full_ms = stage["sw_full"][0] / a.steps
    # sw_vector: every window the device scores (a superset of the reference's calls), 4 integer-pipe
    # instructions per cell (8 per packed pair of cells), against the measured VIADDMNMX.S16x2 peak
    dev_cells = st["device_vector_cells"]
    vec_gcups = dev_cells / (vec_ms * 1e-3) / 1e9 if vec_ms > 0 else 0.0
    vec_ginstr = vec_gcups * 4.0
    # seed scan: algorithmic bytes = packed read (both strands) + 2 table words per k-mer + 4 B per
    # gathered list entry + 48 B per hit written
    mkp = 1 if w.colour else 0
    kmers = 2 * n_reads * sum(max(0, w.read_len - s.span + 1 - mkp) for s in seeds)
    scan_bytes = 2 * n_reads * 4 * packed.shape[1] + kmers * 8 + st["list_entries"] * 4 + st["hits"] * 48
    scan_gbs = scan_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    # full SW: band cells x integer-pipe instructions per cell (counted from SASS)
    full_ipc = FULL_INSTR_PER_CELL["cs" if w.colour else "ls"]
    full_gcells = st["full_cells"] / (full_ms * 1e-3) / 1e9 if full_ms > 0 else 0.0
    dominant = max(stage.items(), key=lambda kv: kv[1][0])[0]
    roofs = {
        "sw_vector": {"kernel": "sw_vector_kernel", "bound": "int-dpx", "achieved": vec_ginstr, "peak": dpx_peak,
                      "unit": "G thread-instr/s", "frac": vec_ginstr / dpx_peak if dpx_peak else None,
                      "traffic": None, "gcups": vec_gcups, "ms_per_launch": vec_ms,
                      "peak_source": "measured live: register-resident VIADDMNMX.S16x2 chains (shrimp_gpu_dpx_peak)"},
        "seed_scan": {"kernel": "scan_cta_kernel" if st["scan_big_strands"] > n_reads else "scan_kernel", "bound": "hbm", "achieved": scan_gbs, "peak": hbm_peak, "unit": "GB/s",
                      "frac": scan_gbs / hbm_peak, "traffic": None, "ms_per_launch": scan_ms, "peak_source": peak_src,
                      "algorithmic_bytes_per_launch": scan_bytes},
        "sw_full": {"kernel": "sw_full_cs_quad_kernel" if w.colour else "sw_full_ls_ring_kernel", "bound": "int-alu",
                    "achieved": full_gcells * full_ipc, "peak": 2.0 * dpx_peak, "unit": "G integer ops/s",
                    "frac": full_gcells * full_ipc / (2.0 * dpx_peak) if dpx_peak else None, "traffic": None,
                    "gcells_per_s": full_gcells, "ops_per_cell": full_ipc, "ms_per_step": full_ms,
                    "peak_source": "integer issue peak = 128 lanes per clock and SM = 2 x the measured half-rate "
                                   "VIADDMNMX.S16x2 peak (shrimp_gpu_dpx_peak)"},
    }
    post_ms = stage.get("post_sw", (0.0, 0))[0] / a.steps
    if post_ms > 0:
        # post_sw: per aligned column 16 nodes x (3 exp + 2 log: forward, backward, posterior), 11 / 20 FP64
        # instructions each in the libm transcription (glibc_math.cuh) -- round 1's count, kept so that the fractions
        # of the rounds compare (the quad kernel of round 2 evaluates 48 exp + 8 log per column, the half-warp kernel
        # evaluated 48 + 32); against the FP64 issue rate measured live (shrimp_gpu_fp64_peak)
        fp64_instr = st["post_sw_columns"] * 16.0 * (3 * 11 + 2 * 20)
        fp64_peak = fp64_peak_live
        roofs["post_sw"] = {"kernel": "post_sw_quad_kernel", "bound": "fp64", "achieved": fp64_instr / (post_ms * 1e-3) / 1e9,
                            "peak": fp64_peak, "unit": "G FP64 instr/s",
                            "frac": fp64_instr / (post_ms * 1e-3) / 1e9 / fp64_peak, "traffic": None,
                            "columns_per_s": st["post_sw_columns"] / (post_ms * 1e-3), "ms_per_step": post_ms,
                            "peak_source": "measured live: register-resident DFMA chains (shrimp_gpu_fp64_peak)"}
    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture per launch; only for the
    # configuration the capture was taken on (profiles/r02b_ncu_full_scan_kernel_raw.csv, r02b_ncu_full_post_sw_quad_raw.csv: C2, 1 M reads per launch)
    if w.key in ("c2", "c2nomq") and n_reads == 1_000_000:
        roofs["seed_scan"]["traffic"] = 14.26e9
        roofs["seed_scan"]["traffic_source"] = "profiles/r02b_ncu_full_scan_kernel_raw.csv"
        if "post_sw" in roofs:
            roofs["post_sw"]["traffic"] = 9.15e9
            roofs["post_sw"]["traffic_source"] = "profiles/r02b_ncu_full_post_sw_quad_raw.csv"
    roofline = dict(roofs.get(dominant, roofs["seed_scan"]))
    roofline["dominant_stage"] = dominant
    roofline["other"] = {k: v for k, v in roofs.items() if k != dominant}

    # ---- CPU baseline: the reference binary on this box's cores, bounded sample; and the drop-in binary
    # (integration/_build: the reference's unchanged gmapper.c / output.c / fasta.c objects + the shims + the library)
    # FASTA -> SAM on the same kind of files, each by its own "Read Mapping Time" ---------------------------------
    cpu = None
    e2e_sam = None
    dropin_job = None
    tmp_holder = None
    n_cpu = min(a.cpu_sample, w.cpu_sample) & ~1
    if world == 1 and not a.no_cpu_baseline and not os.path.exists(ref_bin) and w.genome_len <= 400_000_000:
        try:
            v, sample_s = run_oracle_port(w, n_cpu)
            cpu = {"value": v, "unit": "reads/s", "cores": 1, "kind": "port", "sample": sample_s}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": "reads/s", "cores": 1, "kind": "port", "sample": f"failed: {e}"}
    if world == 1 and not a.no_cpu_baseline and os.path.exists(ref_bin) and w.genome_len <= 400_000_000:
        try:
            sample, _ = w.reads(n_cpu, 1000)
            tmp_holder = tempfile.TemporaryDirectory()
            if True:
                d = tmp_holder.name
                how = reference_setup(w, d, sample, ctx)
                s, gcu, _ = run_reference(w, d, ncores, "reads.fa", "proj")
                cpu = {"value": n_cpu / s, "unit": "reads/s", "cores": ncores, "kind": "reference",
                       "sample": f"{n_cpu} reads of the same workload, oracle/_ref/{w.binary} -N {ncores} "
                                 f"{' '.join(w.args)} -L <projection> (Read Mapping Time, index load excluded); {how}",
                       "sw_vector_gcups": gcu}
                dropin = os.path.join(ROOT, "integration", "_build", w.binary)
                if os.path.exists(dropin):
                    n_sam = (2 * n_reads if compact else 6 * n_reads) & ~1
                    big, _ = w.reads(n_sam, 1001)
                    w.write_reads_fasta(os.path.join(d, "big.fa"), big)
                    dropin_job = (dropin, d, n_sam)
        except Exception as e:  # noqa: BLE001
            cpu = cpu or {"value": None, "unit": "reads/s", "cores": ncores, "kind": "reference", "sample": f"failed: {e}"}

    line = {
        "metric": "reads_per_sec_mapped", "value": value, "unit": "reads/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": total_ms_max / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": bench_config(w, n_reads, world),
        "sw_vector_gcups": vec_gcups, "sw_full_mcells_per_s": full_gcells * 1e3,
        "reads_mapped_frac": n_mapped / n_reads,
        "index_build_s": index_s,
        "stage_ms_per_step": {k: v[0] / a.steps for k, v in stage.items()},
        "pipeline_stats": st,
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "host_threads_per_gpu": n_host},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e_sam": e2e_sam,
    }
    if compact:   # the short form that goes into the `workloads` block of the default run
        line = {"value": value, "unit": "reads/s", "ms_per_step": total_ms_max / a.steps, "steps": a.steps,
                "reads_per_step": n_reads, "genome_len": w.genome_len, "reads_mapped_frac": n_mapped / n_reads,
                "e2e": line["e2e"], "e2e_sam": e2e_sam, "cpu_baseline": cpu, "index_build_s": index_s,
                "stage_ms_per_step": line["stage_ms_per_step"], "sw_vector_gcups": vec_gcups,
                "sw_full_mcells_per_s": full_gcells * 1e3,
                "roofline": {k: {kk: v.get(kk) for kk in ("kernel", "bound", "achieved", "peak", "unit", "frac")}
                             for k, v in roofs.items()},
                "dominant_stage": dominant, "gpu_launches": launches, "config": w.desc}
    for cx in extra_ctx:
        cx.close()
    ctx.close()
    if dropin_job is not None:
        # the drop-in binary runs with the GPU to itself, as it would in production: this process's contexts are closed
        dropin, d, n_sam = dropin_job
        th, ck = min(ncores, 16), max(2000, min(50_000, n_sam // 32)) & ~1
        rd = ["-1", "big.fa.1", "-2", "big.fa.2"] if w.paired else ["big.fa"]
        # two runs, the faster one counts, both are reported: the binary's time is host-bound (sixteen threads on
        # sixteen cores, a serial reader) and moves by +-25 % from run to run on the same box
        runs, detail = [], []
        for verbose in (False, True):   # the second run also times its reader (two rdtsc per entry) and its threads
            try:
                r = subprocess.run([dropin, "-N", str(th), "-K", str(ck), *w.load_args(), "-L", "proj", *rd], cwd=d,
                                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=900,
                                   env=dict(child_env(), **({"SHRIMP_B200_VERBOSE": "1"} if verbose else {})))
            except subprocess.TimeoutExpired:
                r = subprocess.CompletedProcess([dropin], 124, "", "timed out after 900 s")
                m = None
                break
            m = re.search(r"Read Mapping Time:\s+([0-9.]+) seconds", r.stderr)
            if r.returncode != 0 or not m or float(m.group(1)) <= 0:
                break
            runs.append(n_sam / float(m.group(1)))
            # where the run's time went: the serial reader (inside gmapper.c's critical section), the threads' summed
            # wait for that lock, and one thread's own split
            rd_ns = re.search(r"reader: \d+ entries in ([0-9.]+) s \((\d+) ns each\)", r.stderr)
            wait = re.search(r"Wait Time:\s+([0-9.]+) seconds", r.stderr)
            thr = re.search(r"device calls ([0-9.]+) \(the first ([0-9.]+)\), record rebuild ([0-9.]+), output ([0-9.]+)", r.stderr)
            detail.append({"map_s": float(m.group(1)), "reader_s": float(rd_ns.group(1)) if rd_ns else None,
                           "reader_ns_per_entry": int(rd_ns.group(2)) if rd_ns else None,
                           "threads_wait_s_sum": float(wait.group(1)) if wait else None,
                           "one_thread": {"device_calls_s": float(thr.group(1)), "first_call_s": float(thr.group(2)),
                                          "rebuild_s": float(thr.group(3)), "output_s": float(thr.group(4))} if thr else None})
        # the same link line with the reference's OWN reader and SAM formatter (gmapper-b200-refio: nothing of
        # fast_io.cpp; the strict reading of "input/output and SAM emission unchanged"), one run
        ref_io = None
        refio_bin = os.path.join(os.path.dirname(dropin), "refio", w.binary)
        if runs and os.path.exists(refio_bin):
            try:
                r2 = subprocess.run([refio_bin, "-N", str(th), "-K", str(ck), *w.load_args(), "-L", "proj", *rd], cwd=d,
                                    stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=900, env=child_env())
                m2 = re.search(r"Read Mapping Time:\s+([0-9.]+) seconds", r2.stderr)
            except subprocess.TimeoutExpired:
                r2, m2 = None, None
            if r2 is not None and r2.returncode == 0 and m2 and float(m2.group(1)) > 0:
                ref_io = {"value": n_sam / float(m2.group(1)), "unit": "reads/s",
                          "how": "integration/_build/refio: the reference's unchanged fasta.o / util.o / output.o in "
                                 "place of integration/fast_io.cpp, same options"}
        if runs:
            e2e_sam = {"value": max(runs), "unit": "reads/s", "reads": n_sam, "runs": runs, "runs_detail": detail,
                       "reference_io": ref_io,
                       "reference_value": cpu["value"] if cpu else None,
                       "how": f"integration/_build/{w.binary} -N {th} -K {ck} {' '.join(w.load_args())} -L <projection> "
                              "<reads.fa>: FASTA in, SAM out, the binary's own Read Mapping Time (the reference's clock, "
                              "gmapper.c:3015-3021): the reference's unchanged gmapper.c / genome.c / output.c with the "
                              "mapping shims and integration/fast_io.cpp (reader + SAM formatter, SURVEY 8 f2); the reader "
                              "runs inside gmapper.c's omp critical section (gmapper.c:338) and bounds it, not the device"}
        else:
            e2e_sam = {"value": None, "unit": "reads/s", "how": "failed: " + r.stderr[-300:]}
        line["e2e_sam"] = e2e_sam
    if tmp_holder is not None:
        tmp_holder.cleanup()
    return line


def other_workloads(a, ncores):
    """The other BASELINE.json configs next to the headline one, a few steps each (one GPU): C1, C3 at 300 Mb and at
    hg18 size (100 k reads per step), C4 with default options and with -M mirna, C5."""
    import copy
    out = {}
    legs = [("c1", "c1", 0, 0), ("c3_300mb", "c3", 300, 0), ("c3_3gb", "c3", 3000, 100_000), ("c4", "c4", 0, 0),
            ("c4mirna", "c4mirna", 0, 0), ("c5", "c5", 0, 0)]
    for name, key, mb, n in legs:
        w = copy.copy(WORKLOADS[key])
        w._genome = None
        if mb:
            w.resize(mb)
        b = copy.copy(a)
        b.steps, b.warmup, b.e2e_threads = 3, 3, 2
        try:
            if mb >= 3000:
                import psutil
                if psutil.virtual_memory().available < 40e9:
                    raise RuntimeError("less than 40 GB of host memory free for the 3 Gb genome")
            out[name] = measure(b, w, n or w.default_reads, 0, 1, 0, ncores, compact=True)
        except Exception as e:  # noqa: BLE001
            out[name] = {"value": None, "error": str(e)[-300:]}
        w._genome = None
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU per step (0 = the workload's default)")
    ap.add_argument("--genome-mb", type=int, default=0, help="scale the synthetic genome (c3: default 300, 3000 = hg18 size)")
    ap.add_argument("--cpu-sample", type=int, default=200_000, help="reads in the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-workloads-block", action="store_true",
                    help="skip the compact C1 / C3 / C4 / C5 legs the default C2 run appends")
    ap.add_argument("--host-threads", type=int, default=0, help="OpenMP threads of the library's host stages per call "
                    "(0 = the rank's share of the host cores)")
    ap.add_argument("--e2e-threads", type=int, default=0,
                    help="host threads (one context each) of the end-to-end leg; 0 = 4 on one GPU, 3 per GPU on several")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1
    w = WORKLOADS[a.workload]
    if a.genome_mb or w.key == "c3":
        w.resize(a.genome_mb or 300)
    n_reads = a.reads or w.default_reads
    ref_bin = os.path.join(REF_DIR, w.binary)

    if a.impl == "reference":
        if rank != 0:
            return 0
        if not os.path.exists(ref_bin):   # the compiled reference did not travel: time the oracle port instead
            vals = []
            for _ in range(max(1, min(a.steps, 2))):
                v, sample_s = run_oracle_port(w, a.cpu_sample)
                vals.append(v)
            val = sum(vals) / len(vals)
            print(json.dumps({"impl": "reference", "metric": "reads_per_sec_mapped", "value": val, "unit": "reads/s",
                              "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": None,
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16",
                              "data": "synthetic", "config": {"workload": w.desc},
                              "cpu_baseline": {"value": val, "unit": "reads/s", "cores": 1, "kind": "port",
                                               "sample": sample_s},
                              "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
            return 0
        n_cpu = min(a.cpu_sample, w.cpu_sample) & ~1
        sample, _ = w.reads(n_cpu, 1000)
        # the projection the timed runs load with -L is the reference's own (gmapper -S, minutes at 100 Mb, untimed),
        # kept under /tmp for the other invocations on this box; nothing of this repo's library runs in this arm
        cache = os.path.join(tempfile.gettempdir(), f"shrimp_ref_proj_{w.key}_{w.genome_len}")
        os.makedirs(cache, exist_ok=True)
        if not os.path.exists(os.path.join(cache, "proj.genome")):
            w.write_genome_fasta(os.path.join(cache, "genome.fa"))
            r = subprocess.run([ref_bin, *[x for x in w.args if x in ("-s", "w11", "-M", "mirna")], "-S", "proj.tmp",
                                "genome.fa"], cwd=cache, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=child_env())
            if r.returncode != 0:
                raise RuntimeError("reference gmapper -S failed: " + r.stderr[-500:])
            for f in sorted(os.listdir(cache)):
                if f.startswith("proj.tmp"):
                    os.rename(os.path.join(cache, f), os.path.join(cache, "proj" + f[len("proj.tmp"):]))
            os.remove(os.path.join(cache, "genome.fa"))
        how = "projection built by the reference itself (gmapper -S)"
        with tempfile.TemporaryDirectory() as d:
            w.write_reads_fasta(os.path.join(d, "reads.fa"), sample)
            prefix = os.path.join(cache, "proj")
            for _ in range(min(a.warmup, 1)):
                run_reference(w, d, ncores, "reads.fa", prefix)
            secs, gc = [], []
            for _ in range(a.steps):
                s_, g, _ = run_reference(w, d, ncores, "reads.fa", prefix)
                secs.append(s_)
                gc.append(g)
        tot = sum(secs)
        val = n_cpu * a.steps / tot
        sample_s = (f"{n_cpu} reads of the {w.key.upper()} workload per step, {w.binary} -N {ncores} "
                    f"{' '.join(w.args)} -L <projection> (Read Mapping Time: FASTA parsing and SAM printing included); {how}")
        line = {"impl": "reference", "metric": "reads_per_sec_mapped", "value": val, "unit": "reads/s",
                "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot / a.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16",
                "data": "synthetic", "config": bench_config(w, n_reads, a.gpus),
                "sw_vector_gcups": gc[-1],
                "cpu_baseline": {"value": val, "unit": "reads/s", "cores": ncores, "kind": "reference",
                                 "sample": sample_s},
                "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = measure(a, w, n_reads, rank, world, local_rank, ncores)
    if rank == 0:
        if world == 1 and a.workload == "c2" and not a.no_workloads_block and not a.genome_mb and not a.reads:
            line["workloads"] = other_workloads(a, ncores)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
