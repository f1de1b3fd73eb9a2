"""The N > 1 path on CPU: world_size-2 gloo processes shard a read set by chunks, map their chunks with the
CPU oracle (standing in for the GPU mapper, which is per-rank and collective-free) and rank 0 gathers in chunk
order -- the result must equal the single-process mapping of the same reads."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from shrimp_b200 import shard  # noqa: E402

N_READS, CHUNK = 120, 16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_mapper():
    from mapcases import LsCase
    from oracle import pipeline as op
    case = LsCase("c1_small")
    g = op.Genome(case.contig_codes, False)
    ix = op.Index(g, case.seeds)
    opts = op.MapOptions(scores=case.scores, list_cutoff=case.list_cutoff)

    def map_chunk(a, b):
        hits, nper, _, _ = op.map_reads(g, ix, opts, case.packed[a:b], case.read_len[a:b])
        return [(a + int(h["read_idx"]), int(h["cn"]), int(h["gen_st"]), int(h["sfr"]["genome_start"]),
                 int(h["score_full"])) for h in hits], nper.tolist()
    return map_chunk


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = shard.map_sharded(N_READS, CHUNK, _oracle_mapper(), rank, world)
        slowest = shard.max_over_ranks(10.0 + rank, world)
        if rank == 0:
            torch.save({"res": res, "slowest": slowest}, out_path)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_chunk_plan_round_robin_keeps_order_and_covers_everything():
    plan = shard.chunk_plan(103, 10, 4)
    seen = sorted(c for p in plan for c in p)
    assert [c[0] for c in seen] == list(range(11))
    assert seen[0][1] == 0 and seen[-1][2] == 103
    assert all(seen[i][2] == seen[i + 1][1] for i in range(10))
    assert [c[0] for c in plan[1]] == [1, 5, 9]
    assert sum(shard.units_per_rank(103, 10, 4)) == 103
    with pytest.raises(ValueError):
        shard.chunk_plan(10, 0, 2)


def test_two_gloo_ranks_gather_in_read_order(tmp_path):
    out = str(tmp_path / "gathered.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    assert got["slowest"] == 11.0
    single = shard.map_sharded(N_READS, CHUNK, _oracle_mapper())
    assert len(got["res"]) == len(single) == (N_READS + CHUNK - 1) // CHUNK
    assert got["res"] == single
    # and the chunked result is the unchunked one: reads are independent
    whole_hits, whole_nper = _oracle_mapper()(0, N_READS)
    assert [h for c in single for h in c[0]] == whole_hits
    assert [n for c in single for n in c[1]] == whole_nper
