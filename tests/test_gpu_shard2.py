"""Two GPUs, the real mapper: shard.map_sharded deals chunks of one read set to two ranks (one process per GPU, its own
replica of the index, no collective on the data path), rank 0 gathers in chunk order -- the records equal those of one
GPU mapping the whole set.  Skipped on a one-GPU box."""
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

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")]

CHUNK = 250


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mapper(device):
    import shrimp_b200
    from mapcases import LsCase
    from shrimp_b200.api import MapParams, _pack_codes
    case = LsCase("c1_small")
    ctx = shrimp_b200.GpuContext(device)
    ctx.sw_setup(1400, 1000, case.scores)
    ctx.load_genome([_pack_codes(c.astype(np.uint32)) for c in case.contig_codes], [c.size for c in case.contig_codes])
    ctx.build_index(case.seeds)
    params = MapParams(list_cutoff=case.list_cutoff)

    def map_chunk(a, b):
        res = ctx.map_reads(params, case.scores, case.packed[a:b], case.read_len[a:b])
        recs = [(a + int(h["read_idx"]), int(h["cn"]), int(h["gen_st"]), int(h["genome_start"]), int(h["score_full"]),
                 bytes(res.edits[int(h["edit_off"]): int(h["edit_off"]) + int(h["edit_len"])])) for h in res.hits]
        return recs, res.n_hits_per_read.tolist()
    return case, map_chunk, ctx


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from shrimp_b200 import shard
        case, map_chunk, ctx = _mapper(rank)
        res = shard.map_sharded(len(case.read_len), CHUNK, map_chunk, rank, world)
        if rank == 0:
            torch.save(res, out_path)
        dist.barrier()
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_two_gpus_equal_one(tmp_path):
    out = os.path.join(str(tmp_path), "sharded.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    sharded = torch.load(out, weights_only=False)
    from shrimp_b200 import shard
    case, map_chunk, ctx = _mapper(0)
    try:
        single = shard.map_sharded(len(case.read_len), CHUNK, map_chunk, 0, 1)
    finally:
        ctx.close()
    assert len(sharded) == len(single) == (len(case.read_len) + CHUNK - 1) // CHUNK
    assert sharded == single
    assert sum(len(r) for r, _ in single) > 1000
