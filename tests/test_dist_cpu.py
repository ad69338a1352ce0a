"""CPU, world_size 2, gloo: the segment-wise gradient all-reduce used by the N>1 path (vcvits_b200/dist.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vcvits_b200.dist import SegmentReducer


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)
        flat = torch.randn(1000)
        mine = flat.clone()
        ranges = [(0, 10), (10, 10), (10, 400), (400, 1000)]  # includes an empty segment
        red = SegmentReducer(flat, ranges, dist.group.WORLD)
        for seg in range(len(ranges)):
            red.segment_done(seg)
        red.finish()
        gathered = [torch.zeros(1000) for _ in range(world)]
        dist.all_gather(gathered, mine)
        expect = sum(gathered) / world
        ok = torch.allclose(flat, expect, atol=1e-6)
        # every rank ends with identical gradients
        same = [torch.zeros(1000) for _ in range(world)]
        dist.all_gather(same, flat)
        ok = ok and all(torch.equal(same[0], t) for t in same)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_segment_reducer_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_segment_reducer_single_process_is_noop():
    flat = torch.arange(10.0)
    red = SegmentReducer(flat, [(0, 5), (5, 10)], None)
    red.segment_done(0)
    red.segment_done(1)
    red.finish()
    assert torch.equal(flat, torch.arange(10.0))
