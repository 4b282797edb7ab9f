"""N>1 path on CPU: world_size-2 gloo run of the replica plumbing (shard by volume, max-over-ranks timing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dg_tta_b200 import replicas


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = replicas.shard_items(8, rank, world)
        # pretend every volume costs (index+1) time units and has 1000 voxels
        seconds = float(sum(i + 1 for i in mine))
        thr = replicas.aggregate_throughput(1000.0 * len(mine), seconds)
        slowest = replicas.max_over_ranks(seconds)
        out.put((rank, mine, thr, slowest))
    finally:
        dist.destroy_process_group()


def test_two_rank_replicas_shard_and_time():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, t0, s0), (r1, m1, t1, s1) = res
    assert m0 == [0, 2, 4, 6] and m1 == [1, 3, 5, 7]          # disjoint, complete
    assert sorted(m0 + m1) == list(range(8))
    assert s0 == s1 == 20.0                                      # rank 1: 2+4+6+8 is the slowest
    assert abs(t0 - 8000.0 / 20.0) < 1e-9 and t0 == t1          # whole-job units / slowest rank


def test_single_process_is_identity():
    assert replicas.max_over_ranks(3.5) == 3.5
    assert replicas.shard_items(5, 0, 1) == [0, 1, 2, 3, 4]
    assert replicas.aggregate_throughput(10.0, 2.0) == 5.0
