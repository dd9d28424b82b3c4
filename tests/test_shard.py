"""Host-side sharding logic, including a world_size-2 gloo run on CPU (no data-path collective exists;
the only cross-rank traffic is the barrier / max-over-ranks used for timing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opencl_fft_b200.shard import max_over_ranks, shard_range, sum_over_ranks


def test_shard_range_covers_and_balances():
    for total in (0, 1, 7, 64, 1024, 1027):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(1024, rank, world)
    spans = [None] * world
    dist.all_gather_object(spans, (lo, hi))
    slow = max_over_ranks(10.0 + rank)  # the slowest rank defines the step time
    tot = sum_over_ranks(hi - lo)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, spans, slow, tot))


def test_two_rank_gloo_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, spans, slow, tot in res:
        assert spans == [(0, 512), (512, 1024)]
        assert slow == 11.0
        assert tot == 1024
