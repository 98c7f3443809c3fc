"""CPU test of the N>1 path: world_size-2 gloo processes partition streams round-robin with
no overlap and reduce their timings with max / counts with sum (no data-path collective)."""
import os
import sys

import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from x264vfw_b200.sharding import streams_of_rank, max_over_ranks, sum_over_ranks
    mine = streams_of_rank(8, rank, world)
    ms = max_over_ranks(10.0 + 5 * rank)
    total = sum_over_ranks(len(mine))
    dist.barrier()
    q.put((rank, mine, ms, total))
    dist.destroy_process_group()


def test_two_rank_stream_partition_and_reductions():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29640 + os.getpid() % 200
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5, 7]
    assert all(r[2] == 15.0 for r in res)          # max over ranks
    assert all(r[3] == 8 for r in res)             # every stream owned exactly once


def test_partition_is_a_disjoint_cover_for_any_world_size():
    sys.path.insert(0, ROOT)
    from x264vfw_b200.sharding import streams_of_rank
    for world in (1, 2, 3, 4, 8):
        for n in (1, 5, 8, 64):
            got = sorted(s for r in range(world) for s in streams_of_rank(n, r, world))
            assert got == list(range(n))
