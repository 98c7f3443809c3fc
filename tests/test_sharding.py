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


# ---- GOP-segmented clips ---------------------------------------------------------------
class _FakeSession:
    """Stands in for lookahead.Lookahead on the CPU: decides every frame 'P' except the first ('IDR'), two
    frames late, in arrival order -- enough to check segment ownership, renumbering and stitching."""

    def __init__(self):
        self.n, self.pending, self.flushed, self.closed = 0, [], False, False

    def put_frame(self, frame):
        self.pending.append(dict(i_frame=self.n, i_type=1 if self.n == 0 else 3, payload=frame))
        self.n += 1

    def decisions(self):
        keep = 0 if self.flushed else 2
        ready, self.pending = self.pending[:len(self.pending) - keep], self.pending[len(self.pending) - keep:]
        return ready

    def flush(self):
        self.flushed = True

    def close(self):
        self.closed = True


def test_gop_segments_cover_the_clip():
    sys.path.insert(0, ROOT)
    from x264vfw_b200.sharding import gop_segments, segments_of_rank
    assert gop_segments(0, 8) == []
    assert gop_segments(20, 8) == [(0, 8), (8, 16), (16, 20)]
    for world in (1, 2, 3, 8):
        segs = sorted((k, a, b) for r in range(world) for k, a, b in segments_of_rank(300, 50, r, world))
        assert [s[0] for s in segs] == list(range(6))
        assert [f for _, a, b in segs for f in range(a, b)] == list(range(300))
    import pytest
    with pytest.raises(ValueError):
        gop_segments(10, 0)


def test_segmented_clip_runs_one_fresh_session_per_segment_and_stitches_in_order():
    sys.path.insert(0, ROOT)
    from x264vfw_b200.sharding import run_clip_segments, stitch_segments
    frames = [f"frame{i}" for i in range(21)]
    sessions = []

    def opener():
        sessions.append(_FakeSession())
        return sessions[-1]

    parts = {}
    for rank in range(2):
        parts.update(run_clip_segments(opener, frames, 8, rank, 2))
    assert len(sessions) == 3 and all(s.closed for s in sessions)
    clip = stitch_segments(parts)
    assert [d["i_frame"] for d in clip] == list(range(21))                     # every frame decided once, in order
    assert [d["payload"] for d in clip] == frames                              # ... from the right input
    assert [d["i_frame"] for d in clip if d["i_type"] == 1] == [0, 8, 16]      # each segment opens with its own IDR
    assert [d["segment"] for d in clip] == [0] * 8 + [1] * 8 + [2] * 5


def _segment_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from x264vfw_b200.sharding import run_clip_segments, gather_segments, stitch_segments
    frames = list(range(100, 130))
    mine = run_clip_segments(_FakeSession, frames, 7, rank, world)
    clip = stitch_segments(gather_segments(mine))
    dist.barrier()
    q.put((rank, sorted(mine), [d["i_frame"] for d in clip], [d["payload"] for d in clip]))
    dist.destroy_process_group()


def test_two_rank_gop_segmented_clip():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + os.getpid() % 100
    ps = [ctx.Process(target=_segment_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]                      # 30 frames / 7 = 5 segments, round-robin
    for r in res:                                                              # both ranks end up with the whole clip
        assert r[2] == list(range(30)) and r[3] == list(range(100, 130))
