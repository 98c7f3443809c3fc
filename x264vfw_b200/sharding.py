"""Multi-GPU partitioning of the path: frames of one stream are serially
dependent through the lookahead, independent streams are not, so the unit of sharding is the
stream: stream i -> rank i mod world_size, one session per stream, no data-path collective.
Timing of a multi-rank run is the max over ranks (each rank times its own device).

A single clip shards only as GOP segments: fixed-length runs of frames, each fed to its OWN session, so
each starts with an IDR and is a closed GOP by construction -- exactly what the reference produces when
one encoder instance (one CODEC) is opened per segment and the bitstreams are concatenated.  It is NOT
what one encoder produces over the whole clip (scene-cut / keyint counters, mb-tree and B-frame decisions
near a boundary differ), which is why SURVEY 8e allows it only when the CPU reference is cut at the same
points; otherwise a clip is "replicas only".  Segments are independent, so again no collective."""
import torch
import torch.distributed as dist


def streams_of_rank(n_streams: int, rank: int, world_size: int):
    """Global stream ids owned by `rank` (round-robin, like 'stream i -> GPU i mod G')."""
    return [s for s in range(n_streams) if s % world_size == rank]


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: int, device=None) -> int:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(value)
    t = torch.tensor([value], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


# ---- GOP-segmented clips -----------------------------------------------------------------------
def gop_segments(n_frames: int, segment_frames: int):
    """[start, end) frame ranges of a clip cut every `segment_frames` frames (the last one may be shorter)."""
    if n_frames < 0 or segment_frames <= 0:
        raise ValueError("n_frames >= 0 and segment_frames > 0 required")
    return [(s, min(s + segment_frames, n_frames)) for s in range(0, n_frames, segment_frames)]


def segments_of_rank(n_frames: int, segment_frames: int, rank: int, world_size: int):
    """(segment index, start, end) of the segments `rank` owns: segment k -> rank k mod world_size."""
    return [(k, a, b) for k, (a, b) in enumerate(gop_segments(n_frames, segment_frames)) if k % world_size == rank]


def stitch_segments(per_segment):
    """per_segment: {segment index: (start frame, decisions in that session's coded order)}.  Returns the
    clip's decisions in coded order with `i_frame` renumbered to clip frame numbers.  Concatenation is the
    coded order of the clip because no segment references a frame outside itself."""
    out = []
    for k in sorted(per_segment):
        start, decisions = per_segment[k]
        for d in decisions:
            e = dict(d)
            e["i_frame"] = d["i_frame"] + start
            e["segment"] = k
            out.append(e)
    return out


def run_clip_segments(open_session, frames, segment_frames: int, rank: int = 0, world_size: int = 1):
    """Feeds the segments this rank owns through one fresh session each (open_session() -> an object with
    put_frame / flush / decisions / close, e.g. lookahead.Lookahead) and returns {segment index: (start,
    decisions)} for stitch_segments.  `frames` is indexable by clip frame number."""
    done = {}
    for k, a, b in segments_of_rank(len(frames), segment_frames, rank, world_size):
        la = open_session()
        try:
            got = []
            for i in range(a, b):
                la.put_frame(frames[i])
                got += la.decisions()
            la.flush()
            got += la.decisions()
        finally:
            la.close()
        done[k] = (a, got)
    return done


def gather_segments(mine: dict) -> dict:
    """Union of every rank's {segment: (start, decisions)} (control data only; all_gather_object)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(mine)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, mine)
    out = {}
    for p in parts:
        out.update(p)
    return out
