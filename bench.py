#!/usr/bin/env python
"""bench.py -- 1080p frames/s through csp + lookahead on B200 (BASELINE.json metric).

Workload (config.workload "C5"): per GPU, S concurrent independent 1080p RGB32 bottom-up DIB
streams, each: BGRA|VFLIP -> I420 (stage 1) -> AQ statistics + lowres planes -> x264 lookahead
with preset medium (bframes 3, b-adapt 1, rc-lookahead 40, mb-tree, scenecut 40, weightp 2,
aq-mode 1), one encoder session per stream, one native host thread per stream.  Every stream plays
the clip SURVEY 8(d) pins (300 frames, hard cuts at 100 and 200, two-frame white flash at 150; seed
0x264 + stream id), back and forth.  A "step" is one pass of the hot path over one batch: every
stream consumes F consecutive frames and returns the frame types / qp offsets decided meanwhile.
Sessions persist across steps (steady state; the lookahead window is pre-filled before the warm-up).
N GPUs = N ranks x S streams (weak scaling, no collective: streams are independent).

  value : frames/s with the packed clips already resident in HBM.
  e2e   : the same through the C-ABI call with HOST (pinned) buffers: H2D of every packed
          frame and D2H of the converted planes (codec->conv_pic) + decisions inside the timed
          region.
  --impl reference : the SAME job on the host cores: persistent sessions, window pre-filled, same
          clips / streams / preset; stage 1 = the unmodified reference csp.c object (oracle/_ref),
          stage 2 = the CPU restatement of the libx264 lookahead (libx264 itself is not vendored by
          the reference).  Each step is a bounded sample (16 of the F frames per stream).  This arm
          never imports the product package.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# one hardware work queue per CUDA stream (8 sessions x 3 streams); must be set before CUDA initialises
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

W, H = 1920, 1080
PRESET = "medium"
BGRA_FLIP = 9 | 0x1000
SRC_BYTES = W * H * 4
DST_BYTES = W * H * 3 // 2
MB_COUNT = 120 * 68
CSP_ALGO_BYTES = SRC_BYTES + DST_BYTES            # 11,404,800 B / frame
LOWRES_ALGO_BYTES = 1920 * 1088 + 4 * 960 * 544   # 4,177,920 B / frame
FUSED_ALGO_BYTES = CSP_ALGO_BYTES + 4 * 960 * 544  # csp + lowres, luma never re-read: 13,493,760 B / frame
HPEL_ALGO_BYTES = 5 * 1920 * 1088                 # reconstructed plane in, 4 reference planes out: 10,444,800 B / frame
CLIP_FRAMES, CLIP_CUTS, CLIP_FLASH = 300, (100, 200), 150
REF_SAMPLE_FRAMES = 16                            # frames per stream per step of the CPU arms (a sample of F)
WORKLOAD = ("C5: 1080p RGB32 bottom-up DIB -> I420 + x264 lookahead preset medium, independent streams, one session per stream; "
            "SURVEY 8(d) clip (300 frames, cuts at 100/200, 2-frame flash at 150) played back and forth")


def bind_to_gpu_numa(index):
    """Run this rank (its session threads, its pinned buffers by first touch) on the CPUs NVML reports as local to the
    GPU: at N > 1 a rank whose threads or pinned memory sit on the other socket pays for every copy twice."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(index)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(hnd, n)
        cpus = {64 * i + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=8, help="independent streams per GPU")
    ap.add_argument("--frames-per-step", type=int, default=160, help="frames per stream per step (timed region >= 2 s at the default 20 steps)")
    ap.add_argument("--clip-frames", type=int, default=CLIP_FRAMES)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-worst-case", action="store_true")
    ap.add_argument("--stage1-batch", type=int, default=96, help="frames per launch for the stage-1 roofline probe")
    return ap.parse_args()


def workload_config(args):
    """The keys that define the job; identical in both arms."""
    return {"workload": WORKLOAD, "preset": PRESET, "width": W, "height": H, "streams_per_gpu": args.streams,
            "frames_per_step_per_stream": args.frames_per_step, "clip_frames": args.clip_frames,
            "rc_lookahead": 40, "bframes": 3, "b_adapt": 1, "mbtree": 1, "weightp": 2, "aq_mode": 1, "lookahead_threads": 1}


def clip_of(stream_id, n_frames):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from clipgen import SyntheticClip        # test / bench infrastructure, not part of the product package
    cuts = tuple(c * n_frames // CLIP_FRAMES for c in CLIP_CUTS)
    return SyntheticClip(W, H, n_frames=n_frames, stream_id=stream_id, cuts=cuts, flash=CLIP_FLASH * n_frames // CLIP_FRAMES, flash_len=2)


def generate_clips(stream_ids, n_frames, sink):
    """Generates every stream's packed BGRA frames (one worker thread per clip: a clip caches the texture of the
    scene it is in) and hands each frame to sink(k, n, frame)."""
    from concurrent.futures import ThreadPoolExecutor

    def one(k):
        clip = clip_of(stream_ids[k], n_frames)
        for n in range(n_frames):
            sink(k, n, clip.packed(n, "bgra"))

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(stream_ids)))) as ex:
        list(ex.map(one, range(len(stream_ids))))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def _run_nvml(self):
        # same counters as the nvidia-smi clocks line of B200_PROFILING.md, sampled in-process so
        # that short timed regions still get many samples
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(hnd, pynvml.NVML_CLOCK_SM))
        bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop_evt.is_set():
            self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM)))
            r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hnd)
            for n, b in bits.items():
                if r & b:
                    self.reasons.add(n)
            self._stop_evt.wait(0.01)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.sm_max = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=10)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(s)}


def run_phase(torch, dist, sessions, frames, on_device, conv, args, sampler_index, steps, profile=False):
    """Prefill + warm-up + timed steps, then end of stream.  Returns a dict: ms_per_step (max over ranks), clocks,
    launches, per-kernel-class profile, host counters, device work counters, frames fed / decided."""
    import x264vfw_b200 as xv
    from x264vfw_b200.harness import StreamSet
    from x264vfw_b200.sharding import max_over_ranks, sum_over_ranks
    S, F = len(sessions), args.frames_per_step
    # one NATIVE host thread per stream (host/x264vfw_harness.c), like the reference's app threads
    streams = StreamSet(sessions, frames, on_device, conv)
    # prefill the lookahead window (rc-lookahead 40 + bframes) so that timed steps are steady state
    streams.run(sessions[0].p.rc_lookahead + sessions[0].p.bframes + 5)
    for _ in range(args.warmup):
        streams.run(F)
    if profile:
        for la in sessions:
            la.profile(1)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(sampler_index)
    sampler.start()
    c0 = [la.counters() for la in sessions]
    d0 = sum(streams.decided)
    n0 = xv.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        streams.run(F)
    torch.cuda.synchronize()
    e1.record()
    e1.synchronize()
    if dist is not None:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = xv.launch_count() - n0
    clocks = sampler.stop()
    c1 = [la.counters() for la in sessions]
    decided_timed = sum(streams.decided) - d0
    out = {"host": {k: sum(b[k] - a[k] for a, b in zip(c0, c1)) for k in c1[0]}, "prof": None, "stats": None}
    if profile:
        prof, stats = {}, {}
        for la in sessions:
            for k, v in la.stats().items():
                stats[k] = stats.get(k, 0) + v
            for k, (t, n) in la.profile(0).items():
                a = prof.setdefault(k, [0.0, 0])
                a[0] += t
                a[1] += n
        out["prof"], out["stats"] = prof, stats
    # end of stream: every frame fed must come back decided (no work skipped or left behind)
    fed = sum(int(streams.arr[s].pos) for s in range(S))
    streams.flush()
    decided = sum(streams.decided)
    assert decided == fed, f"decisions drained {decided} != frames fed {fed}"
    assert abs(decided_timed - steps * S * F) <= S * 2 * (sessions[0].p.bframes + 2), (decided_timed, steps * S * F)
    streams.close()
    out.update(ms_per_step=max_over_ranks(ms, device="cuda") / steps, clocks=clocks, launches=sum_over_ranks(launches, device="cuda"),
               frames_fed=fed, decisions_drained=decided, decisions_in_timed_region=decided_timed)
    return out


def stage1_roofline(torch, args):
    """Device-resident batch launches of the stage-1 kernels (the HBM-bound part of the path):
    CUDA events on the launching stream, inputs larger than L2."""
    import x264vfw_b200 as xv
    from x264vfw_b200 import csp, lowres, hpel
    ctx = xv._lib.Context(torch.cuda.current_device())
    st = torch.cuda.ExternalStream(ctx.stream)
    nf = args.stage1_batch
    sfb, dfb = csp.frame_bytes(BGRA_FLIP, 2, W, H)
    src = torch.randint(0, 256, (nf * sfb,), dtype=torch.uint8, device="cuda")
    dst = torch.empty(nf * dfb, dtype=torch.uint8, device="cuda")
    g = lowres.geometry(W, H)
    lr = torch.empty(nf * 4 * g.lplane_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    def timeit(fn, iters=10):
        for _ in range(3):
            fn()
        ctx.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(iters):
            fn()
        b.record(st)
        b.synchronize()
        return a.elapsed_time(b) / iters * 1e-3

    out = {}

    def leg(name, algo_bytes, n, fn):
        try:          # a failing leg is reported in the line, never hidden, and cannot take the headline numbers down
            t = timeit(fn)
            out[name] = {"frames_per_launch": n, "ms_per_launch": t * 1e3, "gbs": algo_bytes * n / t / 1e9, "algorithmic_bytes_per_frame": algo_bytes}
        except Exception as e:
            out[name] = {"error": f"{type(e).__name__}: {e}"}

    leg("csp_bgra_to_i420", CSP_ALGO_BYTES, nf, lambda: csp.convert_batch(ctx, src.data_ptr(), dst.data_ptr(), BGRA_FLIP, 2, 2, 0, W, H, nf))
    leg("lowres_init", LOWRES_ALGO_BYTES, nf, lambda: lowres.lowres_init(ctx, lr.data_ptr(), dst.data_ptr(), W, W, H, dfb, 4 * g.lplane_bytes, nf))
    if hasattr(lowres, "FusedBatch"):
        # csp -> I420 planes + AQ statistics + the four lowres planes in ONE pass over the packed rows
        fe = lowres.FusedBatch(ctx, W, H, nf)
        leg("fused_csp_aq_lowres", FUSED_ALGO_BYTES, nf, lambda: fe.run(src.data_ptr(), dst.data_ptr(), lr.data_ptr()))
    # half-pel reference planes: a mod-16 reconstructed plane per frame, half the batch
    hg = hpel.geometry(W, 1088)
    hn = max(1, nf // 2)
    rec = torch.randint(0, 256, (hn * W * 1088,), dtype=torch.uint8, device="cuda")
    hp = torch.empty(hn * 4 * hg.plane_bytes, dtype=torch.uint8, device="cuda")
    leg("hpel_filter", HPEL_ALGO_BYTES, hn, lambda: hpel.hpel_filter(ctx, hp.data_ptr(), rec.data_ptr(), W, W, 1088, W * 1088, 4 * hg.plane_bytes, hn))
    # integral image behind the half-pel planes (me esa / tesa): 8x8 box sums of the padded plane 0 just produced
    try:
        from x264vfw_b200 import b3
        prow = hg.plane_bytes // hg.stride
        s8 = torch.empty(hn * prow * hg.stride, dtype=torch.int16, device="cuda")
        leg("integral_init8", prow * hg.stride * 3, hn,
            lambda: b3.integral_init(ctx, s8.data_ptr(), 0, hp.data_ptr(), hg.stride, prow, 4 * hg.plane_bytes, prow * hg.stride, hn))
        ctx.sync()
        del s8
    except Exception as e:
        out["integral_init8"] = {"error": f"{type(e).__name__}: {e}"}
    # decoder-side output conversion: decoded yuv420p pictures -> bottom-up RGB32 DIBs, the default VfW
    # decompress target; algorithmic bytes = 1.5 read + 4 written per pixel.  `dst` doubles as the yuv420p source.
    try:
        from x264vfw_b200 import decode
        dn = max(1, nf // 2)
        dd = decode.Decompressor(csp.X264VFW_CSP_BGRA | csp.X264VFW_CSP_VFLIP, W, H, decode.AVCOL_SPC_BT709, False, ctx=ctx)
        dib = torch.empty(dn * W * H * 4, dtype=torch.uint8, device="cuda")
        yuv = torch.randint(0, 256, (dn * dfb,), dtype=torch.uint8, device="cuda")
        b0 = yuv.data_ptr()
        leg("decode_yuv420p_to_bgra", W * H * 3 // 2 + W * H * 4, dn,
            lambda: dd.decompress_batch(dib.data_ptr(), W * H * 4, (b0, b0 + W * H, b0 + W * H * 5 // 4), (W, W // 2, W // 2), dfb, dn))
        ctx.sync()
        dd.close()
        if "gbs" in out["decode_yuv420p_to_bgra"] and not args.no_cpu_baseline:
            # the reference's own CPU path for this stage is libswscale (codec.c:2292): time the copy this image carries,
            # driven like the reference drives it, one thread, five pictures
            try:
                sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
                import numpy as np
                import swsref
                if swsref.available():
                    rng = np.random.default_rng(0)
                    py, pu, pv = (rng.integers(0, 256, sh, dtype=np.uint8) for sh in ((H, W), (H // 2, W // 2), (H // 2, W // 2)))
                    tm = []
                    swsref.decompress_convert(py, pu, pv, 9 | 0x1000, 1, 0, repeat=5, timing=tm)
                    out["decode_yuv420p_to_bgra"]["cpu_reference"] = {"kind": "libswscale " + swsref.version() + " sws_scale, context built once, 1 thread",
                                                                      "ms_per_picture": tm[0] * 1e3}
            except Exception as e:
                out["decode_yuv420p_to_bgra"]["cpu_reference"] = {"error": f"{type(e).__name__}: {e}"}
    except Exception as e:
        out["decode_yuv420p_to_bgra"] = {"error": f"{type(e).__name__}: {e}"}
    try:
        ctx.close()
    except Exception:
        pass
    return out


# ------------------------------------------------------------------------------------------------------------------
# CPU arms (oracle/ may only be executed here: cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(n_streams, clips, frames_per_step, steps, warmup, prefill):
    """The job of the GPU arm on host cores: persistent sessions (one per stream, one thread per stream like
    codec.c:1728), window pre-filled, `frames_per_step` frames per stream per step, clip played back and forth.
    Stage 1 = unmodified reference csp.c (oracle/_ref) when present, else the csp port; stage 2 = CPU restatement of
    the libx264 lookahead.  Returns (frames/s mean over timed steps, seconds per step, kind, decisions drained)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    # the CPU baseline gets the best scalar code gcc gives on THIS box: oracle/Makefile `native` (-O3 -march=native,
    # same sources, same results); the reference csp.c object stays the -O2 build that travels from the authoring container
    native = os.path.join(ROOT, "oracle", "_ref", "liboracle_native.so")
    if "oracle_lib" not in sys.modules:
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "native"], capture_output=True)
        if r.returncode == 0 and os.path.exists(native):
            os.environ["X264VFW_ORACLE_SO"] = native
    import oracle_lib as ol
    cpu_reference_run.flags = "gcc -O3 -march=native" if ol.ORACLE_SO == native else "gcc -O3 -march=x86-64-v2"
    have_ref = ol.have_ref_csp()
    conv = ol.ref_convert if have_ref else ol.oracle_convert
    ol.oracle()
    ol.la_params(PRESET, W, H)          # build tables before threading
    sessions = [ol.OracleLookahead(ol.la_params(PRESET, W, H)) for _ in range(n_streams)]
    pos = [0] * n_streams
    decided = [0] * n_streams

    def feed(s, count):
        la, clip = sessions[s], clips[s % len(clips)]
        n = len(clip)
        for _ in range(count):
            k = pos[s] % (2 * n - 2) if n > 1 else 0
            if k >= n:
                k = 2 * n - 2 - k
            la.put_i420(conv(clip[k], BGRA_FLIP, 2, 2, 0, W, H))
            decided[s] += len(la.decisions())
            pos[s] += 1

    def step(count):
        ths = [threading.Thread(target=feed, args=(s, count)) for s in range(n_streams)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0

    step(prefill)
    for _ in range(warmup):
        step(frames_per_step)
    times = [step(frames_per_step) for _ in range(steps)]
    for s, la in enumerate(sessions):
        la.flush()
        decided[s] += len(la.decisions())
        la.close()
    assert decided == pos, (decided, pos)
    fps = [n_streams * frames_per_step / t for t in times]
    return sum(fps) / len(fps), sum(times) / len(times), ("reference csp.c + port" if have_ref else "port"), sum(decided)


def host_clips(n_clips, n_frames, clip_len):
    """The first n_frames frames of n_clips clips of length clip_len (same content as the GPU arm's clips)."""
    from concurrent.futures import ThreadPoolExecutor

    def one(k):
        clip = clip_of(k, clip_len)
        return [clip.packed(n, "bgra") for n in range(n_frames)]

    with ThreadPoolExecutor(max_workers=min(8, max(1, n_clips))) as ex:
        return list(ex.map(one, range(n_clips)))


def libx264_probe():
    """a libx264 dropped under baseline/_ref would be the real stage-2 reference (oracle/Makefile builds
    oracle/_ref/x264_ref_probe against it).  Absent from this image; reported so that the label is never wrong."""
    for base in (os.path.join(ROOT, "baseline", "_ref"), os.path.join(ROOT, "oracle", "_ref")):
        for dirpath, _, files in os.walk(base):
            for f in files:
                if f.startswith("libx264.so") or f == "libx264.a":
                    return os.path.join(dirpath, f)
    return None


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], capture_output=True)
    cores = os.cpu_count() or 1
    n_streams = args.streams * max(1, args.gpus)          # the same job: S streams per GPU
    prefill = 40 + 3 + 5
    # frames the arm will actually play (the clip is played back and forth from frame 0)
    need = min(args.clip_frames, prefill + (args.warmup + args.steps) * REF_SAMPLE_FRAMES)
    clips = host_clips(min(n_streams, args.streams), need, args.clip_frames)
    fps, sec, kind, decided = cpu_reference_run(n_streams, clips, REF_SAMPLE_FRAMES, args.steps, args.warmup, prefill)
    x264 = libx264_probe()
    sample = (f"same job as the GPU arm: {n_streams} persistent sessions (one thread each; {cores} host cores), window pre-filled with {prefill} frames, "
              f"{args.warmup} warm-up + {args.steps} timed steps; each step is a bounded sample of {REF_SAMPLE_FRAMES} of the {args.frames_per_step} frames "
              f"per stream (first {need} frames of each clip, back and forth); {min(n_streams, args.streams)} distinct clips reused round-robin; "
              f"stage 1 = {'unmodified reference csp.c (oracle/_ref, gcc -O2)' if 'reference' in kind else 'csp port'}, "
              f"stage 2 = CPU restatement of the libx264 lookahead, scalar C ({cpu_reference_run.flags}), no asm -- NOT libx264"
              + (f"; a libx264 exists at {x264} but is not wired in" if x264 else " (none in the image or under baseline/_ref)"))
    line = {"impl": "reference", "metric": "1080p frames/sec through csp+lookahead", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": min(n_streams, cores), "kind": "port", "sample": sample,
                             "threads": n_streams, "decisions_drained": decided},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def ncu_metrics():
    """Per-kernel numbers taken from the tracked ncu captures of this round (profiles/ncu_metrics_r2.json, written by
    scripts/summarize_profiles.py from the .ncu-rep files): static properties of the kernels (instructions per MB
    search, DRAM bytes per launch), never timings."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_metrics_r2.json")))
    except Exception:
        return {}


def main():
    args = parse_args()
    if args.impl == "reference":
        return main_reference(args)
    if args.warmup < 3:
        args.warmup = 3
    import torch
    import x264vfw_b200 as xv  # noqa: F401
    from x264vfw_b200 import csp, lookahead

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 and os.environ.get("X264VFW_BENCH_NUMA", "1") != "0" else None
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    n_gpus = world
    S, F = args.streams, args.frames_per_step

    # one caller thread + one session worker per stream; they spin while they wait (lowest latency).  On a node with fewer
    # cores than such threads: first the CALLERS stop polling (they sleep on the interrupt while their copies run; with
    # resident clips they sleep on a condition variable anyway), and when even the workers alone outnumber the cores the
    # workers yield the core between polls.
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    cores = os.cpu_count() or 1
    if "X264VFW_CUDA_SYNC" not in os.environ and local_world * S * 2 > cores:
        os.environ.setdefault("X264VFW_CUDA_CALLER_BLOCK", "1")
        if local_world * (S + 2) > cores:
            os.environ["X264VFW_CUDA_SYNC"] = "yield"
    from x264vfw_b200.sharding import streams_of_rank
    my_streams = streams_of_rank(S * world, rank, world)   # global stream ids of this rank (S per GPU)
    assert len(my_streams) == S

    # ---- clips: resident in HBM (value); a pinned host copy for e2e, sized to the node's memory ----
    e2e_frames = 0
    if not args.no_e2e:
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 64 << 30
        e2e_frames = int(min(args.clip_frames, 0.25 * avail / (local_world * S * SRC_BYTES)))
        if e2e_frames < 64:
            e2e_frames = min(args.clip_frames, 64)
    dev_frames = [[None] * args.clip_frames for _ in range(S)]
    pin_frames = [[None] * e2e_frames for _ in range(S)]

    dev = torch.device("cuda", local_rank)

    def sink(k, n, f):
        t = torch.from_numpy(f)
        if n < e2e_frames:
            t = t.pin_memory()
            pin_frames[k][n] = t.numpy()
        dev_frames[k][n] = t.to(dev)          # explicit device: the generator threads' current CUDA device is not this rank's

    generate_clips(my_streams, args.clip_frames, sink)
    dev_ptrs = [[t.data_ptr() for t in clip] for clip in dev_frames]
    torch.cuda.synchronize()

    def open_sessions():
        return [lookahead.Lookahead(lookahead.params_preset(PRESET, W, H), in_csp=BGRA_FLIP, out_csp=csp.X264_CSP_I420,
                                    colmatrix=2, fullrange=0, device=local_rank) for _ in range(S)]

    def phase(frames, on_device, conv, steps, profile=False, env=None):
        old = {}
        for k, v in (env or {}).items():
            old[k] = os.environ.get(k)
            os.environ[k] = v
        try:
            sessions = open_sessions()
            r = run_phase(torch, dist, sessions, frames, on_device, conv, args, local_rank, steps, profile=profile)
            for la in sessions:
                la.close()
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        return r

    # resident clips: X264VFW_CUDA_SRC_RESIDENT (the session never waits for the reader of a source frame)
    main_r = phase(dev_ptrs, 2, None, args.steps)
    # same workload once more with per-kernel CUDA-event timing and the device work counters switched on
    prof_steps = max(2, args.steps // 4)
    prof_r = phase(dev_ptrs, 2, None, prof_steps, profile=True)
    # and with the speculation made useless: the ordered verification keeps nothing (0 % hit rate), same results
    worst_r = None if args.no_worst_case else phase(dev_ptrs, 2, None, max(1, args.steps // 10), env={"X264VFW_CUDA_ME_FORCE_MISS": "1"})
    frames_per_step_all = S * F * n_gpus
    value = frames_per_step_all / (main_r["ms_per_step"] * 1e-3)

    e2e = None
    if not args.no_e2e:
        conv = [[torch.empty(DST_BYTES, dtype=torch.uint8).pin_memory().numpy() for _ in range(4)] for _ in range(S)]
        e2e_r = phase(pin_frames, 0, conv, args.steps)
        e2e = {"value": frames_per_step_all / (e2e_r["ms_per_step"] * 1e-3), "unit": "frames/s",
               "h2d_bytes_per_step": n_gpus * S * F * SRC_BYTES, "d2h_bytes_per_step": n_gpus * S * F * (DST_BYTES + 2 * 4 * MB_COUNT + 32),
               "ms_per_step": e2e_r["ms_per_step"], "clip_frames_pinned": e2e_frames,
               "frames_fed": e2e_r["frames_fed"], "decisions_drained": e2e_r["decisions_drained"]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"

    stage1 = stage1_roofline(torch, args)
    for k in stage1:
        if "gbs" in stage1[k]:
            stage1[k].update(frac=stage1[k]["gbs"] / hbm_peak, peak=hbm_peak, bound="hbm")

    # ---- the dominant kernels of the step by device time, from the profiled phase ----
    prof, stats, pdelta = prof_r["prof"], prof_r["stats"], prof_r["host"]
    t_prof = prof_r["ms_per_step"] * prof_steps * 1e-3
    tot = sum(v[0] for v in prof.values()) or 1.0
    shares = {k: {"ms": v[0], "launches": v[1], "share": v[0] / tot} for k, v in prof.items()}
    dom = max(prof, key=lambda k: prof[k][0])
    # algorithmic bytes of one search: fenc plane + 4 ref planes + per-MB mv/cost
    me_bytes_per_search = 960 * 544 * 5 + MB_COUNT * 8
    me_ms, me_n = prof["me"]
    pass_ms, pass_n = prof.get("me_pass", (0.0, 0))
    me_avg = (me_ms + pass_ms) / max(1, me_n)
    searches = pdelta["mb_searches"] / MB_COUNT
    searches_per_launch = searches / max(1, me_n)
    ncu = ncu_metrics()
    clock_hz = (main_r["clocks"].get("sm_mhz") or 1965.0) * 1e6
    issue_peak = 148 * 4 * clock_hz                      # warp instructions per second the SM sub-partitions can issue
    mb_pass = stats["pass0"] + stats["pass1"] + stats["pass2"] + stats["pass3"]
    mb_order = stats["researched"]
    integer = {"unit": "warp instructions/s", "peak": issue_peak,
               "peak_source": "148 SMs x 4 schedulers x SM clock sampled in the timed region",
               "mb_searches_parallel_passes_per_s": mb_pass / t_prof, "mb_searches_in_order_per_s": mb_order / t_prof,
               "pixel_diff_ops_per_s": 64 * (stats["sad8x8"] + stats["satd8x8"]) / t_prof,
               "sad8x8_per_s": stats["sad8x8"] / t_prof, "satd8x8_per_s": stats["satd8x8"] / t_prof,
               "sad8x8_per_mb_search": stats["sad8x8"] / max(1, mb_pass + mb_order), "satd8x8_per_mb_search": stats["satd8x8"] / max(1, mb_pass + mb_order)}
    ipm = (ncu.get("me_pass_kernel") or {}).get("warp_inst_per_mb_search")
    if ipm:
        integer.update(warp_inst_per_mb_search=ipm, achieved=ipm * mb_pass / t_prof, frac=ipm * mb_pass / t_prof / issue_peak,
                       source="instructions per MB search: profiles/ncu_metrics_r2.json (ncu smsp__inst_executed.sum of me_pass_kernel / MBs it searched); "
                              "MB searches per second: device counters of this run")
    traffic = None
    if (ncu.get("me_pass_kernel") or {}).get("dram_bytes_per_search") is not None and (ncu.get("me_verify_kernel") or {}).get("dram_bytes_per_search") is not None:
        traffic = (ncu["me_pass_kernel"]["dram_bytes_per_search"] + ncu["me_verify_kernel"]["dram_bytes_per_search"]) * searches_per_launch
    kept = stats["kept"] / max(1, stats["kept"] + stats["researched"])
    roofline = {"kernel": "motion search: me_pass_kernel (speculative parallel passes) + me_verify_kernel (exact ordered "
                          "verification); integer pipe / dependent-MB latency, see DESIGN.md 4.1",
                "bound": "hbm", "achieved": (me_bytes_per_search * searches_per_launch / (me_avg * 1e-3) / 1e9) if me_n else None,
                "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": traffic,
                "traffic_source": "profiles/ncu_metrics_r2.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per search)" if traffic is not None else None,
                "peak_source": peak_src,
                "avg_launch_ms": me_avg, "launches": me_n, "searches_per_launch": searches_per_launch,
                "avg_pass_ms": pass_ms / max(1, pass_n), "avg_verify_ms": me_ms / max(1, me_n),
                "algorithmic_bytes_per_search": me_bytes_per_search, "mb_searches_per_s": pdelta["mb_searches"] / t_prof,
                "share_of_step_device_time": shares["me"]["share"] + shares.get("me_pass", {"share": 0.0})["share"],
                "integer": integer,
                "speculation": {"kept_fraction": kept, "mbs_kept": stats["kept"], "mbs_researched_in_order": stats["researched"],
                                "mbs_searched_by_pass": [stats["pass0"], stats["pass1"], stats["pass2"], stats["pass3"]],
                                "searches_speculative": stats["spec_jobs"], "searches_on_demand": stats["ondemand_jobs"],
                                "searches_the_decision_logic_asked_for": stats["searches_asked_for"],
                                "wasted_speculation": 1.0 - stats["searches_asked_for"] / max(1, stats["spec_jobs"] + stats["ondemand_jobs"]),
                                "note": "search counts since the sessions were opened (prefill and warm-up included)"},
                "note": "dominant kernels by device time; their bound is neither HBM nor tensor (SURVEY 8(d)): algorithmic traffic is ~2.7 MB per search "
                        "against 8160 MB searches with a dependent chain, so the GB/s figure is tiny by construction -- the integer roofline (issue slots) "
                        "is in `integer`, the HBM-bound kernels of the path are in `stage1`."}
    if roofline["achieved"] is not None:
        roofline["frac"] = roofline["achieved"] / hbm_peak
    tree_ms, tree_n = prof.get("mbtree", (0.0, 0))
    mbtree = {"walks": stats["tree_walks"], "steps": stats["tree_steps"], "steps_per_walk": stats["tree_steps"] / max(1, stats["tree_walks"]),
              "device_ms_per_walk": tree_ms / max(1, tree_n), "steps_per_s_inside_a_walk": stats["tree_steps"] / max(1e-9, tree_ms * 1e-3),
              "steps_per_s_whole_job": stats["tree_steps"] / t_prof,
              "qp_offset_parity": "bit-exact vs the checker (0 ulp; north_star asks 1e-5 relative): tests/test_lookahead_gpu.py, test_baseline_configs_gpu.py"}

    worst = None
    if worst_r is not None:
        worst = {"value": frames_per_step_all / (worst_r["ms_per_step"] * 1e-3), "unit": "frames/s", "ms_per_step": worst_r["ms_per_step"],
                 "what": "X264VFW_CUDA_ME_FORCE_MISS=1: the ordered verification keeps no speculative result (0 % hit rate), every MB is searched again in "
                         "dependency order after the parallel passes; identical decisions"}

    cpu_baseline = None
    if not args.no_cpu_baseline and n_gpus == 1:        # reported at N = 1 only (the --impl reference arm covers every N)
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], capture_output=True)
        cores = os.cpu_count() or 1
        clips = host_clips(S, 96, args.clip_frames)
        fps, sec, kind, decided = cpu_reference_run(S, clips, REF_SAMPLE_FRAMES, 4, 1, 48)
        cpu_baseline = {"value": fps, "unit": "frames/s", "cores": min(S, cores), "kind": "port", "threads": S,
                        "sample": f"{S} persistent sessions (one thread each; {cores} host cores), window pre-filled with 48 frames, 1 warm-up + 4 timed steps of "
                                  f"{REF_SAMPLE_FRAMES} frames per stream on the first 96 frames of the same clips; "
                                  f"stage 1 = {'unmodified reference csp.c (oracle/_ref, gcc -O2)' if 'reference' in kind else 'csp port'}, "
                                  f"stage 2 = CPU restatement of the libx264 lookahead (scalar C, {cpu_reference_run.flags}, no asm; libx264 is not vendored by the reference)"}

    host_delta = main_r["host"]
    line = {"metric": "1080p frames/sec through csp+lookahead", "value": value, "unit": "frames/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args),
            "run": {"host_wait": os.environ.get("X264VFW_CUDA_SYNC", "spin") + ("+callers block" if os.environ.get("X264VFW_CUDA_CALLER_BLOCK") == "1" else ""), "rank0_cpu_affinity": (f"{len(numa_cpus)} CPUs local to the GPU" if numa_cpus else "unrestricted"),
                    "l2_policy": f"inputs larger than L2: {S * F * SRC_BYTES / 1e6:.0f} MB of packed frames per step per GPU",
                    "timed_region_s": main_r["ms_per_step"] * args.steps * 1e-3,
                    "frames_fed": main_r["frames_fed"], "decisions_drained": main_r["decisions_drained"],
                    "decisions_in_timed_region": main_r["decisions_in_timed_region"], "frames_in_timed_region": args.steps * S * F},
            "clocks": main_r["clocks"], "gpu_launches": main_r["launches"], "e2e": e2e, "roofline": roofline, "stage1": stage1,
            "worst_case": worst, "mbtree": mbtree,
            "kernel_shares": shares, "dominant_kernel_class": dom, "cpu_baseline": cpu_baseline,
            "host_us_per_frame": {k: host_delta[k] / max(1, host_delta["frames"]) for k in ("put_us", "decide_us", "sync_us")},
            "per_frame": {k: host_delta[k] / max(1, host_delta["frames"]) for k in ("frame_costs", "launches", "syncs")}}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
