"""Turns the ncu artefacts in gpurun_out/ into the small text summaries kept under profiles/."""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio"]


def launch_list(src, dst, title):
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    agg = collections.defaultdict(lambda: [0.0, 0])
    for r in rows[hdr + 1:]:
        if len(r) < 15:
            continue
        name = r[4].split("(")[0].replace("void ", "").replace("xv::", "")
        agg[name][0] += float(r[-1]); agg[name][1] += 1
    tot = sum(v[0] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n# gpu__time_duration.sum per launch (ncu --clock-control none; cold-cache, serialised: compare SHARES)\n")
        f.write(f"{'kernel':46s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{k:46s} {v[1]:8d} {v[0] / 1e3:12.1f} {v[0] / v[1] / 1e3:10.1f} {v[0] / tot * 100:6.1f}%\n")


def raw_metrics(rep, dst, title):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    with open(dst, "w") as f:
        f.write(f"# {title}\n# ncu --set full --clock-control none; one row per captured launch\n")
        for r in rows[2:]:
            f.write(f"\n## {r[hdr.index('Kernel Name')]}\n")
            for k in hdr:
                if k in KEYS or ("pcsamp_warps_issue_stalled" in k and "not_issued" not in k and r[hdr.index(k)] not in ("0", "")):
                    f.write(f"{k:78s} {r[hdr.index(k)]} {rows[1][hdr.index(k)]}\n")


def metrics_json(tag):
    """profiles/ncu_metrics_<tag>.json: the static per-kernel numbers bench.py quotes (instructions per MB search,
    DRAM bytes per search) computed from the captured launches, never typed in."""
    import json
    out = {}
    for kern in ("me_pass_kernel", "me_verify_kernel", "tree_chain_kernel"):
        rep = os.path.join(GP, f"{kern}_{tag}.ncu-rep")
        if not os.path.exists(rep):
            continue
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            continue
        hdr = rows[0]

        def col(r, k):
            try:
                return float(r[hdr.index(k)].replace(",", ""))
            except Exception:
                return None

        def to_bytes(r, k):
            v, unit = col(r, k), rows[1][hdr.index(k)]
            return None if v is None else v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

        launches = []
        for r in rows[2:]:
            launches.append({"grid": col(r, "launch__grid_size"), "inst": col(r, "smsp__inst_executed.sum"),
                             "dur_us": col(r, "gpu__time_duration.sum"), "issue_pct": col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                             "warps_active_pct": col(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                             "dram": (to_bytes(r, "dram__bytes_read.sum") or 0) + (to_bytes(r, "dram__bytes_write.sum") or 0),
                             "regs": col(r, "launch__registers_per_thread")})
        if kern == "me_pass_kernel":
            big = max(launches, key=lambda l: l["inst"] or 0)          # pass 0 of its batch: every MB is searched
            searches = round(big["grid"] / 1020)          # 2040 warps per 1080p search, 2 warps per block
            out[kern] = {"captured": "pass 0 of a %d-search batch at 1080p (scripts/profile_la.py, kernel alone on the GPU)" % searches,
                         "warp_inst": big["inst"], "mb_searches": searches * 8160, "warp_inst_per_mb_search": big["inst"] / (searches * 8160),
                         "duration_us": big["dur_us"], "issue_active_pct": big["issue_pct"], "warps_active_pct": big["warps_active_pct"],
                         "registers": big["regs"], "dram_bytes_per_search": big["dram"] / searches,
                         "other_passes": [{"warp_inst": l["inst"], "duration_us": l["dur_us"]} for l in launches if l is not big]}
        elif kern == "me_verify_kernel":
            l = launches[0]
            searches = round(l["grid"] / 5) if l["grid"] and l["grid"] >= 5 else 1
            out[kern] = {"captured": "verification of a batch (grid %d) at 1080p" % int(l["grid"] or 0), "warp_inst": l["inst"], "duration_us": l["dur_us"],
                         "issue_active_pct": l["issue_pct"], "registers": l["regs"], "searches": searches, "dram_bytes_per_search": l["dram"] / max(1, searches),
                         "all": [{"grid": x["grid"], "duration_us": x["dur_us"], "warp_inst": x["inst"]} for x in launches]}
        else:
            l = launches[0]
            out[kern] = {"warp_inst": l["inst"], "duration_us": l["dur_us"], "issue_active_pct": l["issue_pct"], "grid": l["grid"], "dram_bytes": l["dram"]}
    if out:
        json.dump(out, open(os.path.join(OUT, f"ncu_metrics_{tag}.json"), "w"), indent=1)
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    jobs = [("launches_bench_r1e.csv", f"launches_bench_{tag}.txt", "bench.py --steps 2 --warmup 3 --streams 2 --frames-per-step 8 --no-e2e --no-cpu-baseline (launches 3000..6000)"),
            ("launches_bench_r1.csv", f"launches_bench_{tag}_wavefront_version.txt", "same command, earlier build (plain wavefront search, one kernel per mb-tree step)"),
            ("launches_la_r1a.csv", f"launches_single_stream_{tag}.txt", "scripts/profile_la.py 10 (one 1080p stream, 10 frames)")]
    for src, dst, title in jobs:
        if os.path.exists(os.path.join(GP, src)):
            launch_list(os.path.join(GP, src), os.path.join(OUT, dst), title)
    reps = [("me_pass_d.ncu-rep", f"ncu_me_pass_{tag}.txt", "me_pass_kernel<1,true>, pass 0 of a 2-search batch (16320 MB searches), alone on the GPU"),
            ("me_pass_a.ncu-rep", f"ncu_me_pass_{tag}_first_version.txt", "me_pass_kernel before the state-machine / code-size work (129 KB of SASS, 167 registers)"),
            ("me_verify_e.ncu-rep", f"ncu_me_verify_{tag}.txt", "me_verify_kernel<true>, verification wavefront of a 2-search batch, alone on the GPU"),
            ("me2_b.ncu-rep", f"ncu_me_band_{tag}_rejected.txt", "rejected design: 4 MB rows per warp in lockstep (me_band_kernel), 2 searches"),
            ("me_prof_r1c.ncu-rep", f"ncu_me_wavefront_{tag}.txt", "me_wavefront_kernel (plain wavefront, now the fallback), 7 searches of one 1080p frame, alone on the GPU"),
            ("me_prof_r1a.ncu-rep", f"ncu_me_wavefront_{tag}_first_version.txt", "first version of the search kernel (global loads, progress flags) for comparison"),
            ("s1_bgra_r1a.ncu-rep", f"ncu_csp_bgra_{tag}_before_4row.txt", "rgb_to_420 (dp4a version, before the 4-row fast path), 64 frames per launch"),
            ("s1_lowres_r1a.ncu-rep", f"ncu_lowres_{tag}_before.txt", "lowres_init before the cheap-border rewrite, 64 frames per launch"),
            ("s1_bgra_r1b.ncu-rep", f"ncu_csp_bgra_{tag}.txt", "rgb_to_420_fast_kernel<4,false>, 64 frames (730 MB algorithmic) per launch"),
            ("s1_lowres_r1b.ncu-rep", f"ncu_lowres_{tag}.txt", "lowres_init_kernel, 64 frames per launch"),
            ("hpel_a.ncu-rep", f"ncu_hpel_{tag}_v1_unrolled.txt", "hpel_kernel, first version (row loop unrolled by six: 57 KB of SASS, a row's load consumed in the trip that issues it), 48 frames of 1920x1088 per launch"),
            ("hpel_b.ncu-rep", f"ncu_hpel_{tag}_v2_prefetch_unrolled.txt", "hpel_kernel with rows prefetched two trips ahead, still unrolled by six (62 KB of SASS): the stall moves from the load to the instruction cache"),
            ("hpel_c.ncu-rep", f"ncu_hpel_{tag}_v3_compact_loop.txt", "hpel_kernel, one row per loop trip (6 KB loop body), one warp per block; an EXIT inside the loop still waits for the prefetched loads"),
            ("hpel_e.ncu-rep", f"ncu_hpel_{tag}_v4_no_exit.txt", "hpel_kernel without the EXIT in the loop, register double buffer: the copy pre[0]=pre[1] of a register still being loaded holds 25 % of the stall samples"),
            ("hpel_f.ncu-rep", f"ncu_hpel_{tag}.txt", "hpel_kernel as measured at the end of the round: rows prefetched three trips ahead through a per-lane cp.async ring in shared memory; 48 frames of 1920x1088 per launch")]
    if tag != "r1":
        jobs = [(f"launches_bench_{tag}.csv", f"launches_bench_{tag}.txt", "bench.py --steps 2 --warmup 3 --streams 2 --frames-per-step 24 --clip-frames 60 --no-e2e --no-cpu-baseline --no-worst-case (launches 4000..8000)")]
        for src, dst, title in jobs:
            if os.path.exists(os.path.join(GP, src)):
                launch_list(os.path.join(GP, src), os.path.join(OUT, dst), title)
        reps = [(f"{k}_{tag}.ncu-rep", f"ncu_{k.replace('_kernel', '')}_{tag}.txt", f"{k}, scripts/profile_la.py 60 (one 1080p stream, preset medium), kernel alone on the GPU")
                for k in ("me_pass_kernel", "me_verify_kernel", "tree_chain_kernel", "frontend_kernel", "intra_kernel", "finalize_kernel")]
        reps.append((f"dec_packed_kernel_{tag}.ncu-rep", f"ncu_dec_packed_{tag}.txt",
                     "dec_packed_kernel<BGRA,vec>, yuv420p -> bottom-up RGB32, 48 pictures of 1080p per launch (scripts/probe_decode.py 48 2 bgra_bottom_up), kernel alone on the GPU"))
        print(metrics_json(tag))
    for src, dst, title in reps:
        if os.path.exists(os.path.join(GP, src)):
            raw_metrics(os.path.join(GP, src), os.path.join(OUT, dst), title)
    print(sorted(os.listdir(OUT)))
