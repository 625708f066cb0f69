#!/usr/bin/env python3
"""Turns gpurun_out ncu artefacts into the committed summaries under profiles/.
usage: summarize_ncu.py <tag> <launches.csv> <prof.ncu-rep | -> [workload] [rays_per_frame]   ("-": launch list only)
With rays_per_frame the per-launch counters of the frame kernels (executed warp instructions, L2 and DRAM bytes) are
also written to ncu_counts.json, which bench.py reads for the issue / L2 / HBM terms of its roofline."""
import collections, csv, json, os, subprocess, sys
tag, launches, rep = sys.argv[1:4]
workload = sys.argv[4] if len(sys.argv) > 4 else "synth1m"
rays_per_frame = float(sys.argv[5]) if len(sys.argv) > 5 else None
here = os.path.dirname(os.path.abspath(__file__))
# ---- launch list -> per-kernel shares
rows = list(csv.reader(open(launches)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("nsecond", "ns") else v * 1e3 if r[ui] in ("msecond", "ms") else v
    a = agg.setdefault(r[ki][:100], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(here, f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none), bench.py --steps 5 --warmup 3 --no-cpu --no-others\n\n")
    f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f}% |\n")
if rep == "-":
    print(open(os.path.join(here, f"{tag}_launches.md")).read()); sys.exit(0)
# ---- full capture -> key metrics + DRAM traffic
# (a .ncu-rep, or the `ncu -i rep --page raw --csv` dump of one made on the GPU box: the reports themselves are too
# large to bring back)
raw = open(rep).read() if rep.endswith(".csv") else \
    subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]; idx = {x: i for i, x in enumerate(hdr)}
keep = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
units = rows[1]
def mb(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
traffic = {}
counts = {"inst_executed": 0.0, "lts_bytes": 0.0, "dram_bytes": 0.0, "kernels": [], "rays": rays_per_frame,
          "source": f"profiles/{tag}_ncu_full.md (ncu --set full, one launch per kernel of a frame)"}
FRAME_KERNELS = ("k_frame", "k_traverse", "k_shade", "k_paths", "k_resolve")
with open(os.path.join(here, f"{tag}_ncu_full.md"), "w") as f:
    f.write(f"# {tag}: ncu --set full --clock-control none (one launch per kernel), {workload}\n\n")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        f.write(f"## `{name}`\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in keep:
            if k in idx: f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
        f.write("\n")
        # the first launch of each frame kernel is the whole 1-GPU frame (tests/gpu_profile_r2.py profiles a rank's
        # share of an 8-GPU frame afterwards: same kernels again, or k_frame_push)
        short_name = name.split("(")[0]
        if any(k in name for k in FRAME_KERNELS) and "k_frame_push" not in name and short_name not in counts["kernels"] \
                and "smsp__inst_executed.sum" in idx:
            counts["inst_executed"] += float(r[idx["smsp__inst_executed.sum"]].replace(",", ""))
            if "lts__t_sectors.sum" in idx:
                counts["lts_bytes"] += 32.0 * float(r[idx["lts__t_sectors.sum"]].replace(",", ""))
            counts["dram_bytes"] += mb(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                                    mb(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            counts["kernels"].append(short_name)
        short = "k_traverse<primary>" if "k_traverse<0" in name else "k_traverse<queue>" if "k_traverse<1" in name else \
                "k_traverse<shadow>" if "k_traverse<2" in name else \
                "k_shade" if "k_shade" in name else "k_resolve" if "k_resolve" in name else name.split("(")[0].split("::")[-1]
        traffic[short] = mb(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                         mb(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
tp = os.path.join(here, "traffic.json")
tj = json.load(open(tp)) if os.path.exists(tp) else {}
tj[workload] = traffic
tj["_source"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/{tag}_ncu_full.md"
json.dump(tj, open(tp, "w"), indent=1, sort_keys=True)
if rays_per_frame:
    cp = os.path.join(here, "ncu_counts.json")
    cj = json.load(open(cp)) if os.path.exists(cp) else {}
    cj[workload] = counts
    json.dump(cj, open(cp, "w"), indent=1, sort_keys=True)
print(open(os.path.join(here, f"{tag}_launches.md")).read())
print(json.dumps(traffic))
print(json.dumps(counts))
