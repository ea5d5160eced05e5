#!/usr/bin/env python
"""Turn the ncu outputs a gpurun call left in gpurun_out/ into the small tracked summaries in
profiles/ (the .ncu-rep itself is scratch).

    python profiles/summarize.py launches gpurun_out/launches_c4.csv profiles/r1_c4_launches.csv
    python profiles/summarize.py full gpurun_out/prof_density.ncu-rep profiles/r1_c4_density_full.csv
    python profiles/summarize.py stalls gpurun_out/prof_density.ncu-rep profiles/r1_c4_density_stalls.txt
"""
import collections
import csv
import io
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"]
        if not (name.startswith("k_") or " k_" in name[:12]):
            continue          # only this repo's kernels (torch's synthetic-data kernels dropped)
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}[row["Metric Unit"]]
        short = name.split("(")[0].replace("void ", "")
        a = agg.setdefault(short, [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "min_us", "max_us", "share_pct"])
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            w.writerow([k, a[0], f"{a[1]:.1f}", f"{a[1] / a[0]:.2f}", f"{a[2]:.2f}", f"{a[3]:.2f}",
                        f"{100 * a[1] / tot:.1f}"])
        w.writerow(["TOTAL (this repo's kernels; ncu serialises launches, cold caches)",
                    sum(a[0] for a in agg.values()), f"{tot:.1f}", "", "", "", "100.0"])


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def full(src, dst):
    r = raw(src)
    hdr, units = r[0], r[1]
    cols = ["Kernel Name"] + [m for m in FULL_METRICS if m in hdr]
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[hdr.index(c)] for c in cols])
        for row in r[2:]:
            w.writerow([row[hdr.index(c)][:90] for c in cols])


def stalls(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    i_src, i_s, i_ie = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    sc = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break                       # first captured launch only
        if len(r) < len(hdr) or not r[i_s].isdigit():
            continue
        data.append((int(r[i_s]), r))
    tot = sum(s for s, _ in data) or 1
    agg = collections.Counter()
    for s, r in data:
        for h, i in sc:
            if r[i].isdigit():
                agg[h] += int(r[i])
    with open(dst, "w") as f:
        f.write(f"{rows[0][1]}\nwarp-state samples: {tot}\n\nstall reason        samples   share\n")
        for h, v in agg.most_common(10):
            f.write(f"{h:22s} {v:6d}  {100 * v / tot:5.1f}%\n")
        f.write("\ntop instructions by samples (samples, times executed, SASS, top reasons)\n")
        for s, r in sorted(data, key=lambda x: -x[0])[:20]:
            top = sorted(((int(r[i]) if r[i].isdigit() else 0, h) for h, i in sc), reverse=True)[:2]
            f.write(f"{s:6d} {r[i_ie]:>9s}  {r[i_src][:60]:60s} {top}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full, "stalls": stalls}[sys.argv[1]](sys.argv[2], sys.argv[3])
