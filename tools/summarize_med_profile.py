"""profiles/<tag>_generate_launches.csv (compact: kernel, grid, block, duration_ns) and profiles/<tag>_summary.md for the caption
decoder / CapFilt path, from the scratch files a gpurun session left in gpurun_out/:

    python tools/summarize_med_profile.py r01e gpurun_out/med_launches_e.csv gpurun_out/capfilt_d.log [more bench logs...]

Every gpurun_out/<tag>_*.ncu-rep (`ncu --set full`) contributes a row of selected metrics.
"""
import collections
import csv
import glob
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "smsp__inst_executed.sum"]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void vidil::<unnamed>::", "").replace("vidil::<unnamed>::", "")


def read_launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ns = v if unit == "ns" else v * 1e3 if unit in ("us", "usecond") else v * 1e6
        out.append((short(row["Kernel Name"]), row["Grid Size"], row["Block Size"], int(round(ns))))
    return out


def main():
    tag, launches_csv, logs = sys.argv[1], sys.argv[2], sys.argv[3:]
    L = read_launches(launches_csv)
    with open(os.path.join(ROOT, "profiles", f"{tag}_generate_launches.csv"), "w") as f:
        f.write("kernel,grid,block,duration_ns\n")
        w = csv.writer(f)
        for r in L:
            w.writerow(r)
    agg = collections.defaultdict(lambda: [0, 0])
    for k, g, b, ns in L:
        agg[(k, g)][0] += 1
        agg[(k, g)][1] += ns
    total = sum(ns for *_, ns in L)
    md = [f"# Profile summary {tag}: caption decoder (vidil_med_generate) and CapFilt stages", ""]
    md += ["## bench.py --workload capfilt lines of this round (not under a profiler), oldest first", "", "```json"]
    for p in logs:
        for line in open(p):
            line = line.strip()
            if line.startswith("{"):
                d = json.loads(line)
                d.pop("config", None)
                md.append(json.dumps(d))
    md += ["```", "", f"## Launch list of one `vidil_med_generate` call ({os.path.basename(launches_csv)}: tools/med_profile.py, 1024 frames x 197 "
           "ViT-L tokens, beams 3, max_length 20; `ncu --metrics gpu__time_duration.sum --clock-control none`, serialised and cold-cache)", "",
           f"total {total / 1e6:.2f} ms over {len(L)} launches", "", "| kernel | grid | launches | total ms | share | mean us |", "|---|---|---:|---:|---:|---:|"]
    for (k, g), (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        md.append(f"| `{k}` | {g} | {c} | {ns / 1e6:.2f} | {100 * ns / total:.1f}% | {ns / c / 1e3:.1f} |")
    md += ["", "## `ncu --set full --clock-control none` captures", ""]
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_*.ncu-rep"))):
        r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        md.append(f"### {os.path.basename(rep)}")
        md.append("")
        for row in rows[2:]:
            name = short(row[hdr.index("Kernel Name")])
            md.append(f"* `{name}` grid {row[hdr.index('launch__grid_size')] if 'launch__grid_size' in hdr else '?'}")
            for m in METRICS:
                if m in hdr:
                    md.append(f"  * {m} = {row[hdr.index(m)]} {units[hdr.index(m)]}")
        md.append("")
    with open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w") as f:
        f.write("\n".join(md) + "\n")
    print("\n".join(md[:60]))


if __name__ == "__main__":
    main()
