"""Turn the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py <tag> [--launches FILE] [--bench FILE] [--reports PREFIX]      e.g.  r02
      --launches  launch-list csv inside gpurun_out/ (default launches.csv)
      --bench     bench JSON line inside gpurun_out/ (default bench.json)
      --reports   only *.ncu-rep files whose name starts with PREFIX (default: all)
Also refreshes profiles/traffic.json (DRAM bytes per launch of the GEMM kernel, read back by bench.py's roofline.traffic).

Writes profiles/<tag>_launches.csv   one forward's launch list (kernel, grid, duration) from launches.csv
       profiles/<tag>_summary.md     per-kernel totals of that forward + selected `ncu --set full` metrics of every
                                     *.ncu-rep found + the bench JSON line of the same build
"""
import collections
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum",
]


def short(name):
    name = name.replace("void vidil::<unnamed>::", "").replace("vidil::<unnamed>::", "")
    return name.split("(")[0]


def opt(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


def launches(tag, lines):
    path = os.path.join(OUT, opt("--launches", "launches.csv"))
    if not os.path.exists(path):
        return
    with open(path) as f:
        rows = [r for r in csv.DictReader([l for l in f if l.startswith('"')]) if r.get("Metric Name", "gpu__time_duration.sum") == "gpu__time_duration.sum"]
    starts = [i for i, r in enumerate(rows) if "im2col" in r["Kernel Name"]]
    if len(starts) < 2:
        return
    fwd = rows[starts[-2]:starts[-1]]
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "block", "duration_ns"])
        for r in fwd:
            w.writerow([short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], r["Metric Value"].replace(",", "")])
    agg = collections.OrderedDict()
    total = 0.0
    for r in fwd:
        k = short(r["Kernel Name"])
        ns = float(r["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    lines.append(f"## Launch list of one forward (ncu gpu__time_duration.sum, --clock-control none; {len(fwd)} launches, "
                 f"{total / 1e6:.2f} ms serialised)\n")
    lines.append("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {ns / 1e6:.3f} | {ns / n / 1e3:.1f} | {ns / total * 100:.1f}% |")
    lines.append("")


def reports(lines):
    traffic = {}
    for rep in sorted(glob.glob(os.path.join(OUT, opt("--reports", "") + "*.ncu-rep"))):
        r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(r.stdout.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units, data = rows[0], rows[1], rows[2:]
        lines.append(f"## `ncu --set full` — {os.path.basename(rep)} ({len(data)} launches)\n")
        kn = hdr.index("Kernel Name")
        lines.append("| metric | unit | " + " | ".join(f"#{i} `{short(d[kn])[:38]}`" for i, d in enumerate(data)) + " |")
        lines.append("|---|---|" + "---:|" * len(data))
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                lines.append(f"| {m} | {units[i]} | " + " | ".join(d[i] for d in data) + " |")
        lines.append("")
        # DRAM bytes per launch of the per-layer GEMMs (a capture of 4 consecutive gemm_tcgen05 launches = qkv, proj, fc1, fc2)
        if all("gemm_tcgen05_kernel" in d[kn] for d in data) and len(data) == 4 and "dram__bytes_read.sum" in hdr:
            def to_bytes(v, u):
                return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            per = [to_bytes(d[ri], units[ri]) + to_bytes(d[wi], units[wi]) for d in data]
            traffic = {"bytes_per_launch": sum(per) / len(per),
                       "source": f"profiles/{sys.argv[1]}_summary.md ({os.path.basename(rep)}): ncu --set full, dram__bytes_read.sum + "
                                 f"dram__bytes_write.sum, mean over 4 consecutive per-layer GEMM launches "
                                 f"({', '.join(f'{x / 1e6:.0f}' for x in per)} MB)"}
    if traffic:
        path = os.path.join(ROOT, "profiles", "traffic.json")
        allt = json.load(open(path)) if os.path.exists(path) else {}
        allt.setdefault("gemm_tcgen05_kernel", {})["vit_large_224_b256"] = traffic
        with open(path, "w") as f:
            json.dump(allt, f, indent=1)


def main():
    tag = sys.argv[1]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    lines = [f"# Profile summary {tag}\n"]
    bj = os.path.join(OUT, opt("--bench", "bench.json"))
    if os.path.exists(bj):
        txt = open(bj).read().strip().splitlines()
        if txt:
            b = json.loads(txt[-1])
            lines.append("## bench.py line of the same build (not under a profiler)\n")
            lines.append("```json\n" + json.dumps(b, indent=1) + "\n```\n")
    launches(tag, lines)
    reports(lines)
    with open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote profiles/%s_summary.md" % tag)


if __name__ == "__main__":
    main()
