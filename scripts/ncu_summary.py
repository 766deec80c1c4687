"""Summarise an ncu launch list (csv) and/or a full .ncu-rep into a markdown file under profiles/.

    python scripts/ncu_summary.py --launches gpurun_out/launches.csv --rep gpurun_out/prof.ncu-rep --out profiles/x.md --title "..."
"""
import argparse, csv, subprocess, io, collections

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = collections.OrderedDict()
    for r in rows[1:]:
        try:
            d.setdefault(r[ki].split("(")[0][-60:], []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in d.values())
    out = ["| kernel | launches | total ms | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k}` | {len(v)} | {sum(v)/1e6:.3f} | {sum(v)/len(v)/1e3:.1f} | {sum(v)/tot:.3f} |")
    return "\n".join(out)


def ncu_csv(rep, page, extra=()):
    r = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True)
    return list(csv.reader(io.StringIO(r.stdout)))


def raw_metrics(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    out = []
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        out.append(f"\n**`{name}`**\n")
        out.append("| metric | value | unit |\n|---|---:|---|")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                out.append(f"| {m} | {r[i]} | {units[i]} |")
    return "\n".join(out), sorted(seen)


def stalls(rep, kernel):
    rows = ncu_csv(rep, "source", ("--kernel-name", "regex:" + kernel))
    if len(rows) < 3:
        return ""
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = dict.fromkeys(names, 0)
    lines = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            n = int(r[idx["# Samples"]])
        except ValueError:
            continue
        for s in names:
            try:
                tot[s] += int(r[idx[s]])
            except ValueError:
                pass
        lines.append((n, r[idx["Source"]][:80]))
    T = max(sum(tot.values()), 1)
    S = max(sum(n for n, _ in lines), 1)
    out = ["stall reasons (share of samples): " + ", ".join(f"{k[6:]} {v/T:.2f}" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v / T >= 0.01)]
    out.append("\nhottest SASS lines:\n")
    out.append("| share | instruction |\n|---:|---|")
    for n, src in sorted(lines, key=lambda x: -x[0])[:8]:
        out.append(f"| {n/S:.3f} | `{src.strip()}` |")
    return "\n".join(out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches"); ap.add_argument("--rep"); ap.add_argument("--out", required=True); ap.add_argument("--title", default="ncu summary")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    md = [f"# {a.title}\n", a.note, ""]
    if a.launches:
        md += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)\n", launches(a.launches), ""]
    if a.rep:
        txt, kernels = raw_metrics(a.rep)
        md += ["## `ncu --set full` (one launch per kernel)\n", txt, ""]
        for k in kernels:
            short = k.split("::")[-1].split("<")[0].replace("void ", "").strip()
            md += [f"### stall profile of `{short}`\n", stalls(a.rep, short), ""]
    open(a.out, "w").write("\n".join(md))
    print("wrote", a.out)
