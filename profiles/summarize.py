#!/usr/bin/env python3
"""Turn the ncu outputs a gpurun call left in gpurun_out/ into the tracked summary under profiles/.

  python profiles/summarize.py r01            # reads gpurun_out/r01_launches.csv + gpurun_out/r01_full*.ncu-rep

Writes profiles/<round>_launches.csv (copy of the launch list), profiles/<round>_summary.md (tables) and
profiles/<round>_metrics.json (per-kernel means of the raw-page metrics, used for bench.py's roofline.traffic).
Needs `ncu` on PATH to read the .ncu-rep (no GPU needed).
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PUSH_KERNELS = ("k_classify", "k_update", "k_borders")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__block_size"]
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def kname(full: str, keep_template: bool = False) -> str:
    """'void (anonymous namespace)::k_axis<0>(AxisParams)' -> 'k_axis' (or 'k_axis<0>')"""
    s = full.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    s = s.split("(")[0]
    return s if keep_template else s.split("<")[0]


def launch_table(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    per = collections.defaultdict(list)
    for r in rows:
        per[kname(r[4])].append(float(r[14]) / 1e3)
    tot = sum(sum(v) for v in per.values())
    out.append("| kernel | launches | mean us | min us | max us | share of listed GPU time |")
    out.append("|---|---|---|---|---|---|")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| {k} | {len(v)} | {sum(v)/len(v):.1f} | {min(v):.1f} | {max(v):.1f} | {100*sum(v)/tot:.1f} % |")
    # steady-state share inside one push: medians, so the first (cold) launches do not dominate
    med = {k: sorted(per[k])[len(per[k]) // 2] for k in PUSH_KERNELS if k in per and len(per[k]) >= 10}
    ptot = sum(med.values())
    if ptot:
        out.append("\nMedian launch inside one push: " + ", ".join(f"{k} {med[k]:.1f} us ({100*med[k]/ptot:.0f} %)" for k in med) + ".")
    return per


def raw_tables(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rr[2:]:
        k = kname(r[idx["Kernel Name"]], keep_template=True)
        for w in WANT:
            if w in idx:
                try:
                    v = float(r[idx[w]].replace(",", ""))
                except ValueError:
                    continue
                u = units[idx[w]]
                if u in BYTES:
                    v *= BYTES[u]
                if w == "gpu__time_duration.sum":
                    v *= TIME_US.get(u, 1.0)
                agg[k][w].append(v)

    def m(k, w):
        v = agg[k].get(w, [])
        return sum(v) / len(v) if v else float("nan")

    out.append("| kernel | n | time us | dram read MB | dram write MB | DRAM % of peak | SM % | regs | warps active % | warp instr | IPC/SM | L2 hit % | L1 hit % | grid x block |")
    out.append("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for k in agg:
        out.append(f"| {k} | {len(agg[k]['gpu__time_duration.sum'])} | {m(k,'gpu__time_duration.sum'):.1f} | "
                   f"{m(k,'dram__bytes_read.sum')/1e6:.2f} | {m(k,'dram__bytes_write.sum')/1e6:.2f} | "
                   f"{m(k,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {m(k,'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                   f"{m(k,'launch__registers_per_thread'):.0f} | {m(k,'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
                   f"{m(k,'smsp__inst_executed.sum'):.3g} | {m(k,'sm__inst_executed.avg.per_cycle_elapsed'):.2f} | "
                   f"{m(k,'lts__t_sector_hit_rate.pct'):.1f} | {m(k,'l1tex__t_sector_hit_rate.pct'):.1f} | "
                   f"{m(k,'launch__grid_size'):.0f} x {m(k,'launch__block_size'):.0f} |")
    return {k: {w: m(k, w) for w in WANT} for k in agg}


def stall_table(rep, kernel, out, top=14):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{kernel}", "-c", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = None
    for i, r in enumerate(rows):
        if "stall_long_sb" in r:
            hdr, start = r, i + 1
            break
    if hdr is None:
        return
    stall = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    tot = {c: 0 for c in stall}
    iE = hdr.index("Instructions Executed")
    lines = collections.OrderedDict()
    cur = None
    for r in rows[start:]:
        if len(r) < len(hdr):
            continue
        if r[0].isdigit():
            cur = (int(r[0]), r[1].strip()[:90])
            continue
        if cur is None:
            continue
        a = lines.setdefault(cur, [0, 0])
        try:
            a[0] += int(r[iE])
        except ValueError:
            pass
        for c in stall:
            try:
                v = int(r[hdr.index(c)])
            except ValueError:
                continue
            tot[c] += v
            a[1] += v
    s = sum(tot.values()) or 1
    out.append(f"\n### {kernel}: warp-stall samples (one launch)\n")
    out.append(", ".join(f"{c[6:]} {100*v/s:.1f} %" for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]) + ".\n")
    ti = sum(v[0] for v in lines.values()) or 1
    out.append("| line | share of warp instr | stall samples | source |")
    out.append("|---|---|---|---|")
    for (ln, text), (e, st) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        out.append(f"| {ln} | {100*e/ti:.1f} % | {st} | `{text.replace('|', '/')}` |")


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    tag = sys.argv[2] if len(sys.argv) > 2 else ""
    go = os.path.join(ROOT, "gpurun_out")
    import glob
    launches = os.path.join(go, f"{rnd}_launches.csv")
    reps = sorted(glob.glob(os.path.join(go, f"{rnd}_full*.ncu-rep")))
    out = [f"# {rnd}: ncu summary\n"]
    if os.path.exists(launches):
        shutil.copy(launches, os.path.join(ROOT, "profiles", f"{rnd}_launches.csv"))
        out.append("## Launch list\n")
        out.append("`ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sweep`"
                   " (the whole bench: map build, pushes, ray casts, ICP, map publication, 10^5-hypothesis scoring)."
                   " Per-launch times are cold-cache and serialised: compare shares, not absolutes. Full list: `" + f"{rnd}_launches.csv`.\n")
        launch_table(launches, out)
    metrics = {}
    for rep in reps:
        out.append(f"\n## ncu --set full, {os.path.basename(rep)} (per launch, mean over the captured launches)\n")
        m = raw_tables(rep, out)
        for k in m:
            if k.split("<")[0] in ("k_update", "k_raycast", "k_icp", "k_classify", "k_score_tsd", "k_score_rnm", "k_score_pdf"):
                stall_table(rep, k.split("<")[0], out)
        metrics.update({k.split("<")[0]: v for k, v in m.items()})
        if tag:  # the workload the captured bench ran (bench.py's roofline.traffic looks up "k_update@<workload>")
            metrics.update({k.split("<")[0] + "@" + tag: v for k, v in m.items()})
    if metrics:
        json.dump(metrics, open(os.path.join(ROOT, "profiles", f"{rnd}_metrics.json"), "w"), indent=1)
    extra = os.path.join(ROOT, "profiles", f"{rnd}_notes.md")
    if os.path.exists(extra):
        out.append("\n" + open(extra).read())
    open(os.path.join(ROOT, "profiles", f"{rnd}_summary.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
