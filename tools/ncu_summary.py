"""Summarise ncu artefacts from gpurun_out/ into small tracked text files under profiles/.
usage: python tools/ncu_summary.py <tag>"""
import csv, glob, os, subprocess, sys, collections
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = open(os.path.join(root, "profiles", f"{tag}_summary.md"), "w")
lf = os.path.join(root, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lf):
    rows = [r for r in csv.reader(open(lf)) if len(r) > 5]
    hdr = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
if os.path.exists(lf) and hdr is not None:
    H = rows[hdr]; kn, mv = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        try: t = float(r[mv].replace(",", ""))
        except ValueError: continue
        a = agg.setdefault(r[kn].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += t
    tot = sum(a[1] for a in agg.values())
    unit = rows[hdr + 1][H.index("Metric Unit")]
    out.write(f"# ncu launch list `{tag}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
    out.write(f"| kernel | launches | total {unit} | avg {unit} | share |\n|---|---|---|---|---|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"| `{k}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / tot:.1f}% |\n")
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
for rep in sorted(glob.glob(os.path.join(root, "gpurun_out", f"prof_{tag}_*.ncu-rep"))):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3: continue
    H, U = rows[0], rows[1]
    out.write(f"\n## `ncu --set full` capture: {os.path.basename(rep)}\n\n")
    for r in rows[2:3]:
        out.write(f"kernel: `{r[H.index('Kernel Name')]}`\n\n| metric | value | unit |\n|---|---|---|\n")
        for w in want:
            if w in H: out.write(f"| {w} | {r[H.index(w)]} | {U[H.index(w)]} |\n")
out.close()
print(open(out.name).read())
