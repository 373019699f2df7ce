"""Summarise the round-2 ncu artefacts (raw-page CSV exports of `ncu --set full` captures + the launch list) from gpurun_out/ into
profiles/<tag>_summary.md and profiles/ncu_traffic.json entries.   usage: python tools/ncu_csv_summary.py r02p"""
import collections, csv, glob, json, os, sys
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = open(os.path.join(root, "profiles", f"{tag}_summary.md"), "w")
lf = os.path.join(root, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lf):
    rows = [r for r in csv.reader(open(lf)) if len(r) > 5]
    hdr = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
    if hdr is not None:
        H = rows[hdr]; kn, mv = H.index("Kernel Name"), H.index("Metric Value")
        agg = collections.OrderedDict()
        for r in rows[hdr + 1:]:
            try: t = float(r[mv].replace(",", ""))
            except ValueError: continue
            a = agg.setdefault(r[kn].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += t
        tot = sum(a[1] for a in agg.values())
        unit = rows[hdr + 1][H.index("Metric Unit")]
        out.write(f"# ncu launch list `{tag}`: `python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline` (C2), 2500 launches inside the timed gradient step\n\n"
                  f"(gpu__time_duration.sum, --clock-control none; cold-cache, serialised replay: compare SHARES, not absolutes)\n\n")
        out.write(f"| kernel | launches | total {unit} | avg {unit} | share |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write(f"| `{k}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / tot:.1f}% |\n")
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__average_warp_latency_issue_stalled_barrier.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
traffic = {}
for f in sorted(glob.glob(os.path.join(root, "gpurun_out", f"prof_{tag}_*.raw.csv"))):
    rows = list(csv.reader(open(f)))
    if len(rows) < 3: continue
    H, U, r = rows[0], rows[1], rows[2]
    name = os.path.basename(f)[len(f"prof_{tag}_"):-len(".raw.csv")]
    out.write(f"\n## `ncu --set full --clock-control none` capture: {name}\n\nkernel: `{r[H.index('Kernel Name')][:150]}`\n\n| metric | value | unit |\n|---|---|---|\n")
    for w in want:
        if w in H: out.write(f"| {w} | {r[H.index(w)]} | {U[H.index(w)]} |\n")
    def val(k):
        v = float(r[H.index(k)].replace(",", "")); u = U[H.index(k)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    try:
        traffic[name] = {"dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"), "ms": float(r[H.index("gpu__time_duration.sum")].replace(",", "")) *
                         {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(U[H.index("gpu__time_duration.sum")], 1)}
    except Exception as e:
        traffic[name] = {"error": str(e)}
out.close()
json.dump(traffic, open(os.path.join(root, "profiles", f"{tag}_traffic.json"), "w"), indent=1)
print(open(out.name).read()[:6000]); print(json.dumps(traffic, indent=1))
