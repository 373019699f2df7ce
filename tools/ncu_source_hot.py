"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line: stall samples, the dominant stall
reasons and executed instructions, sorted by samples.  usage: python tools/ncu_source_hot.py file.source.csv [top_n]"""
import csv
import sys


def main(path, top=40):
    rows = []
    cur_file, hdr = None, None
    with open(path, newline="") as f:
        for r in csv.reader(f):
            if not r:
                continue
            if r[0] == "File Path":
                cur_file = r[1].split("/")[-1]; continue
            if r[0] == "Function Name":
                continue
            if r[0] == "Line No":
                hdr = r; continue
            if hdr is None or r[0] == "" or not r[0].isdigit():
                continue                      # SASS rows (their samples are already summed into the CUDA line row)
            d = dict(zip(hdr, r))
            def num(k):
                try:
                    return float(d.get(k, "0"))
                except ValueError:
                    return 0.0
            stalls = {k[6:]: num(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
            rows.append((num("Warp Stall Sampling (All Samples)"), num("Instructions Executed"), cur_file, int(r[0]), r[1].strip(), stalls))
    tot = sum(x[0] for x in rows) or 1.0
    tot_i = sum(x[1] for x in rows) or 1.0
    agg = {}
    for _, _, _, _, _, st in rows:
        for k, v in st.items():
            agg[k] = agg.get(k, 0.0) + v
    print(f"total samples {tot:.0f}, warp instructions {tot_i:.0f}")
    print("stall mix: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    print(f"{'samples%':>8} {'inst%':>6}  {'file:line':<28} top stalls | source")
    for s, ins, fn, ln, src, st in sorted(rows, key=lambda x: -x[0])[:top]:
        top2 = ", ".join(f"{k} {100 * v / max(s, 1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
        print(f"{100 * s / tot:8.2f} {100 * ins / tot_i:6.2f}  {fn + ':' + str(ln):<28} {top2:<34} | {src[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
