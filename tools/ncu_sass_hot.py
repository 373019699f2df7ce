"""Top SASS instructions by stall samples of an `ncu --page source --csv --print-source cuda,sass` export, with the CUDA line each
belongs to.  usage: python tools/ncu_sass_hot.py file.source.csv [top_n]"""
import csv
import sys


def main(path, top=25):
    rows, cur = [], None
    for r in csv.reader(open(path, newline="")):
        if not r or r[0] in ("File Path", "Function Name", "Line No"):
            continue
        if r[0].isdigit():
            cur = (r[0], r[1].strip()[:80]); continue
        if r[0] == "" and len(r) > 6 and r[2].startswith("0x"):
            try:
                rows.append((float(r[4]), r[2], r[3].strip()[:60], cur))
            except ValueError:
                pass
    # an instruction is listed once under every CUDA line it is attributed to (inlining): keep the first
    seen, uniq = set(), []
    for x in rows:
        if x[1] in seen:
            continue
        seen.add(x[1]); uniq.append(x)
    tot = sum(x[0] for x in uniq) or 1.0
    for s, addr, ins, c in sorted(uniq, key=lambda x: -x[0])[:top]:
        print(f"{100 * s / tot:6.2f}  {ins:<60} <- {c[0]}: {c[1]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
