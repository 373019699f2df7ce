"""Top stall sites of an ncu --set full capture (SASS level).  usage: python tools/ncu_stalls.py rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
H = rows[1]
iS, iN = H.index("Source"), H.index("# Samples")
stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
body = rows[2:]
tot = sum(int(r[iN]) for r in body)
agg = {H[i]: sum(int(r[i]) for r in body) for i in stall_cols}
print("total samples", tot)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
order = sorted(range(len(body)), key=lambda k: -int(body[k][iN]))[:N]
for k in sorted(order):
    r = body[k]
    st = {H[i][6:]: int(r[i]) for i in stall_cols if int(r[i]) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{k:5d} {int(r[iN]):6d} {100*int(r[iN])/tot:5.1f}%  {r[iS].strip()[:70]:70s} {top}")
