"""Top stalled SASS instructions from an .ncu-rep source page: python tools/ncu_hot.py file.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# first kernel only: rows[0] = kernel name, rows[1] = header
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == "Kernel Name" or r[0] == "Address":
        if r and r[0] == "Kernel Name":
            break
        continue
    data.append(r)
S = ci["# Samples"]
tot = sum(int(r[S] or 0) for r in data)
print("instructions", len(data), "total samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ci[h]] or 0) for r in data) for h in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(data)), key=lambda k: -int(data[k][S] or 0))[:n]
for k in sorted(order):
    r = data[k]
    top = sorted(stalls, key=lambda h: -int(r[ci[h]] or 0))[:2]
    print(f"{k:5d} {int(r[S]):7d} {100*int(r[S])/tot:5.1f}%  exec={r[ci['Instructions Executed']]:>9s} thr={r[ci['Avg. Threads Executed']]:>5s} "
          f"{r[ci['Source']].strip()[:70]:70s} {top[0]}={r[ci[top[0]]]} {top[1]}={r[ci[top[1]]]}")
