"""Executed-instruction mix by opcode from an .ncu-rep source page (first kernel)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
mix = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    if r[0] == "Kernel Name": break
    src = r[ci["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    try:
        n = int(r[ci["Instructions Executed"]] or 0)
    except ValueError:
        break
    mix[op] += n; tot += n
print("total warp instructions", tot)
for op, n in mix.most_common(40):
    print(f"{op:10s} {n:12d} {100*n/tot:5.1f}%")
