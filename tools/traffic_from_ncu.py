"""Per-kernel DRAM traffic and duration from an ncu report: python tools/traffic_from_ncu.py rep.ncu-rep "C3 T 100000x10000" [skip_launches]
Prints a JSON entry {key: {...}} to merge into profiles/r2_traffic.json (what bench.py reads as `traffic`)."""
import csv
import io
import json
import subprocess
import sys

rep, key = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
units = rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v = r[col[name]].replace(",", "")
    u = units[col[name]]
    f = float(v)
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "usecond": 1e-3, "msecond": 1, "nsecond": 1e-6, "second": 1e3}.get(u, 1)
    return f * scale


per = {}
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[col["Kernel Name"]].split("(")[0].split("::")[-1]
    per.setdefault(name, []).append(r)
entry = {"kernels": {}, "source": f"ncu --set full --clock-control none, {rep}"}
total = 0.0
for name, rs in per.items():
    rs = rs[skip:] if len(rs) > skip else rs
    rd = sum(val(r, "dram__bytes_read.sum") for r in rs) / len(rs)
    wr = sum(val(r, "dram__bytes_write.sum") for r in rs) / len(rs)
    ms = sum(val(r, "gpu__time_duration.sum") for r in rs) / len(rs)
    entry["kernels"][name] = {"launches_averaged": len(rs), "dram_bytes_read": rd, "dram_bytes_write": wr, "duration_ms": ms}
    total += rd + wr
entry["per_update_total_dram_bytes"] = total
print(json.dumps({key: entry}, indent=1))
