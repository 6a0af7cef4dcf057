"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X.csv):
   python tools/launch_summary.py X.csv "command that was profiled" [steps_in_run]
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import csv
import re
import sys
from collections import OrderedDict

STEP_KERNELS = ["update_kernel", "ordered_sum_kernel", "extract_weights_kernel", "tile_sum_kernel", "tile_fn_kernel",
                "resolve_kernel", "emit_kernel", "exact_scan_kernel", "normalize_kernel", "resample_search_kernel",
                "mean_pass1", "mean_pass2"]


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("a3d::<unnamed>::", "").replace("a3d::", "")
    return name[-72:]


def main():
    path, cmd = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "?"
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1e6
    total = sum(a[1] for a in agg.values())
    print(f"# ncu launch list summary: `{cmd}` (B200)")
    print(f"# source: {path} (ncu --metrics gpu__time_duration.sum --clock-control none)")
    print("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:<72} launches={n:4d} total={ms:9.3f} ms  mean={ms / n:8.4f} ms  share={100 * ms / total:5.1f}%")
    # one step of the hot path: kernels of STEP_KERNELS, launches per step inferred from the update kernel count
    upd = next((v for k, v in agg.items() if k.startswith("update_kernel")), None)
    if upd:
        steps = upd[0]
        print(f"\n# per step (launch counts divided by the {steps} update launches of the run)")
        tot = 0.0
        for want in STEP_KERNELS:
            for k, (n, ms) in agg.items():
                if want in k:
                    per = ms / steps
                    tot += per
                    print(f"  {k:<70} {n / steps:5.2f} launches  {per:8.4f} ms")
        print(f"  per step total (sum of kernel times under ncu) {tot:.4f} ms")
        for k, (n, ms) in agg.items():
            if k.startswith("update_kernel") or "ordered_sum" in k:
                print(f"  share of step: {k:<50} {100 * ms / steps / tot:5.1f}%")


if __name__ == "__main__":
    main()
