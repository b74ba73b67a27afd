#!/usr/bin/env python
"""ncu launch list (CSV, `--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`) of ONE bench step ->
per-kernel-function summary: launches, total / average duration, share of the step, average DRAM bytes per launch.

    python scripts/ncu_launch_summary.py gpurun_out/launches.csv profiles/r01_launch_summary.json > profiles/r01_launch_summary.txt

bench.py reads the JSON for `roofline.traffic` (measured DRAM bytes per launch of the dominant kernel)."""
import collections
import csv
import json
import re
import sys


def short(name):
    m = re.search(r"(\w+)(<[^>]*>)?\(", name)
    base = m.group(1) if m else name
    tpl = re.search(base + r"<([^>]*)>", name)
    return base + ("<" + tpl.group(1).replace("(int)", "").replace("__nv_bfloat16", "bf16") + ">" if tpl else "")


def main():
    src, dst = sys.argv[1], sys.argv[2]
    per = collections.defaultdict(lambda: dict(launches=0, ns=0.0, rd=0.0, wr=0.0))
    rows = {}
    with open(src) as f:
        for line in f:
            if line.startswith('"ID"'):
                break
        for r in csv.reader(f):
            if len(r) < 15:
                continue
            kid, name, metric, unit, val = int(r[0]), r[4], r[12], r[13], float(r[14].replace(",", ""))
            rows.setdefault(kid, dict(name=name))[metric] = (val, unit)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}
    for k in rows.values():
        p = per[short(k["name"])]
        p["launches"] += 1
        for key, fld in (("gpu__time_duration.sum", "ns"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr")):
            if key in k:
                v, u = k[key]
                p[fld] += v * scale.get(u, 1.0)
    tot = sum(p["ns"] for p in per.values()) or 1.0
    out = {}
    print(f"{'kernel':<44} {'launches':>8} {'total ms':>9} {'share':>7} {'avg us':>9} {'DRAM MB/launch':>15}")
    for name, p in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
        out[name] = dict(launches=p["launches"], total_ms=p["ns"] / 1e6, share=p["ns"] / tot, avg_us=p["ns"] / 1e3 / p["launches"],
                         dram_bytes_per_launch=(p["rd"] + p["wr"]) / p["launches"])
        print(f"{name:<44} {p['launches']:>8} {p['ns'] / 1e6:>9.3f} {100 * p['ns'] / tot:>6.1f}% {p['ns'] / 1e3 / p['launches']:>9.2f} "
              f"{(p['rd'] + p['wr']) / p['launches'] / 1e6:>15.2f}")
    print(f"total {tot / 1e6:.3f} ms over {sum(p['launches'] for p in per.values())} launches (ncu: cold caches, serialised)")
    json.dump({"source": src, "note": "ncu --clock-control none, one bench step; per-launch averages", "kernels": out}, open(dst, "w"), indent=1)


if __name__ == "__main__":
    main()
