#!/usr/bin/env python
"""Condense the per-launch blocks that scripts/ncu_summary.py writes into one line per launch.

    python scripts/ncu_table.py matcha=profiles/r02_ncu_full_matcha_step_v9.txt vocoder=profiles/r02_ncu_full_vocoder_v9.txt > profiles/r02_ncu_full_table_v9.txt
"""
import re
import sys


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def to_mb(val, unit):
    return num(val) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, float("nan"))


def to_us(val, unit):
    return num(val) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, float("nan"))


def blocks(path):
    cur = None
    for line in open(path):
        if line.startswith("== "):
            if cur:
                yield cur
            m = re.search(r"(\w+)(<[^>(]*>)?\(", line[3:])
            name = (m.group(1) + (m.group(2) or "")) if m else line[3:].strip()
            cur = {"name": name.replace("(int)", "")}
        elif cur is not None and line.startswith("   "):
            label, rest = line[3:37].strip(), line[37:].split()
            cur[label] = (rest[0], rest[1] if len(rest) > 1 else "") if rest else ("nan", "")
    if cur:
        yield cur


print("# ncu --set full --clock-control none, one launch per row (cold caches, serialised); condensed by scripts/ncu_table.py")
print(f"{'kernel':<40}{'us':>8}{'grid':>7}{'tensor pipe active %':>22}{'DRAM rd MB':>12}{'DRAM wr MB':>12}{'L2->SM MB':>11}{'smem bank conflicts':>21}")
for arg in sys.argv[1:]:
    tag, path = arg.split("=", 1)
    print("## " + tag)
    for b in blocks(path):
        g = lambda k: b.get(k, ("nan", ""))
        print(f"{b['name'][:39]:<40}{to_us(*g('duration')):8.1f}{num(g('grid')[0]):7.0f}{num(g('tensor pipe (hmma) active %')[0]):22.1f}"
              f"{to_mb(*g('dram read')):12.1f}{to_mb(*g('dram write')):12.1f}{to_mb(*g('L2->SM bytes')):11.1f}{num(g('smem bank conflicts')[0]):21.0f}")
