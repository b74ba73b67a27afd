#!/usr/bin/env python
"""Summarise an .ncu-rep (from `ncu --set full`) into a small text table: one block per captured kernel with the metrics
the roofline discussion uses.   python scripts/ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg", "sm cycles"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (hmma) active %"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem pipe inst %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    print(f"# {rep}: ncu --set full --clock-control none (values are per launch)")
    for r in rows[2:]:
        print(f"\n== {r[h.index('Kernel Name')][:110]}")
        for key, label in WANT:
            if key in h:
                i = h.index(key)
                print(f"   {label:<34} {r[i]} {units[i]}")


if __name__ == "__main__":
    main()
