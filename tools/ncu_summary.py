"""Key metrics of every kernel in an ncu report (--set full) as a small text table for profiles/.

    python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_<workload>_summary.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs(CTA/SM)"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem(CTA/SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%peak"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h, units = rows[0], rows[1]
        print(f"# {rep.split('/')[-1]}  (ncu --set full --clock-control none; cold-cache, serialised replays)")
        for r in rows[2:]:
            print(f"kernel: {r[h.index('Kernel Name')]}   grid {r[h.index('Grid Size')]} block {r[h.index('Block Size')]}")
            for m, label in METRICS:
                if m in h:
                    i = h.index(m)
                    print(f"    {label:24s} {r[i]:>16s} {units[i]}")
        print()


if __name__ == "__main__":
    main()
