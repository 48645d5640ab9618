"""profiles/rNN_ncu_traffic.json from ncu reports (--set full): per workload and kernel group the counters SURVEY.md 8(d)
names -- DRAM bytes per launch, DRAM / SM throughput as % of peak, achieved occupancy, warp-execution efficiency,
instruction counts -- for bench.py's `roofline.traffic` and for DESIGN.md.

    python tools/ncu_traffic.py profiles/r02_ncu_traffic.json c4=gpurun_out/x_prof_c4.ncu-rep c3=... c5=...

A report may hold several launches of a kernel: the LAST one is taken (steady state).  Kernel groups are keyed the way
bench.py names them: k_setup, k_vis, k_shade, k_fill, plus the binning kernels for the record.
"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",      # (ncu 2025: dram__throughput... lives under this name)
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
    "smsp__inst_executed.sum": "warp_inst",
    "launch__registers_per_thread": "registers",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
GROUPS = ("k_setup", "k_vis", "k_shade", "k_fill", "k_raster", "k_bin_fill", "k_bin_scan", "k_chunk_cull", "k_upload")


def parse(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    res = {}
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        g = next((k for k in GROUPS if name.split("<")[0].split("(")[0].strip().endswith(k) or f" {k}" in f" {name}"), None)
        if g is None:
            continue
        e = {"kernel": name.split("(")[0], "grid": r[h.index("Grid Size")], "block": r[h.index("Block Size")]}
        for m, key in WANT.items():
            if m in h and key not in e:
                i = h.index(m)
                v = float(r[i].replace(",", ""))
                e[key] = v * UNIT_SCALE.get(units[i], 1.0) if key in ("time_us", "dram_read_bytes", "dram_write_bytes") else v
        res[g] = e          # later launches overwrite earlier ones
    return res


def main():
    dst = sys.argv[1]
    out = {}
    for arg in sys.argv[2:]:
        wl, rep = arg.split("=", 1)
        d = parse(rep)
        for e in d.values():
            e["capture"] = rep.split("/")[-1]
        out[wl] = d
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    for wl, d in out.items():
        for g, e in d.items():
            print(f"{wl:4s} {g:12s} {e.get('time_us', 0):9.1f} us  dram {e.get('dram_read_bytes', 0) / 1e6:7.1f} + {e.get('dram_write_bytes', 0) / 1e6:7.1f} MB"
                  f"  dram% {e.get('dram_pct_of_peak', 0):5.1f}  sm% {e.get('sm_pct_of_peak', 0):5.1f}  occ% {e.get('achieved_occupancy_pct', 0):5.1f}"
                  f"  issue% {e.get('issue_active_pct', 0):5.1f}  thr/inst {e.get('threads_per_inst', 0):5.2f}  winst {e.get('warp_inst', 0) / 1e6:8.1f} M  regs {e.get('registers', 0):.0f}")


if __name__ == "__main__":
    main()
