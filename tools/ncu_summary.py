"""Summarises an .ncu-rep (raw page) into one line per profiled launch: duration, DRAM bytes, occupancy, top stalls.
usage: python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...] > profiles/summary.md"""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "dur_us", 1e-3), ("dram__bytes_read.sum", "dram_rd_MB", None), ("dram__bytes_write.sum", "dram_wr_MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct", 1), ("launch__registers_per_thread", "regs", 1),
        ("lts__t_sector_hit_rate.pct", "l2_hit", 1), ("l1tex__t_sector_hit_rate.pct", "l1_hit", 1),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pct", 1),
        ("sm__inst_executed_pipe_fp64.sum", "fp64_inst", 1),
        ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "st_long_sb", 1),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb", 1),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier", 1),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb", 1),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg", 1),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait", 1),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math", 1),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar", 1)]


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            print("(no rows in %s)" % rep)
            continue
        hdr, units = rows[0], rows[1]
        print("## %s" % rep)
        for r in rows[2:]:
            d = {"kernel": r[hdr.index("Kernel Name")], "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
            for key, name, scale in KEYS:
                cols = [i for i, h in enumerate(hdr) if h == key or h.endswith("." + key)]
                if not cols or name in d:
                    continue
                v = to_float(r[cols[0]])
                if v is None:
                    continue
                u = units[cols[0]]
                if name.endswith("_MB"):
                    v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1e-6)
                elif name == "dur_us":
                    v = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
                d[name] = round(v, 3)
            print("- " + ", ".join("%s=%s" % kv for kv in d.items()))


if __name__ == "__main__":
    main()
