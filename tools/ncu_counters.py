"""Merges the counters of an .ncu-rep into profiles/r02_ncu_counters.json (the file bench.py quotes in roofline.counters).
usage: python tools/ncu_counters.py report.ncu-rep WORKLOAD SOURCE_NOTE [name,name,...]   (only these timing-record names)
Kernels are matched by name to the names of the library's timing records (kernels[].name of a bench line)."""
import csv
import json
import os
import subprocess
import sys

NAMES = {"k_ssao": "ssao", "k_lighting_hard": "lighting", "k_pcss_visibility": "pcss_visibility", "k_chain_fused": "pcss_chain",
         "k_resolve_geometry": "resolve_geometry", "k_resolve_shadow": "resolve_shadow", "k_pixel_masks": "pcss_masks",
         "k_shadow_coords": "shadow_coords", "k_classify": "pcss_classify", "k_chunk_index": "pcss_chunk_index", "k_blur_h": "blur_h",
         "k_blur_v": "blur_v", "k_setup": "setup", "k_raster_blocks": "raster_blocks", "k_raster_small": "raster_small"}
COLS = {"gpu__time_duration.sum": ("dur_us", {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}),
        "dram__bytes_read.sum": ("_rd", {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}),
        "dram__bytes_write.sum": ("_wr", {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}),
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": ("dram_pct", None),
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": ("sm_pct", None),
        "sm__warps_active.avg.pct_of_peak_sustained_active": ("occupancy_pct", None),
        "launch__registers_per_thread": ("regs", None),
        "lts__t_sector_hit_rate.pct": ("l2_hit_pct", None),
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": ("fp64_pipe_pct", None),
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": ("stall_long_scoreboard", None),
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": ("stall_barrier", None),
        "smsp__inst_executed.sum": ("warp_instructions", None)}


def main():
    rep, workload, note = sys.argv[1], sys.argv[2], sys.argv[3]
    only = set(sys.argv[4].split(",")) if len(sys.argv) > 4 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    acc = {}
    for r in rows[2:]:
        kname = r[hdr.index("Kernel Name")]
        name = next((v for k, v in NAMES.items() if k in kname), None)
        if not name or (only and name not in only):
            continue
        d = {}
        for key, (field, scale) in COLS.items():
            cols = [i for i, h in enumerate(hdr) if h == key]
            if not cols:
                continue
            try:
                v = float(r[cols[0]].replace(",", ""))
            except ValueError:
                continue
            if scale:
                v *= scale.get(units[cols[0]], 1)
            d[field] = v
        acc.setdefault(name, []).append(d)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_ncu_counters.json")
    data = json.load(open(path))
    for name, ds in acc.items():
        e = {"launches_profiled": len(ds), "source": note}
        for f in set().union(*ds):
            vals = [d[f] for d in ds if f in d]
            e[f] = round(sum(vals) / len(vals), 2)
        e["dram_bytes_per_launch"] = int(e.pop("_rd", 0) + e.pop("_wr", 0))
        data.setdefault(workload, {})[name] = e
        print(name, e)
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
