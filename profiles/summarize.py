#!/usr/bin/env python
"""Turn an .ncu-rep (brought back in gpurun_out/) into the committed summary:

    python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/r01_x.md ["title"]

Reads the report with `ncu -i ... --page raw --csv` and keeps the metrics the
roofline discussion needs (duration, DRAM bytes, L2 traffic/hit rates, RED
sectors, occupancy, stall reasons, registers, grid)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum",
    "lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum", "lts__t_sectors_srcunit_tex_op_atom.sum",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__d_atomic_input_cycles_active.max.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# %s\n\nsource: `%s` (ncu --set full --clock-control none; per-launch values, cold-ish cache)\n\n" % (title, rep))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("## %s  (launch id %s)\n\n| metric | value | unit |\n|---|---|---|\n" % (d["Kernel Name"][:90], d["ID"]))
            for k in KEYS:
                if k in d and d[k] != "":
                    f.write("| %s | %s | %s |\n" % (k, d[k], units[hdr.index(k)]))
            f.write("\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
