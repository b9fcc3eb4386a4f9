#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small CSV: one row per captured launch with the metrics
profiles/README.md quotes.  usage: tools/ncu_summary.py gpurun_out/<tag>/prof.ncu-rep profiles/<name>.csv"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("lts__t_sectors_srcunit_tex.sum", "l2_sectors_from_sm"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
    ("launch__grid_size", "grid"),
    ("launch__cluster_dim_x", "cluster"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall_long_sb_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{n} [{units[cols[m]]}]" if m in cols else n for m, n in WANT])
        for r in rows[2:]:
            w.writerow([r[cols["Kernel Name"]][:60]] + [r[cols[m]] if m in cols else "" for m, _ in WANT])
    print(open(out).read())


if __name__ == "__main__":
    main()
