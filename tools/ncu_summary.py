#!/usr/bin/env python
"""Extracts the judged metrics of an .ncu-rep (CPU box, no GPU needed) into a small text summary for profiles/.

    python tools/ncu_summary.py gpurun_out/p_gatfwd.ncu-rep > profiles/r01_ncu_gatv2_fwd.txt
"""
import csv
import re
import subprocess
import sys
import collections

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("kernel:", vals[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:75s} {vals[i]} {units[i]}")
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(vals[i])
                except ValueError:
                    continue
                if v >= 0.1:
                    print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {v:.3f} warps/issue")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 2:
        hdr, data = rows[1], rows[2:]
        ia, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
        ops = collections.Counter()
        for r in data:
            if len(r) <= max(ia, isrc) or not r[ia].isdigit():          # kernel-name separator rows of multi-kernel reports
                continue
            s = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip())
            ops[s.split()[0].split(".")[0] if s else "?"] += int(r[ia])
        tot = sum(ops.values())
        print("  SASS opcode mix (executed warp instructions):")
        for op, c in ops.most_common(12):
            print(f"    {op:10s} {100 * c / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
