"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion needs.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xxx.txt]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active", "sm__pipe_fmaheavy_cycles_active",
        "sm__inst_executed_pipe_fma", "sm__pipe_tensor", "smsp__inst_executed.avg.per_cycle_active", "sm__warps_active.avg.pct",
        "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate", "smsp__issue_active.avg.pct", "smsp__average_warp",
        "smsp__warp_issue_stalled", "sm__cycles_active.avg", "smsp__inst_executed.sum", "l1tex__t_bytes", "dram__cycles_active",
        "smsp__pcsamp_warps_issue_stalled"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"=== {r[hdr.index('Kernel Name')][:60]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for i, h in enumerate(hdr):
            if any(w in h for w in WANT) and r[i] not in ("", "0", "n/a"):
                print(f"  {h} [{units[i]}] = {r[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
