"""Compact per-kernel summary of `ncu --set full` raw-page CSVs: usage
   python tools/ncu_metrics.py out.json raw_a.csv raw_b.csv ...
Writes {kernel_class: {time_us, dram_read_mb, dram_write_mb, traffic_bytes, tensor_pipe_pct, issue_active_pct, ...}} --
bench.py reads `traffic_bytes` (dram__bytes_read.sum + dram__bytes_write.sum of ONE launch) for its roofline object."""
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
WANT = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "inst_executed",
}
CLASS = [("tc_nodeblock_bwd", "tc_nodeblock_bwd"), ("tc_nodeblock_fwd", "tc_nodeblock"), ("tc_bondffn_bwd", "tc_bondffn_bwd"),
         ("tc_bondffn_fwd", "tc_bondffn"), ("tc_edge_d", "tc_edge_d"), ("tc_node_kernel", "tc_node"), ("tc_bwd_node", "tc_node_bwd"),
         ("tc_edge_tail_bwd", "tc_edge_tail_bwd"), ("transition_step", "transition"), ("bwd_node", "bwd_node"),
         ("bwd_edge_tail", "bwd_edge_tail"), ("node_kernel", "node")]


def main(out, paths):
    res = {}
    for path in paths:
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            cls = next((c for pat, c in CLASS if pat in name), re.sub(r"\W+", "_", name)[:40])
            d = {"kernel": name.split("(")[0].replace("<unnamed>::", "")}
            for m, key in WANT.items():
                if m in hdr and r[hdr.index(m)] not in ("", "n/a"):
                    v = float(r[hdr.index(m)].replace(",", ""))
                    d[key] = v * UNIT.get(units[hdr.index(m)], 1.0)
            if "dram_read_bytes" in d and "dram_write_bytes" in d:
                d["traffic_bytes"] = d["dram_read_bytes"] + d["dram_write_bytes"]
            res[cls] = d
    json.dump(res, open(out, "w"), indent=1, sort_keys=True)
    for k, v in res.items():
        print(k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items() if a != "kernel"})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
