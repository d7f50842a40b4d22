#!/usr/bin/env python
"""Condense an `ncu --set full` report into the JSON summaries kept under profiles/.

    ncu -i report.ncu-rep --page raw --csv | python benchmarks/ncu_summary.py > profiles/rN_ncu_full_....json

One entry per profiled launch (the last launch of each kernel name is kept): duration, DRAM bytes, issue / pipe
utilisation, shared-memory wavefronts against the ideal, occupancy, launch geometry and the top stall reasons."""
import csv
import json
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__waves_per_multiprocessor",
]


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    out = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        e = {"kernel": name[:110]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                e[k] = (r[i] + " " + units[i]).strip()
        st = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(r[i])
              for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and r[i]}
        e["stalls_per_issue"] = {k: round(v, 2) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:5]}
        out[name[:110]] = e
    json.dump(list(out.values()), sys.stdout, indent=1)


if __name__ == "__main__":
    main()
