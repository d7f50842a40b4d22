#!/usr/bin/env python
"""Table of the sharded config-5 sweep: one line per (size, kernel), one column per GPU count.

    python benchmarks/summarize_sweep_sharded.py profiles/r2_multi_gpu/sweep_config5_n{1,2,4,8}.json"""
import json
import sys


def main():
    runs = {}
    for p in sys.argv[1:]:
        d = json.load(open(p))
        runs[d["n_gpus"]] = d["rows"]
    ns = sorted(runs)
    keys = []
    for r in runs[ns[0]]:
        k = (r.get("Hq"), r.get("D"), r.get("H"), r["kernel"])
        if k not in keys:
            keys.append(k)
    print("== config 5 sweep, job of 8 pairs sharded over N GPUs (benchmarks/sweep_sharded.py): job GB/s and per-GPU fraction of "
          "the measured HBM peak")
    for k in keys:
        hq, dd, h, name = k
        head = f"Hq={hq:3d} D={dd:3d} {name:<26s}" if hq is not None else f"H={h:4d} {name:<33s}"
        cols = []
        for n in ns:
            r = next((r for r in runs[n] if (r.get("Hq"), r.get("D"), r.get("H"), r["kernel"]) == k), None)
            if r is None:
                cols.append(f"N={n}        --")
            elif hq is not None:
                cols.append(f"N={n} {r['ms']:7.3f} ms {r['job_GBps']:8.0f} GB/s {r['per_gpu_frac_of_measured_peak']:.2f}")
            else:
                cols.append(f"N={n} {r['ms']:7.3f} ms")
        print(head + " | " + " | ".join(cols))


if __name__ == "__main__":
    main()
