#!/usr/bin/env python
"""Condense `ncu --page raw --csv` exports into the per-kernel summary committed under profiles/.

  ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
  python tools/ncu_summary.py out_summary.csv raw1.csv [raw2.csv ...]
"""
import csv
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    out = sys.argv[1]
    rows_out = []
    for path in sys.argv[2:]:
        rows = list(csv.reader(open(path)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        stall = [i for i, h in enumerate(hdr) if "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")]
        for r in data:
            name = r[hdr.index("Kernel Name")]
            if not r[hdr.index("gpu__time_duration.sum")] or "nan" in r[hdr.index("gpu__time_duration.sum")]:
                continue
            rec = {"kernel": name}
            for k in KEEP:
                if k in hdr:
                    rec[k + " [" + units[hdr.index(k)] + "]"] = r[hdr.index(k)]
            top = sorted(((float(r[i] or 0), hdr[i]) for i in stall), reverse=True)[:5]
            rec["top_stalls (warps per issue)"] = "; ".join(
                "%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, h in top)
            rows_out.append(rec)
    keys = []
    for r in rows_out:
        for k in r:
            if k not in keys:
                keys.append(k)
    w = csv.DictWriter(open(out, "w", newline=""), fieldnames=keys)
    w.writeheader()
    for r in rows_out:
        w.writerow(r)
    print("wrote", out, len(rows_out), "kernels")


if __name__ == "__main__":
    main()
