#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` raw exports: DRAM bytes (read + write) per pipeline stage of one frame.

  python tools/traffic_from_ncu.py cfg3=gpurun_out/r2prof/full_cfg3_raw.csv[:SKIP:COUNT] cfg5=... > profiles/traffic.json

(:SKIP:COUNT selects the launches of ONE frame when the capture window is not aligned with a frame.)

Stage = what bench.py's stage timers bracket; a stage that runs as two launches (row bands) is summed and the
launch count is recorded next to it (the bench line's roofline.traffic is per stage, i.e. per frame)."""
import csv
import json
import re
import sys

STAGE = [("k_prefilter", "prefilter"), ("k_cost", "cost"), ("k_horizontal", "horizontal"), ("k_sweep", None), ("k_vertical", None),
         ("k_pad_cost", "cost"), ("k_fill", "init"), ("k_init_wta", "init"), ("k_lrcheck", "lrcheck"), ("k_lr_median", "median"), ("k_median", "median"), ("k_cc_", "speckle"), ("k_reproject", "tail"),
         ("k_compact", "tail"), ("k_disp_to_float", "tail")]


def main():
    out = {}
    for arg in sys.argv[1:]:
        cfg, path = arg.split("=")
        skip, count = 0, None
        if path.count(":") == 2:
            path, a, b = path.split(":")
            skip, count = int(a), int(b)
        rows = list(csv.reader(open(path)))
        h, u = rows[0], rows[1]
        rows = rows[:2] + rows[2 + skip: (2 + skip + count) if count else None]
        ik, ir, iw, it = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        st, ms, n = {}, {}, {}
        nsweep = 0
        for r in rows[2:]:
            name = r[ik]
            stage = None
            for pre, s in STAGE:
                if pre in name:
                    stage = s
                    if pre in ("k_sweep", "k_vertical"):
                        nsweep += 1
                        # MODE_HH / HH4 run two sweeps (spill, then winner-take-all); the other modes one
                        # k_sweep<NREG, LPC, SAT, WROLE, RPS>: the spilling sweep of MODE_HH is the kernel without the W role
                        m = re.search(r"k_sweep<\s*\d+,\s*\d+,\s*\d+,\s*(\d+)", name)
                        stage = "vertical_fwd" if (m and m.group(1) == "0") else "vertical_wta"
                    break
            if stage is None:
                continue
            b = float(r[ir]) * scale[u[ir]] + float(r[iw]) * scale[u[iw]]
            st[stage] = st.get(stage, 0.0) + b
            ms[stage] = ms.get(stage, 0.0) + float(r[it])
            n[stage] = n.get(stage, 0) + 1
        out[cfg] = {k: int(v) for k, v in st.items()}
        out[cfg]["_launches"] = n
        out[cfg]["_ncu_ms"] = {k: round(v, 4) for k, v in ms.items()}
        out[cfg]["_frame_total"] = int(sum(st.values()))
    out["_source"] = "dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of each stage of ONE frame, ncu --set full " \
                     "--clock-control none (round 2: profiles/r02_ncu_full_cfg3_summary.csv, r02_ncu_full_cfg5_summary.csv)"
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
