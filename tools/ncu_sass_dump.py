#!/usr/bin/env python
"""Per-instruction executed counts and stall samples of one kernel launch from an ncu report captured with
--import-source on:   python tools/ncu_sass_dump.py REPORT.ncu-rep KERNEL_REGEX [LAUNCH_SKIP] > out.csv
Columns: address, executed warp instructions, stall samples, SASS.  (Small enough to bring back from the GPU box.)"""
import csv
import io
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + kre, "--launch-skip", skip, "--launch-count", "1", "--csv",
                      "--page", "source"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, x in enumerate(rows) if x and x[0] == "Address"]
h = rows[hi[0]]
end = hi[1] - 1 if len(hi) > 1 else len(rows)
si, ii, src = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
w = csv.writer(sys.stdout)
w.writerow(["address", "executed", "samples", "sass"])
for x in rows[hi[0] + 1:end]:
    if len(x) == len(h):
        w.writerow([x[0], x[ii], x[si], x[src].strip()])
