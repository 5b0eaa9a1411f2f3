#!/usr/bin/env python
"""Instruction-level summary of one kernel from an ncu report captured with --import-source on.

  python tools/ncu_source_summary.py REPORT.ncu-rep KERNEL_REGEX [LAUNCH_SKIP] > profiles/NAME.md

Reads `ncu --page source --csv` (SASS view: opcode mix, stall samples) and `--print-source cuda,sass`
(per CUDA line totals).  Spin iterations of bounded mbarrier waits (everything after the first failed
try_wait of a wait site up to the back-branch behind the clock read) are reported separately: ncu's
instrumented pass makes waiting warps spin far longer than a normal run does.
"""
import collections
import csv
import io
import re
import subprocess
import sys


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    sel = ["-i", rep, "--kernel-name", "regex:" + kre, "--launch-skip", skip, "--launch-count", "1", "--csv"]
    raw = list(csv.reader(io.StringIO(ncu(sel + ["--page", "raw"]))))
    h = raw[0]
    r = raw[2]
    get = lambda k: r[h.index(k)] if k in h else "n/a"
    print("# %s\n" % r[h.index("Kernel Name")])
    print("Report `%s`, launch skip %s.  Hardware counters of the launch: %s ms, %s warp instructions, issue-active %s %%, "
          "DRAM %s %% of peak, registers/thread %s, dynamic shared memory %s KB.\n"
          % (rep.split("/")[-1], skip, get("gpu__time_duration.sum"), get("smsp__inst_executed.sum"),
             get("smsp__issue_active.avg.pct_of_peak_sustained_active"), get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
             get("launch__registers_per_thread"), get("launch__shared_mem_per_block_dynamic")))
    rows = list(csv.reader(io.StringIO(ncu(sel + ["--page", "source"]))))
    hi = [i for i, x in enumerate(rows) if x and x[0] == "Address"]
    h = rows[hi[0]]
    end = hi[1] - 1 if len(hi) > 1 else len(rows)
    d = [x for x in rows[hi[0] + 1:end] if len(x) == len(h)]
    si, ii, src = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    spin = [False] * len(d)
    for i, x in enumerate(d):
        if "TRYWAIT" in x[src]:
            seen, e = False, None
            for j in range(i, min(i + 24, len(d))):
                if "CS2R" in d[j][src]:
                    seen = True
                if seen and "BRA" in d[j][src]:
                    e = j
                    break
            if e is not None:
                for j in range(i + 2, e + 1):
                    spin[j] = True
    tot_i = sum(num(x[ii]) for x in d)
    spin_i = sum(num(x[ii]) for x, m in zip(d, spin) if m)
    ti = tot_i - spin_i
    ts = sum(num(x[si]) for x in d)
    print("SASS: %d instructions.  Instrumented pass: %d warp instructions, %.1f %% of them spin iterations of bounded waits "
          "(excluded below: %d).\n" % (len(d), tot_i, 100.0 * spin_i / max(tot_i, 1), ti))
    ops = collections.Counter()
    smp = collections.Counter()
    for x, m in zip(d, spin):
        mm = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", x[src])
        op = mm.group(2) if mm else "?"
        smp[op] += num(x[si])
        if not m:
            ops[op] += num(x[ii])
    print("## Opcode mix (share of executed warp instructions without spin; share of stall samples)\n")
    print("| opcode | instructions | samples |\n|---|---|---|")
    for op, c in ops.most_common(24):
        print("| %s | %.1f %% | %.1f %% |" % (op, 100.0 * c / max(ti, 1), 100.0 * smp[op] / max(ts, 1)))
    print("\n## Stall reasons (share of %d samples)\n" % ts)
    st = [(sum(num(x[h.index(c)]) for x in d), c) for c in h if c.startswith("stall_") and "Not Issued" not in c]
    print(", ".join("%s %.1f %%" % (c.replace("stall_", ""), 100.0 * v / max(ts, 1)) for v, c in sorted(st, reverse=True) if v > 0.01 * ts))
    rows = list(csv.reader(io.StringIO(ncu(sel + ["--page", "source", "--print-source", "cuda,sass"]))))
    cur, agg, done = None, [], set()
    for x in rows:
        if not x:
            continue
        if x[0] == "File Path":
            cur = x[1].split("/")[-1]
            continue
        if len(x) > 8 and x[2] == "-":
            key = (cur, x[0])
            if key in done:
                continue
            done.add(key)
            try:
                agg.append((cur, int(x[0]), x[1].strip(), num(x[6]), num(x[7])))
            except ValueError:
                pass
    t2, s2 = sum(a[4] for a in agg), sum(a[3] for a in agg)
    print("\n## CUDA lines with the most executed instructions (instrumented pass, spin included)\n")
    print("| file:line | instructions | samples | source |\n|---|---|---|---|")
    for a in sorted(agg, key=lambda a: -a[4])[:28]:
        print("| %s:%d | %.2f %% | %.2f %% | `%s` |" % (a[0], a[1], 100.0 * a[4] / max(t2, 1), 100.0 * a[3] / max(s2, 1),
                                                     a[2][:88].replace("|", "\\|")))


if __name__ == "__main__":
    main()
