#!/usr/bin/env python
"""Per-role instruction counts of the two cfg3 sweeps from tools/ncu_sass_dump.py exports (spin iterations of the bounded
waits removed, regions split at the roles' EXITs):   python tools/sweep_roles_table.py DIR > profiles/...md
DIR holds sass_wta.csv and sass_fwd.csv."""
import collections
import csv
import re
import sys

D = sys.argv[1]
# optional: workload tag of the second table set (cfg2: 144 strips x 720 rows, k_sweep<8,8,0,1,2> with two rows per ring stage)
STRIPS, ROWS = 148, 2160


def analyse(nm, roles, STRIPS=STRIPS, ROWS=ROWS):
    rows = list(csv.DictReader(open('%s/sass_%s.csv' % (D, nm))))
    ex = [int(float(r['executed'] or 0)) for r in rows]
    sass = [r['sass'] for r in rows]
    spin = [False] * len(rows)
    for i, s in enumerate(sass):
        if 'TRYWAIT' in s:
            seen, e = False, None
            for j in range(i, min(i + 24, len(rows))):
                if 'CS2R' in sass[j]:
                    seen = True
                if seen and 'BRA' in sass[j]:
                    e = j
                    break
            if e is not None:
                for j in range(i + 2, e + 1):
                    spin[j] = True
    idx = [i for i, s in enumerate(sass) if 'EXIT' in s]
    prev, regions = 0, []
    for i in idx + [len(rows) - 1]:
        tot = sum(ex[prev:i + 1])
        ns = sum(e for e, m in zip(ex[prev:i + 1], spin[prev:i + 1]) if not m)
        if tot > 0.005 * sum(ex):
            regions.append((prev, i, tot, ns))
        prev = i + 1
    allns = sum(e for e, m in zip(ex, spin) if not m)
    packed = sum(ex[k] for k in range(len(rows)) if not spin[k] and
                 (re.search(r'VIMNMX3?\.U16x2|VIADDMNMX\.U16x2|VIMNMX3?\.U32', sass[k]) or re.match(r'\s*(@!?U?P\d+\s+)?PRMT', sass[k])))
    out = []
    for (a, b, tot, ns), (role, nw) in zip(regions, roles):
        ops = collections.Counter()
        for k in range(a, b + 1):
            if spin[k]:
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", sass[k])
            ops[m.group(2) if m else '?'] += ex[k]
        top = ', '.join('%s %.0f' % (o, c / (STRIPS * ROWS * nw)) for o, c in ops.most_common(12))
        out.append('| %s | %d | %.1f %% | %.0f | %s |' % (role, nw, 100 * ns / allns, ns / (STRIPS * ROWS * nw), top))
    return allns, packed, out


txt = ["# Per-role instruction counts of the cfg3 sweeps (round 2, final build)", "",
       "From `ncu --set full --import-source on` (per-instruction executed counts, `tools/ncu_sass_dump.py`), spin iterations of the "
       "bounded waits removed, regions split at the roles' `EXIT`s; counts are warp instructions per warp and image row "
       "(148 strips x 2160 rows).", "",
       "Round 1 for comparison (`r01_sweep_wta_instruction_mix_v9.md`): 2.80 G warp instructions in the WTA sweep (8760 per SM "
       "and row), 2.05 G in the forward sweep; role V ~300 per row.", ""]
import os
for nm, roles, title, geo in (('wta', [('V', 7), ('A', 8), ('C', 8), ('W', 7), ('producer', 1)], 'k_sweep<16,8,0,1,1> (cfg3 backward sweep + WTA)', (148, 2160)),
                              ('fwd', [('V', 7), ('A', 8), ('C', 8), ('producer', 1)], 'k_sweep<16,8,0,0,1> (cfg3 forward sweep, spills S)', (148, 2160)),
                              ('cfg2_wta', [('V', 2), ('A', 4), ('C', 4), ('W', 6), ('producer', 1)],
                               'k_sweep<8,8,0,1,2> (cfg2: 1280x720 D=128 MODE_SGBM, TWO rows per ring stage; 144 strips of 8 columns)', (144, 720))):
    if not os.path.exists('%s/sass_%s.csv' % (D, nm)):
        continue
    allns, packed, out = analyse(nm, roles, *geo)
    txt += ["## %s" % title, "",
            "non-spin instructions (instrumented pass): %.3f G = %.0f per SM and row; packed min / max / add-min / permute share %.1f %%"
            % (allns / 1e9, allns / (geo[0] * geo[1]), 100 * packed / allns), "",
            "| role | warps | share | instr / warp-row | top opcodes (per warp-row) |", "|---|---|---|---|---|"] + out + [""]
print('\n'.join(txt))
