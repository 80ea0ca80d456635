#!/usr/bin/env python
"""dram__bytes_read.sum + dram__bytes_write.sum per launch of each kernel in an ncu --set full capture ->
profiles/dram_traffic.json (bench.py's roofline.traffic).  Usage: python tools/ncu_traffic.py rep.ncu-rep out.json"""
import csv
import io
import json
import re
import subprocess
import sys

NAMES = [('RowFwd', 'fwd_rows'), ('RowInv', 'inv_rows'), ('ColPass', None), ('Line1Fwd', 'fwd_line1'), ('Line1Inv', 'inv_line1'),
         ('topk_collect', 'topk_collect'), ('topk_hist', 'topk_hist'), ('topk_block_bin', 'topk_block_bin'),
         ('topk_rank', 'topk_rank'), ('similarity_bank', 'similarity_bank'), ('transpose_kernel', 'transpose')]
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}

raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
acc = {}
for r in rows[2:]:
    k = r[col['Kernel Name']]
    name = None
    m = re.search(r'(ColPipe|RowPipe)<.*>,\s*(?:\(int\))?\d+,\s*(?:(?:\(int\))?\d+,\s*)?(?:\(bool\))?(\d),\s*(?:\(int\))?\d+>\s*>', k)
    if m:   # persistent pipelines: ColPipe<Plan, G, TEAMS, INVERSE, MINB>, RowPipe<Plan, TEAMS, INVERSE, MINB>
        name = ('inv_' if m.group(2) == '1' else 'fwd_') + ('cols' if m.group(1) == 'ColPipe' else 'rows')
    for pat, nm in ([] if name else NAMES):
        if pat in k:
            name = nm
            if pat == 'ColPass':
                # ColPass<Plan<...>, G, TEAMS, INVERSE, MINB>: ncu prints the bool as (bool)1 or as 1 depending on the version
                m = re.search(r'ColPass<.*>,\s*\d+,\s*\d+,\s*(?:\(bool\))?(\d),\s*\d+>\s*>', k)
                name = 'inv_cols' if (m and m.group(1) == '1') else 'fwd_cols'
            break
    if not name:
        continue
    tot = 0.0
    for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        tot += float(r[col[m]]) * UNIT.get(units[col[m]], 1.0)
    acc.setdefault(name, []).append(tot)
out = {k: sum(v) / len(v) for k, v in acc.items()}
json.dump(out, open(sys.argv[2], 'w'), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
