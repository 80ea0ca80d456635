#!/usr/bin/env python
"""Aggregate the ncu source page (SASS) of one kernel launch: stall samples per barrier-delimited
phase and per opcode.  Usage: python tools/ncu_source.py rep.ncu-rep <launch-skip> [top]"""
import csv
import io
import subprocess
import sys
import collections

rep, skip = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--launch-skip', skip,
                      '--launch-count', '1'], capture_output=True, text=True).stdout
lines = raw.splitlines()
print(lines[0][:200])
rows = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
S = col['# Samples']; I = col['Instructions Executed']
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
phase = 0
per_phase = collections.OrderedDict()
per_op = collections.Counter(); per_op_inst = collections.Counter()
tot = 0
seen = set()
for r in rows[1:]:
    if len(r) < len(hdr) or r[S] == '# Samples' or r[0] in seen:
        continue
    seen.add(r[0])
    src = r[col['Source']].strip()
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    op = op.split('.')[0]
    n = int(r[S] or 0); ie = int(r[I] or 0)
    d = per_phase.setdefault(phase, {'samples': 0, 'inst': 0, 'st': collections.Counter()})
    d['samples'] += n; d['inst'] += ie
    for h in stalls:
        v = int(r[col[h]] or 0)
        if v:
            d['st'][h] += v
    per_op[op] += n; per_op_inst[op] += ie
    tot += n
    if op == 'BAR':
        phase += 1
print('total samples', tot)
for ph, d in per_phase.items():
    top = ', '.join('%s %d' % (k.replace('stall_', ''), v) for k, v in d['st'].most_common(5))
    print('phase %2d: samples %6d (%.1f%%)  warp-inst %9d   %s' % (ph, d['samples'], 100.0 * d['samples'] / max(tot, 1), d['inst'], top))
print('by opcode (samples, warp-inst):')
for op, n in per_op.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 14):
    print('   %-8s %7d  %10d' % (op, n, per_op_inst[op]))
