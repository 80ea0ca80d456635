#!/usr/bin/env python
"""SASS evidence for profiles/: per-kernel counts of the Blackwell data-path instructions in libssw.so (TMA tensor /
bulk copies, mbarrier operations, packed FP32, packed conversions) and the SASS around the first TMA load / store of the
column and row pipelines.  Usage: python tools/sass_summary.py [libssw.so] > profiles/rN_sass_pipelines.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'spread_spectrum_watermarking_b200/csrc/libssw.so'
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
WANT = ('UTMALDG', 'UTMASTG', 'UBLKCP', 'UTMACMDFLUSH', 'SYNCS', 'ACQBULK', 'BAR', 'FFMA2', 'FADD2', 'FMUL2', 'F2IP', 'VIMNMX', 'LDGSTS',
        'UTCMMA', 'LDTM', 'HMMA')
funcs = re.split(r'\n\s*Function : ', txt)
tot = collections.Counter()
rows = []
for f in funcs[1:]:
    name = f.split('\n', 1)[0]
    c = collections.Counter(m.group(1) for m in re.finditer(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', f, re.M))
    tot.update(c)
    rows.append((name, c, f))
print('# SASS of %s (cuobjdump -sass), sm_100a' % lib)
print('# UTMALDG / UTMASTG = cp.async.bulk.tensor (TMA tensor-map copies); UBLKCP = cp.async.bulk (linear bulk copies);')
print('# SYNCS = mbarrier operations (init / arrive / expect_tx / try_wait); ACQBULK = griddepcontrol.wait (programmatic dependent launch);')
print('# FFMA2 / FADD2 / FMUL2 = packed f32x2 arithmetic; F2IP = packed / saturating float -> u8 conversion.')
print('library totals:', {k: tot[k] for k in WANT if tot[k]})
print()
print('%-118s %s' % ('kernel', 'counts'))
for name, c, _ in rows:
    if any(t in name for t in ('pipe_kernel', 'lowrank', 'transpose_push', 'topk_', 'similarity_bank')):
        short = re.sub(r'N3ssw4fast', '', name)
        print('%-118s %s  (%d instructions)' % (short[:118], {k: c[k] for k in WANT if c[k]}, sum(c.values())))


def excerpt(pattern, opcode, title, before=6, after=10):
    for name, c, f in rows:
        if re.search(pattern, name) and c[opcode]:
            lines = [ln for ln in f.split('\n') if re.search(r'/\*[0-9a-f]{4}\*/', ln)]
            idx = next(i for i, ln in enumerate(lines) if opcode in ln)
            print()
            print('## %s -- %s' % (title, name[:150]))
            for ln in lines[max(0, idx - before):idx + after]:
                print(re.sub(r'\s+/\* 0x[0-9a-f]+ \*/\s*$', '', ln.rstrip()))
            return


excerpt(r'col_pipe_kernel.*Li2160.*Li4ELi4ELb0', 'UTMALDG', 'forward column pipeline (4K frames): producer issues the TMA tensor loads of a tile')
excerpt(r'col_pipe_kernel.*Li2160.*Li4ELi4ELb0', 'UTMASTG', 'forward column pipeline: TMA tensor store of a finished tile')
excerpt(r'row_pipe_kernel.*Li3840.*Lb0', 'UBLKCP', 'forward row pipeline (4K frames): bulk copy of a row pair into shared memory')
excerpt(r'row_pipe_kernel.*Li3840.*Lb1', 'UBLKCP', 'inverse row pipeline (4K frames): bulk copies in / out')
excerpt(r'row_pipe_kernel.*Li3840.*Lb1', 'F2IP', 'inverse row pipeline: packed FP32 colour arithmetic and saturating conversion to RGB8', 8, 14)
