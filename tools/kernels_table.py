#!/usr/bin/env python
"""Print the per-kernel table of bench.py JSON lines (one file per argument)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        j = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(path, 'unreadable:', e)
        continue
    print('%s: n_gpus %d value %.0f Mpix/s, %.3f ms/step, embed %.0f extract %.0f, e2e %.0f' % (
        path, j.get('n_gpus', 1), j['value'], j['ms_per_step'], j.get('embed_mpix_s', 0), j.get('extract_mpix_s', 0),
        j['e2e']['value']))
    for k in j['kernels']:
        print('    %-16s x%-5.1f %8.2f us  share %.3f  %s' % (
            k['name'], k['launches_per_step'], k['avg_us'], k['share'],
            ('%.0f GB/s frac %.3f' % (k['gbs'], k['frac'])) if k['frac'] else ''))
