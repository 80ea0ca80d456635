#!/bin/bash
# step-time distribution of bench c2 per SSW_PDL_MODE (several fresh processes each): bash tools/ab_modes.sh <tag> <repeats> <mode>...
OUT=gpurun_out; mkdir -p $OUT
TAG=$1; REP=$2; shift; shift
for r in $(seq 1 $REP); do
  for m in "$@"; do
    SSW_PDL_MODE=$m timeout 200 python bench.py --no-cpu-baseline --no-e2e > $OUT/ab_${TAG}_c2_m${m}_$r.json 2>/dev/null
  done
done
python - <<PY
import json, glob
for m in "$*".split():
    v = []
    for f in sorted(glob.glob('$OUT/ab_${TAG}_c2_m%s_*.json' % m)):
        d = json.load(open(f)); v.append((round(d['ms_per_step'] * 1e3, 1), round(d['embed_mpix_s'] / 1e3, 1), round(d['extract_mpix_s'] / 1e3, 1)))
    print('mode', m, '(us/step, embed Gpix/s, extract Gpix/s):', v)
PY
