#!/bin/bash
# End-to-end check on one box: the copy ceiling of the box, then the c2 line with the full e2e leg, partial inverse on / off.
TAG=${1:-e2e}
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python tools/h2d_ceiling.py > $OUT/h2d_ceiling_n1_$TAG.json 2> $OUT/h2d_ceiling_n1_$TAG.err; echo "ceiling rc=$?"; cat $OUT/h2d_ceiling_n1_$TAG.json | cut -c1-600
for pi in 1 0; do
  SSW_PARTIAL_INV=$pi timeout 300 python bench.py --no-extra --no-cpu-baseline > $OUT/bench_${TAG}_pi$pi.json 2> $OUT/bench_${TAG}_pi$pi.err; echo "bench pi$pi rc=$?"
  python - <<PY
import json
d = json.load(open('$OUT/bench_${TAG}_pi$pi.json'))
print('pi$pi: value', round(d['value']), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), 'ms', round(d['e2e']['ms_per_step'], 4), 'steps', d['e2e'].get('steps'))
PY
done
