#!/bin/bash
# r5: F2I pack A/B (prebuilt libs), column-histogram A/B, parity tests, c2 bench with the asynchronous e2e path
TAG=${1:-r5}
OUT=gpurun_out; mkdir -p $OUT
bash tools/ab_libs.sh $TAG v5a v5 2>&1 | grep -E "json|inv_rows|fwd_rows"
for v in 0 1; do
  SSW_COL_HIST=$v timeout 300 python bench.py --workload c2 --steps 100 --no-cpu-baseline --no-e2e > $OUT/bench_c2_${TAG}_h$v.json 2> $OUT/bench_c2_${TAG}_h$v.err; echo "c2 hist=$v rc=$?"; tail -n 2 $OUT/bench_c2_${TAG}_h$v.err
  SSW_COL_HIST=$v timeout 300 python bench.py --workload c3 --steps 10 --no-cpu-baseline --no-e2e > $OUT/bench_c3_${TAG}_h$v.json 2> $OUT/bench_c3_${TAG}_h$v.err; echo "c3 hist=$v rc=$?"; tail -n 2 $OUT/bench_c3_${TAG}_h$v.err
done
python tools/kernels_table.py $OUT/bench_c2_${TAG}_h*.json $OUT/bench_c3_${TAG}_h*.json 2>&1 | grep -E "json|fwd_cols|topk"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 15 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py --workload c2 --no-cpu-baseline > $OUT/bench_c2_${TAG}_e2e.json 2> $OUT/bench_c2_${TAG}_e2e.err; echo "c2 e2e rc=$?"; tail -n 3 $OUT/bench_c2_${TAG}_e2e.err
python - $TAG <<'PY'
import json, sys
j = json.loads(open('gpurun_out/bench_c2_%s_e2e.json' % sys.argv[1]).read().strip().splitlines()[-1])
print('c2: %.3f ms/step (regions %s), e2e %.0f Mpix/s = %.3f ms/step' % (j['ms_per_step'], j['regions_ms_per_step'], j['e2e']['value'], j['e2e']['ms_per_step']))
PY
