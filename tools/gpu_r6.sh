#!/bin/bash
# r6: parity tests, then the driver's default invocation (c2 line + c1/c3/c5 sub-objects) with its wall time, and a low-rank A/B
TAG=${1:-r6}
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 12 $OUT/pytest_gpu_$TAG.log
for v in 0 1; do
  SSW_LOWRANK=$v timeout 300 python bench.py --workload c2 --steps 100 --no-cpu-baseline --no-e2e > $OUT/bench_c2_${TAG}_l$v.json 2> $OUT/bench_c2_${TAG}_l$v.err; echo "c2 lowrank=$v rc=$?"; tail -n 2 $OUT/bench_c2_${TAG}_l$v.err
  SSW_LOWRANK=$v timeout 300 python bench.py --workload c3 --steps 10 --no-cpu-baseline --no-e2e > $OUT/bench_c3_${TAG}_l$v.json 2> $OUT/bench_c3_${TAG}_l$v.err; echo "c3 lowrank=$v rc=$?"; tail -n 2 $OUT/bench_c3_${TAG}_l$v.err
done
python tools/kernels_table.py $OUT/bench_c2_${TAG}_l*.json $OUT/bench_c3_${TAG}_l*.json 2>&1
SECONDS=0
timeout 900 python bench.py > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; echo "default bench rc=$? wall ${SECONDS}s"; tail -n 3 $OUT/bench_default_$TAG.err
python - $TAG <<'PY'
import json, sys
j = json.loads(open('gpurun_out/bench_default_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
def line(name, d):
    print('%-3s %9.0f Mpix/s  %.4f ms/step  e2e %8.0f Mpix/s (%.3f ms/step)  roofline %s %.3f  min_sim %.2f' % (
        name, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d.get('min_similarity', 0)))
line('c2', j)
for k in ('c1', 'c3', 'c5', 'c4'):
    if k in j and j[k]: line(k, j[k])
print('cpu_baseline', j.get('cpu_baseline'))
print('clocks', j.get('clocks'))
PY
SSW_REF_BUDGET_S=40 timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-400 $OUT/bench_ref_$TAG.json
