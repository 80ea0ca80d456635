#!/bin/bash
# One gpurun call on one GPU: parity tests, smoke, the driver's default bench line (c2 + c1/c3/c5 sub-objects), c4 on one rank,
# ncu launch list + full capture (summaries only -- the .ncu-rep stays on the box), reference arm.  Usage: bash tools/gpu_full.sh <tag> [ref]
TAG=${1:-f1}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/clocks_before_$TAG.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 4 $OUT/pytest_gpu_$TAG.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -n 2 $OUT/smoke_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; grep -c "bench rank" $OUT/bench_$TAG.err
timeout 600 python bench.py --workload c4 --no-cpu-baseline > $OUT/bench_c4_n1_$TAG.json 2> $OUT/bench_c4_n1_$TAG.err; echo "bench c4 rc=$?"; tail -n 2 $OUT/bench_c4_n1_$TAG.err | cut -c1-300
python tools/kernels_table.py $OUT/bench_$TAG.json $OUT/bench_c4_n1_$TAG.json
python - <<PY
import json
d = json.load(open('$OUT/bench_$TAG.json'))
for k in ('c1', 'c3', 'c5'):
    s = d.get(k) or {}
    print(k, round(s.get('value', 0)), 'Mpix/s', round(s.get('ms_per_step', 0), 4), 'ms; roofline', (s.get('roofline') or {}).get('kernel'), (s.get('roofline') or {}).get('frac'), 'e2e', round((s.get('e2e') or {}).get('value', 0)), s.get('error'))
print('c2 e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --no-extra --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
KEEP_REP=1 bash tools/gpu_ncu.sh $TAG c2 16 8 'row_pipe|col_pipe|topk_' > $OUT/ncu_stdout_$TAG.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_traffic.py $OUT/prof_$TAG.ncu-rep $OUT/dram_traffic_$TAG.json 2>&1 | tr -d "\n" | cut -c1-600; echo
rm -f $OUT/*.ncu-rep
if [ "$2" == "ref" ]; then
  SSW_REF_BUDGET_S=60 timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-400 $OUT/bench_ref_$TAG.json
fi
du -sh $OUT
