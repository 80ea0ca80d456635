#!/bin/bash
# N-GPU box: host copy ceiling, multi-rank parity of the sharded path, the driver's default bench at N
TAG=${1:-n8}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/h2d_ceiling.py > $OUT/h2d_ceiling_n${N}_$TAG.json 2> $OUT/h2d_ceiling_n${N}_$TAG.err; echo "ceiling N=$N rc=$?"; cat $OUT/h2d_ceiling_n${N}_$TAG.json | cut -c1-700
timeout 120 python tools/h2d_ceiling.py > $OUT/h2d_ceiling_n1_$TAG.json 2> $OUT/h2d_ceiling_n1_$TAG.err; echo "ceiling N=1 rc=$?"; cat $OUT/h2d_ceiling_n1_$TAG.json | cut -c1-500
timeout 1200 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "multi_rank" > $OUT/pytest_multi_$TAG.log 2>&1; echo "pytest multi rc=$?"
tail -n 6 $OUT/pytest_multi_$TAG.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
    > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err; echo "bench N=$N rc=$?"; grep -v "^\s*$" $OUT/bench_n${N}_$TAG.err | grep -n "FAILED\|Fatal\|Error\|error" | head -20 | cut -c1-250
python - <<PY
import json
try:
    d = json.load(open('$OUT/bench_n${N}_$TAG.json'))
    print('c2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    for k in ('c1', 'c3', 'c5', 'c4'):
        if k in d: print(k, d[k].get('value'), d[k].get('ms_per_step'), d[k].get('error'), {a: b for a, b in d[k].items() if a in ('kernel_ms_per_step', 'exposed_comm_ms_per_step', 'scaling')}, 'e2e', (d[k].get('e2e') or {}).get('value'))
    if 'c4' in d:
        for e in d['c4'].get('kernels', [])[:8]: print('   %-18s x%.1f %9.2f us share %.3f frac %s' % (e['name'], e['launches_per_step'], e['avg_us'], e['share'], e['frac']))
except Exception as e:
    print('no line', e)
PY
