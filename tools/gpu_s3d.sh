#!/bin/bash
TAG=${1:-s3d}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/pipe_trace.py > $OUT/pipe_trace_$TAG.txt 2> $OUT/pipe_trace_$TAG.err; echo "trace rc=$?"; tail -n 3 $OUT/pipe_trace_$TAG.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
    > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err; echo "bench N=$N rc=$?"; grep -v "^\s*$" $OUT/bench_n${N}_$TAG.err | grep -n "bench rank\|Fatal\|File\|Error\|error" | head -60 | cut -c1-250
python - <<PY
import json
try:
    d = json.load(open('$OUT/bench_n${N}_$TAG.json'))
    print('c2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
    for k in ('c1', 'c3', 'c5', 'c4'):
        if k in d: print(k, d[k].get('value'), d[k].get('ms_per_step'), d[k].get('error'), {a: b for a, b in d[k].items() if a in ('kernel_ms_per_step', 'exposed_comm_ms', 'scaling')})
except Exception as e:
    print('no line', e)
PY
