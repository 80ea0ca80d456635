#!/bin/bash
# N-GPU box: multi-rank parity tests of the sharded C-ABI path + the driver's default bench invocation at N (includes c4)
# Usage: bash tools/gpu_multi.sh <tag> <N> [full]
TAG=${1:-m2}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "multi_rank or gigapixel" > $OUT/pytest_multi_$TAG.log 2>&1; echo "pytest multi rc=$?"
tail -n 15 $OUT/pytest_multi_$TAG.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
    > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err; echo "bench N=$N rc=$?"; tail -n 5 $OUT/bench_n${N}_$TAG.err | cut -c1-300
python - <<PY
import json
try:
    d = json.load(open('$OUT/bench_n${N}_$TAG.json'))
    print('c2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
    for k in ('c1', 'c3', 'c5', 'c4'):
        if k in d: print(k, d[k].get('value'), d[k].get('ms_per_step'), {a: b for a, b in d[k].items() if a in ('kernel_ms_per_step', 'exposed_comm_ms', 'scaling')})
except Exception as e:
    print('no line', e)
PY
if [ "$3" == "full" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c4 \
      > $OUT/bench_c4_n${N}_$TAG.json 2> $OUT/bench_c4_n${N}_$TAG.err; echo "bench c4 N=$N rc=$?"
fi
du -sh $OUT
