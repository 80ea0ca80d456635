#!/bin/bash
# sharded frame on N GPUs: parity of the in-tree library, then c4 A/B (previous library / in-tree / in-tree with 8 slices)
TAG=${1:-c4}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 tests/dist/run_sharded_cabi.py 4096 4096 1000 > $OUT/cabi_n${N}_$TAG.log 2>&1; echo "parity 4096 rc=$?"; grep -E "sharded|SHARDED" $OUT/cabi_n${N}_$TAG.log | cut -c1-250
run() { local name=$1; shift
  env "$@" timeout 300 $TR --master-port 29532 bench.py --gpus $N --workload c4 > $OUT/bench_c4_n${N}_${TAG}_$name.json 2> $OUT/bench_c4_n${N}_${TAG}_$name.err; echo "c4 $name rc=$?"; }
run old SSW_LIB=$PWD/tools/ab/libssw_t10.so
run new A=1
run new8 SSW_SHARD_CHUNKS=8
run new2 SSW_SHARD_CHUNKS=2
python - <<PY
import json
for n in ('old', 'new', 'new8', 'new2'):
    try:
        d = json.load(open('$OUT/bench_c4_n${N}_${TAG}_%s.json' % n))
        print(n, round(d['value']), 'Mpix/s', round(d['ms_per_step'], 3), 'ms; kernels', round(d['kernel_ms_per_step'], 3), 'exposed', round(d['exposed_comm_ms_per_step'], 3), 'e2e', round(d['e2e']['value']))
        for e in d['kernels'][:5]: print('   %-18s x%.1f %9.2f us share %.3f frac %s' % (e['name'], e['launches_per_step'], e['avg_us'], e['share'], e['frac']))
    except Exception as e:
        print(n, 'no line', e)
PY
