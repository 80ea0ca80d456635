#!/bin/bash
# r7: low-rank A/B (c2, c3), parity tests, tensor-core DCT probe
TAG=${1:-r7}
OUT=gpurun_out; mkdir -p $OUT
for v in 0 1; do
  SSW_LOWRANK=$v timeout 300 python bench.py --workload c2 --steps 100 --no-cpu-baseline --no-e2e > $OUT/bench_c2_${TAG}_l$v.json 2> $OUT/bench_c2_${TAG}_l$v.err; echo "c2 lowrank=$v rc=$?"; tail -n 2 $OUT/bench_c2_${TAG}_l$v.err
  SSW_LOWRANK=$v timeout 300 python bench.py --workload c3 --steps 10 --no-cpu-baseline --no-e2e > $OUT/bench_c3_${TAG}_l$v.json 2> $OUT/bench_c3_${TAG}_l$v.err; echo "c3 lowrank=$v rc=$?"; tail -n 2 $OUT/bench_c3_${TAG}_l$v.err
done
python tools/kernels_table.py $OUT/bench_c2_${TAG}_l*.json $OUT/bench_c3_${TAG}_l*.json 2>&1 | grep -E "json|lowrank|inv_"
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 12 $OUT/pytest_gpu_$TAG.log
timeout 300 python tools/tc_dct_probe.py > $OUT/tensor_core_dct_$TAG.json 2> $OUT/tensor_core_dct_$TAG.err; echo "tc probe rc=$?"; cat $OUT/tensor_core_dct_$TAG.json; tail -n 3 $OUT/tensor_core_dct_$TAG.err
