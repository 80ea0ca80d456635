#!/bin/bash
# r8: tensor-core variant of the low-rank apply kernel: A/B on c2 / c3 + ncu of both apply kernels
TAG=${1:-r8}
OUT=gpurun_out; mkdir -p $OUT
for v in 0 1; do
  SSW_LOWRANK_MMA=$v timeout 300 python bench.py --workload c2 --steps 100 --no-cpu-baseline --no-e2e > $OUT/bench_c2_${TAG}_m$v.json 2> $OUT/bench_c2_${TAG}_m$v.err; echo "c2 mma=$v rc=$?"; tail -n 2 $OUT/bench_c2_${TAG}_m$v.err
  SSW_LOWRANK_MMA=$v timeout 300 python bench.py --workload c3 --steps 10 --no-cpu-baseline --no-e2e > $OUT/bench_c3_${TAG}_m$v.json 2> $OUT/bench_c3_${TAG}_m$v.err; echo "c3 mma=$v rc=$?"; tail -n 2 $OUT/bench_c3_${TAG}_m$v.err
done
python tools/kernels_table.py $OUT/bench_c2_${TAG}_m*.json $OUT/bench_c3_${TAG}_m*.json 2>&1 | grep -E "json|lowrank"
SSW_LOWRANK_MMA=0 bash tools/gpu_ncu.sh ${TAG}a c2 2 2 'lowrank'
SSW_LOWRANK_MMA=1 bash tools/gpu_ncu.sh ${TAG}b c2 2 2 'lowrank'
for f in $OUT/ncu_source_${TAG}a_1.txt $OUT/ncu_source_${TAG}b_1.txt; do echo $f; head -n 40 $f | cut -c1-180; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lowrank or fused or host_batch" > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 6 $OUT/pytest_gpu_$TAG.log
