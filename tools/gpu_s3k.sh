#!/bin/bash
# A/B of two library builds (tools/ab/libssw_t8.so = no swizzle, in-tree = 32-byte swizzle) x column pipeline shapes; parity of the in-tree build first
TAG=${1:-s3k}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or host_batch or cat or topk or order or dct or fast" > $OUT/pytest_${TAG}.log 2>&1; echo "pytest in-tree rc=$?"; tail -n 4 $OUT/pytest_${TAG}.log | cut -c1-300
run() { local name=$1; local wl=$2; local steps=$3; shift; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --steps $steps --no-cpu-baseline --no-e2e > $OUT/bench_${wl}_${TAG}_$name.json 2> $OUT/bench_${wl}_${TAG}_$name.err; echo "$wl $name rc=$?"; }
T8=$PWD/tools/ab/libssw_t8.so
run t8p1 c2 100 SSW_LIB=$T8 SSW_COL_PIPE=1
run t8p2 c2 100 SSW_LIB=$T8 SSW_COL_PIPE=2
run t9p2 c2 100 SSW_COL_PIPE=2
run t8p1b c2 100 SSW_LIB=$T8 SSW_COL_PIPE=1
run t9p2b c2 100 SSW_COL_PIPE=2
run t8 c3 10 SSW_LIB=$T8
run t9 c3 10 A=1
python tools/kernels_table.py $OUT/bench_c2_${TAG}_t*.json $OUT/bench_c3_${TAG}_t*.json 2>&1 | grep -E "json|cols"
