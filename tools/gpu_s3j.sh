#!/bin/bash
# split schedule on the 4-team and the 2-team column pipelines, one 4K frame; parity subset first
TAG=${1:-s3j}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or host_batch or cat or topk or order or dct" > $OUT/pytest_${TAG}.log 2>&1; echo "pytest default rc=$?"; tail -n 4 $OUT/pytest_${TAG}.log | cut -c1-300
run() { local name=$1; shift
  env "$@" timeout 300 python bench.py --workload c2 --steps 100 --no-cpu-baseline --no-e2e > $OUT/bench_c2_${TAG}_$name.json 2> $OUT/bench_c2_${TAG}_$name.err; echo "$name rc=$?"; }
run p1s0 SSW_COL_PIPE=1 SSW_COL_SPLIT=0
run p1s1 SSW_COL_PIPE=1 SSW_COL_SPLIT=1
run p2s1 SSW_COL_PIPE=2 SSW_COL_SPLIT=1
run p1s1b SSW_COL_PIPE=1 SSW_COL_SPLIT=1
run p1s0b SSW_COL_PIPE=1 SSW_COL_SPLIT=0
python tools/kernels_table.py $OUT/bench_c2_${TAG}_p*.json 2>&1 | grep -E "json|cols"
