#!/bin/bash
# column pipeline: fused collect + split schedule A/B on one 4K frame, parity subset first
TAG=${1:-s3g}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or host_batch or cat or topk or order" > $OUT/pytest_${TAG}_p1.log 2>&1; echo "pytest default rc=$?"; tail -n 4 $OUT/pytest_${TAG}_p1.log | cut -c1-300
SSW_COL_PIPE=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or host_batch or cat or topk or order" > $OUT/pytest_${TAG}_p2.log 2>&1; echo "pytest pipe2 rc=$?"; tail -n 4 $OUT/pytest_${TAG}_p2.log | cut -c1-300
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --workload c2 --steps 100 --no-cpu-baseline --no-e2e > $OUT/bench_c2_${TAG}_$name.json 2> $OUT/bench_c2_${TAG}_$name.err; echo "$name rc=$?"
}
run p1c0 SSW_COL_PIPE=1 SSW_COL_COLLECT=0
run p1c1 SSW_COL_PIPE=1 SSW_COL_COLLECT=1
run p2s0c1 SSW_COL_PIPE=2 SSW_COL_SPLIT=0 SSW_COL_COLLECT=1
run p2s1c1 SSW_COL_PIPE=2 SSW_COL_SPLIT=1 SSW_COL_COLLECT=1
run p2s1c0 SSW_COL_PIPE=2 SSW_COL_SPLIT=1 SSW_COL_COLLECT=0
run p2s0c0 SSW_COL_PIPE=2 SSW_COL_SPLIT=0 SSW_COL_COLLECT=0
python tools/kernels_table.py $OUT/bench_c2_${TAG}_p*.json 2>&1 | grep -E "json|cols|collect"
