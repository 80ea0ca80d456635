#!/bin/bash
# column pipeline variants on one 4K frame: bench c2 + pipeline time lines
TAG=${1:-s3f}
OUT=gpurun_out; mkdir -p $OUT
for v in 1 2 3; do
  SSW_COL_PIPE=$v timeout 300 python bench.py --workload c2 --steps 100 --no-cpu-baseline --no-e2e > $OUT/bench_c2_${TAG}_p$v.json 2> $OUT/bench_c2_${TAG}_p$v.err; echo "c2 col_pipe=$v rc=$?"
  SSW_COL_PIPE=$v timeout 300 python tools/pipe_trace.py > $OUT/pipe_trace_${TAG}_p$v.txt 2> $OUT/pipe_trace_${TAG}_p$v.err; echo "trace rc=$?"
done
python tools/kernels_table.py $OUT/bench_c2_${TAG}_p*.json 2>&1 | grep -E "json|cols"
