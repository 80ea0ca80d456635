#!/bin/bash
# A/B of the run-time shapes of the column pipelines on the spill-free (no --split-compile) build + pipeline time lines.
# Usage: bash tools/gpu_r10.sh <tag>      (needs tools/ab/libssw_trace.so = a -DSSW_TRACE build for the time lines)
TAG=${1:-r10}
OUT=gpurun_out; mkdir -p $OUT
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --no-extra --no-cpu-baseline --no-e2e > $OUT/ab_${TAG}_c2_$name.json 2> $OUT/ab_${TAG}_c2_$name.err; echo "c2 $name rc=$?"
}
run default SSW_DUMMY=1
run nosplit SSW_COL_SPLIT=0
run pipe2 SSW_COL_PIPE=2
run pipe2_nosplit SSW_COL_PIPE=2 SSW_COL_SPLIT=0
run nohist SSW_COL_HIST=0
run pdl0 SSW_PDL_MODE=0
SSW_COL_PIPE=2 timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ab_${TAG}_c3_pipe2.json 2> $OUT/ab_${TAG}_c3_pipe2.err; echo "c3 pipe2 rc=$?"
python tools/kernels_table.py $OUT/ab_${TAG}_c*.json
if [ -f tools/ab/libssw_trace.so ]; then
  SSW_LIB=tools/ab/libssw_trace.so timeout 300 python tools/pipe_trace.py > $OUT/pipe_trace_${TAG}_default.txt 2> $OUT/pipe_trace_${TAG}_default.err; echo "trace rc=$?"
  SSW_LIB=tools/ab/libssw_trace.so SSW_COL_SPLIT=0 timeout 300 python tools/pipe_trace.py > $OUT/pipe_trace_${TAG}_nosplit.txt 2> $OUT/pipe_trace_${TAG}_nosplit.err; echo "trace rc=$?"
  SSW_LIB=tools/ab/libssw_trace.so SSW_ROW_INPLACE=1 timeout 300 python tools/pipe_trace.py > $OUT/pipe_trace_${TAG}_inplace.txt 2> $OUT/pipe_trace_${TAG}_inplace.err; echo "trace rc=$?"
fi
du -sh $OUT
