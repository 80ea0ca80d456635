#!/bin/bash
# Parity tests on the library with the partial inverse column pass + the histogram in the producer warp, then A/B on c2 / c3
# and time lines.  Usage: bash tools/gpu_r12.sh <tag>
TAG=${1:-r12}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 6 $OUT/pytest_gpu_$TAG.log | cut -c1-400
run() {  # workload, name, env...
  local wl=$1 name=$2; shift; shift
  local extra="--no-extra"; [ "$wl" != "c2" ] && extra="--workload $wl --steps 10 --warmup 3"
  env "$@" timeout 300 python bench.py $extra --no-cpu-baseline --no-e2e > $OUT/ab_${TAG}_${wl}_$name.json 2> $OUT/ab_${TAG}_${wl}_$name.err; echo "$wl $name rc=$?"; tail -n 1 $OUT/ab_${TAG}_${wl}_$name.err | cut -c1-200
}
run c2 default SSW_DUMMY=1
run c2 nohist SSW_COL_HIST=0
run c2 pipe2 SSW_COL_PIPE=2
run c2 pipe2_nohist SSW_COL_PIPE=2 SSW_COL_HIST=0
run c2 full SSW_PARTIAL_INV=0
run c3 default SSW_DUMMY=1
run c5 default SSW_DUMMY=1
run c5 nohist SSW_COL_HIST=0
python tools/kernels_table.py $OUT/ab_${TAG}_c*.json
if [ -f tools/ab/libssw_trace.so ]; then
  SSW_LIB=tools/ab/libssw_trace.so timeout 300 python tools/pipe_trace.py > $OUT/pipe_trace_${TAG}_default.txt 2> $OUT/pipe_trace_${TAG}_default.err; echo "trace rc=$?"
  SSW_LIB=tools/ab/libssw_trace.so SSW_COL_PIPE=2 timeout 300 python tools/pipe_trace.py > $OUT/pipe_trace_${TAG}_pipe2.txt 2> $OUT/pipe_trace_${TAG}_pipe2.err; echo "trace rc=$?"
fi
du -sh $OUT
