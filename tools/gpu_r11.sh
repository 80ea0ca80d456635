#!/bin/bash
# Parity tests on the library with the histogram in the producer warp, then A/B of the column pipeline shapes / the staggered
# first loads on c2 (c3 for the 1080-point shapes) + time lines.  Usage: bash tools/gpu_r11.sh <tag>
TAG=${1:-r11}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 4 $OUT/pytest_gpu_$TAG.log | cut -c1-300
run() {  # workload, name, env...
  local wl=$1 name=$2; shift; shift
  local extra="--no-extra"; [ "$wl" != "c2" ] && extra="--workload $wl --steps 10 --warmup 3"
  env "$@" timeout 300 python bench.py $extra --no-cpu-baseline --no-e2e > $OUT/ab_${TAG}_${wl}_$name.json 2> $OUT/ab_${TAG}_${wl}_$name.err; echo "$wl $name rc=$?"
}
run c2 default SSW_DUMMY=1
run c2 pipe2 SSW_COL_PIPE=2
run c2 stagger SSW_COL_STAGGER=1
run c2 pipe2_stagger SSW_COL_PIPE=2 SSW_COL_STAGGER=1
run c2 pipe2_stagger_ip SSW_COL_PIPE=2 SSW_COL_STAGGER=1 SSW_ROW_INPLACE=1
run c3 default SSW_DUMMY=1
run c3 pipe2 SSW_COL_PIPE=2
run c3 pipe2_stagger SSW_COL_PIPE=2 SSW_COL_STAGGER=1
python tools/kernels_table.py $OUT/ab_${TAG}_c*.json
if [ -f tools/ab/libssw_trace.so ]; then
  SSW_LIB=tools/ab/libssw_trace.so timeout 300 python tools/pipe_trace.py > $OUT/pipe_trace_${TAG}_default.txt 2> $OUT/pipe_trace_${TAG}_default.err; echo "trace rc=$?"
  SSW_LIB=tools/ab/libssw_trace.so SSW_COL_PIPE=2 SSW_COL_STAGGER=1 timeout 300 python tools/pipe_trace.py > $OUT/pipe_trace_${TAG}_pipe2_stagger.txt 2> $OUT/pipe_trace_${TAG}_pipe2_stagger.err; echo "trace rc=$?"
fi
du -sh $OUT
