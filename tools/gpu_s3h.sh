#!/bin/bash
TAG=${1:-s3h}
OUT=gpurun_out; mkdir -p $OUT
export SSW_LIB=$PWD/tools/ab/libssw_trace.so
SSW_COL_PIPE=2 SSW_COL_SPLIT=1 SSW_COL_COLLECT=0 timeout 300 python tools/pipe_trace.py > $OUT/pipe_trace_${TAG}_p2s1c0.txt 2> $OUT/pipe_trace_${TAG}_p2s1c0.err; echo "trace rc=$?"
SSW_COL_PIPE=2 SSW_COL_SPLIT=1 SSW_COL_COLLECT=1 timeout 300 python tools/pipe_trace.py > $OUT/pipe_trace_${TAG}_p2s1c1.txt 2> $OUT/pipe_trace_${TAG}_p2s1c1.err; echo "trace rc=$?"
SSW_COL_PIPE=2 SSW_COL_SPLIT=0 SSW_COL_COLLECT=0 timeout 300 python tools/pipe_trace.py > $OUT/pipe_trace_${TAG}_p2s0c0.txt 2> $OUT/pipe_trace_${TAG}_p2s0c0.err; echo "trace rc=$?"
tail -n 3 $OUT/pipe_trace_${TAG}_p2s1c0.err
