#!/bin/bash
# Round-2 late experiment: parity tests on the new library, then A/B of library builds (tools/ab/libssw_<tag>.so; here: with and
# without nvcc --split-compile) x inverse row pipeline shapes (SSW_ROW_INPLACE) on c2 and c3.
# Usage: bash tools/gpu_r9.sh <run-tag> <lib-tag> [<lib-tag> ...]      (the in-tree libssw.so runs the tests)
TAG=${1:-r9}; shift
OUT=gpurun_out; mkdir -p $OUT
LIB=spread_spectrum_watermarking_b200/csrc/libssw.so
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/clocks_before_$TAG.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 4 $OUT/pytest_gpu_$TAG.log | cut -c1-300
cp $LIB /tmp/libssw_keep.so
for t in "$@"; do
  cp tools/ab/libssw_$t.so $LIB
  for ip in 0 1; do
    SSW_ROW_INPLACE=$ip timeout 300 python bench.py --no-extra --no-cpu-baseline --no-e2e > $OUT/ab_${TAG}_c2_${t}_ip$ip.json 2> $OUT/ab_${TAG}_c2_${t}_ip$ip.err; echo "c2 $t ip$ip rc=$?"
    SSW_ROW_INPLACE=$ip timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ab_${TAG}_c3_${t}_ip$ip.json 2> $OUT/ab_${TAG}_c3_${t}_ip$ip.err; echo "c3 $t ip$ip rc=$?"
  done
done
cp /tmp/libssw_keep.so $LIB
python tools/kernels_table.py $OUT/ab_${TAG}_c*.json
du -sh $OUT
