#!/bin/bash
# A/B of prebuilt library variants (tools/ab/libssw_<tag>.so): bench c2 (+ c3) per variant, per-kernel table
# Usage: bash tools/ab_libs.sh <run-tag> <tag> [<tag> ...]
OUT=gpurun_out; mkdir -p $OUT
RUN=$1; shift
LIB=spread_spectrum_watermarking_b200/csrc/libssw.so
cp $LIB /tmp/libssw_keep.so
for t in "$@"; do
  cp tools/ab/libssw_$t.so $LIB
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/ab_${RUN}_c2_$t.json 2> $OUT/ab_${RUN}_c2_$t.err
  timeout 300 python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ab_${RUN}_c3_$t.json 2> $OUT/ab_${RUN}_c3_$t.err
done
cp /tmp/libssw_keep.so $LIB
python tools/kernels_table.py $OUT/ab_${RUN}_c*.json | grep -E "json|rows|cols|topk|sim"
