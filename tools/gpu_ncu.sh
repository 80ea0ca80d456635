#!/bin/bash
# ncu --set full of the four line kernels (one launch each) inside bench.py <workload>; summaries are produced on the box
# Usage: bash tools/gpu_ncu.sh <tag> <workload> [skip] [count] [kernel regex]
TAG=${1:-n1}; WL=${2:-c3}; SKIP=${3:-16}; CNT=${4:-4}; KRE=${5:-row_pipe|col_pipe|fast_kernel}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $CNT -f -o $OUT/prof_$TAG \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py $OUT/prof_$TAG.ncu-rep $OUT/ncu_summary_$TAG.md > /dev/null 2>&1
cat $OUT/ncu_summary_$TAG.md | cut -c1-900
for i in $(seq 0 $((CNT-1))); do
  python tools/ncu_source.py $OUT/prof_$TAG.ncu-rep $i 16 > $OUT/ncu_source_${TAG}_$i.txt 2>&1
done
ls -la $OUT/prof_$TAG.ncu-rep
SZ=$(stat -c %s $OUT/prof_$TAG.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 40000000 ] && [ -z "$KEEP_REP" ]; then rm -f $OUT/prof_$TAG.ncu-rep; echo "rep removed (too large)"; fi
