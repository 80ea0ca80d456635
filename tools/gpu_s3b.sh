#!/bin/bash
# ncu launch list + full capture of the line kernels (c2, one frame); small outputs only
TAG=${1:-s3b}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --no-extra --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
tail -n 5 $OUT/ncu_list_$TAG.log | cut -c1-400
wc -l $OUT/launches_$TAG.csv
bash tools/gpu_ncu.sh $TAG c2 16 8 'row_pipe|col_pipe|topk_'
tail -n 8 $OUT/ncu_full_$TAG.log | cut -c1-300
rm -f $OUT/*.ncu-rep
du -sh $OUT
