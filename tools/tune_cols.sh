#!/bin/bash
# tuning build (-DSSW_TUNE): sweep the column-pass variants on both bench workloads
OUT=gpurun_out; mkdir -p $OUT
for v in 0 1 2 3 4 5 6 7; do
  SSW_COL_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/tune_c2_v$v.json 2> $OUT/tune_c2_v$v.err
  SSW_COL_VARIANT=$v timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/tune_c3_v$v.json 2> $OUT/tune_c3_v$v.err
done
python tools/kernels_table.py $OUT/tune_c2_v*.json $OUT/tune_c3_v*.json | grep -E "json|cols|rows" 
