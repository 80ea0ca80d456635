#!/bin/bash
# A/B of an environment switch of libssw on bench c2 / c3: bash tools/ab_env.sh <run-tag> <VAR> <value> [<value> ...]
OUT=gpurun_out; mkdir -p $OUT
RUN=$1; VAR=$2; shift; shift
for v in "$@"; do
  env $VAR=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > $OUT/ab_${RUN}_c2_$v.json 2> $OUT/ab_${RUN}_c2_$v.err
  env $VAR=$v timeout 300 python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ab_${RUN}_c3_$v.json 2> $OUT/ab_${RUN}_c3_$v.err
  env $VAR=$v timeout 300 python bench.py --workload c5 --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/ab_${RUN}_c5_$v.json 2> $OUT/ab_${RUN}_c5_$v.err
done
python tools/kernels_table.py $OUT/ab_${RUN}_c*.json | grep -E "json"
