#!/bin/bash
# Relief schedule for the histogram CTAs of the forward column pipeline: parity of the pipeline variants, then A/B on c2 / c5.
TAG=${1:-r15}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipeline_variants or partial_inverse or fused_batch or cat_" > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 4 $OUT/pytest_gpu_$TAG.log | cut -c1-300
run() {  # workload, name, env...
  local wl=$1 name=$2; shift; shift
  local extra="--no-extra"; [ "$wl" != "c2" ] && extra="--workload $wl --steps 30 --warmup 3"
  env "$@" timeout 300 python bench.py $extra --no-cpu-baseline --no-e2e > $OUT/ab_${TAG}_${wl}_$name.json 2> $OUT/ab_${TAG}_${wl}_$name.err; echo "$wl $name rc=$?"
}
run c2 relief1 SSW_HIST_RELIEF=1
run c2 relief0 SSW_HIST_RELIEF=0
run c2 relief1_p4 SSW_HIST_RELIEF=1 SSW_COL_PIPE=4
run c2 relief0_p4 SSW_HIST_RELIEF=0 SSW_COL_PIPE=4
run c5 relief1 SSW_HIST_RELIEF=1
run c5 relief0 SSW_HIST_RELIEF=0
python tools/kernels_table.py $OUT/ab_${TAG}_c*.json | grep -E "json|fwd_cols"
