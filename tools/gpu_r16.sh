#!/bin/bash
# Tile-wise candidate scan (per-tile coefficient maxima from the forward column pipeline): full parity suite, then A/B.
TAG=${1:-r16}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 4 $OUT/pytest_gpu_$TAG.log | cut -c1-300
run() {  # workload, name, env...
  local wl=$1 name=$2; shift; shift
  local extra="--no-extra"; [ "$wl" != "c2" ] && extra="--workload $wl --steps 20 --warmup 3"
  env "$@" timeout 300 python bench.py $extra --no-cpu-baseline --no-e2e > $OUT/ab_${TAG}_${wl}_$name.json 2> $OUT/ab_${TAG}_${wl}_$name.err; echo "$wl $name rc=$?"
}
run c2 tm1 SSW_TILE_MAX=1

run c3 tm1 SSW_TILE_MAX=1
run c3 tm0 SSW_TILE_MAX=0


python tools/kernels_table.py $OUT/ab_${TAG}_c*.json | grep -E "json|fwd_cols|collect"
