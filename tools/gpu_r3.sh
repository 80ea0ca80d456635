#!/bin/bash
# Round-2 GPU call: parity tests (optional), bench c2 / c3 / c5 with the per-kernel table.
# Usage: bash tools/gpu_r3.sh <tag> [tests|notests] [extra env assignments ...]
TAG=${1:-r3}; TESTS=${2:-tests}; shift 2
OUT=gpurun_out; mkdir -p $OUT
for kv in "$@"; do export "$kv"; done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/device_$TAG.txt 2>&1
if [ "$TESTS" == "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu_$TAG.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
fi
for wl in c2 c3 c5; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-e2e > $OUT/bench_${wl}_$TAG.json 2> $OUT/bench_${wl}_$TAG.err; echo "bench $wl rc=$?"; tail -2 $OUT/bench_${wl}_$TAG.err
done
python tools/kernels_table.py $OUT/bench_c2_$TAG.json $OUT/bench_c3_$TAG.json $OUT/bench_c5_$TAG.json
