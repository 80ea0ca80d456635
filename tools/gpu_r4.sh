#!/bin/bash
# Round-2 GPU call: parity tests, smoke, bench c2 / c3 / c5 (per-kernel table), ncu launch list + --set full capture.
# Usage: bash tools/gpu_r4.sh <tag> [tests|notests] [ncu|noncu]
TAG=${1:-r4}; TESTS=${2:-tests}; NCU=${3:-ncu}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/device_$TAG.txt 2>&1
if [ "$TESTS" == "tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu_$TAG.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
fi
for wl in c2 c3 c5; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-e2e > $OUT/bench_${wl}_$TAG.json 2> $OUT/bench_${wl}_$TAG.err; echo "bench $wl rc=$?"; tail -2 $OUT/bench_${wl}_$TAG.err
done
python tools/kernels_table.py $OUT/bench_c2_$TAG.json $OUT/bench_c3_$TAG.json $OUT/bench_c5_$TAG.json
if [ "$NCU" == "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_$TAG.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fast_kernel|col_pipe|row_pipe|topk_|similarity' -s 68 -c 24 -f -o $OUT/prof_$TAG \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
ls $OUT | wc -l
