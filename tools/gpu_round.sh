#!/bin/bash
# One gpurun call: parity tests, smoke, bench (c2 default + c3 + c5 [+ reference arm]), ncu launch list + full capture.
# Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh <tag> [ref] [skiptests]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python -c "import torch; print(torch.cuda.get_device_name(0))" > $OUT/device_$TAG.txt 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/clocks_before_$TAG.csv 2>&1
if [ "$3" != "skiptests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
  tail -5 $OUT/pytest_gpu_$TAG.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
fi
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -3 $OUT/bench_$TAG.err
timeout 600 python bench.py --workload c3 --no-cpu-baseline > $OUT/bench_c3_$TAG.json 2> $OUT/bench_c3_$TAG.err; echo "bench c3 rc=$?"; tail -3 $OUT/bench_c3_$TAG.err
timeout 600 python bench.py --workload c5 --no-cpu-baseline > $OUT/bench_c5_$TAG.json 2> $OUT/bench_c5_$TAG.err; echo "bench c5 rc=$?"; tail -3 $OUT/bench_c5_$TAG.err
python tools/kernels_table.py $OUT/bench_$TAG.json $OUT/bench_c3_$TAG.json $OUT/bench_c5_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fast_kernel|topk_|similarity' -s 60 -c 20 -f -o $OUT/prof_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
if [ "$2" == "ref" ]; then
  SSW_REF_BUDGET_S=60 timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"; cat $OUT/bench_ref_$TAG.json
fi
ls $OUT | wc -l
