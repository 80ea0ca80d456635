#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import ssw_oracle as so
import spread_spectrum_watermarking_b200 as wm
for (w, h) in [(3840, 67), (1920, 1080), (640, 444)]:
    frame = so.synth_frame(w, h, 2)
    res = []
    for pf in ('0', '1'):
        os.environ['SSW_PREFETCH'] = pf
        ctx = wm.Context(0)
        res.append(wm.Writer.new(frame, ctx=ctx).coefficient_image())
        ctx.close()
    print(w, h, 'prefetch kernel identical to direct loader:', bool((res[0] == res[1]).all()), flush=True)
PY
for cfg in "0 0" "1 0" "1 1" "1 2" "1 4"; do set -- $cfg
  SSW_PREFETCH=$1 SSW_PF_TILES=$2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ab_pf_c2_$1_$2.json 2> $OUT/ab_pf_c2_$1_$2.err
  SSW_PREFETCH=$1 SSW_PF_TILES=$2 timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ab_pf_c3_$1_$2.json 2> $OUT/ab_pf_c3_$1_$2.err
done
python tools/kernels_table.py $OUT/ab_pf_c2_*.json $OUT/ab_pf_c3_*.json | grep -E "json|fwd_rows"
