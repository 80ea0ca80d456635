#!/bin/bash
# tuning build (-DSSW_TUNE): sweep the 3840-point row-plan variants: correctness vs variant 0, then bench c2
OUT=gpurun_out; mkdir -p $OUT
python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import ssw_oracle as so
frame = so.synth_frame(3840, 64, 2)
ref = None
for v in range(8):
    os.environ['SSW_ROW_VARIANT'] = str(v)
    import importlib
    import spread_spectrum_watermarking_b200 as wm
    ctx = wm.Context(0)
    c = wm.Writer.new(frame, ctx=ctx).coefficient_image()
    if ref is None:
        ref = c
    print('variant', v, 'max rel diff vs v0 %.3g' % (np.abs(c - ref).max() / np.abs(ref).max()), flush=True)
    ctx.close()
PY
for v in 0 1 2 3 4 5 6 7; do
  SSW_ROW_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/tune_rows_v$v.json 2> $OUT/tune_rows_v$v.err
done
python tools/kernels_table.py $OUT/tune_rows_v*.json | grep -E "json|fwd_rows"
