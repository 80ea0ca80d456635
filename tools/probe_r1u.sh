#!/bin/bash
# r1u probe (tuning build): packed-FP32 issue rates + fwd_rows with two butterflies per thread (variants 8, 9)
OUT=gpurun_out; mkdir -p $OUT
timeout 120 ./tools/ubench_f32x2 > $OUT/ubench_f32x2_r1u.txt 2>&1; cat $OUT/ubench_f32x2_r1u.txt
for v in 0 8 9 2; do
  SSW_ROW_VARIANT=$v timeout 120 python - <<PY 2>&1 | tee -a $OUT/probe_rows_r1u.txt
import ctypes, sys
sys.path.insert(0, '.')
import torch, numpy as np
import spread_spectrum_watermarking_b200 as wm
from spread_spectrum_watermarking_b200._lib import lib, check
s = torch.cuda.Stream()
ctx = wm.Context(0, stream=s.cuda_stream)
w, h, B = 3840, 2160, 8
fr = torch.empty((B, h, w, 3), dtype=torch.uint8, device='cuda')
check(lib.ssw_synth_frame_rgb8_dev(ctx.handle, w, h, 2, 0, B, fr.data_ptr()))
pl = torch.empty((B, h, w), dtype=torch.float32, device='cuda')
ctx.synchronize()
for i in range(4):
    check(lib.ssw_lines_forward_dev(ctx.handle, 0, fr[i % B].data_ptr(), w, h, pl[i % B].data_ptr()))
ctx.synchronize()
ctx.profile_begin()
for i in range(40):
    check(lib.ssw_lines_forward_dev(ctx.handle, 0, fr[i % B].data_ptr(), w, h, pl[i % B].data_ptr()))
p = ctx.profile_end()
ctx.synchronize()
print('variant $v:', {k: round(v['ms'] / v['launches'] * 1e3, 2) for k, v in p.items()}, 'checksum %.6e' % float(pl[0].double().abs().sum().item()))
PY
done
