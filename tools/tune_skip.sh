#!/bin/bash
# tuning build: fwd_rows with loads / stores / FFT stages switched off (results are garbage; timing only)
OUT=gpurun_out; mkdir -p $OUT
for v in 10 11 12 13 14 15 17; do
  SSW_ROW_VARIANT=$v timeout 120 python - <<PY
import ctypes, sys
sys.path.insert(0, '.')
import torch
import spread_spectrum_watermarking_b200 as wm
from spread_spectrum_watermarking_b200._lib import lib, check
s = torch.cuda.Stream()
ctx = wm.Context(0, stream=s.cuda_stream)
w, h, B = 3840, 2160, 8
fr = torch.empty((B, h, w, 3), dtype=torch.uint8, device='cuda')
check(lib.ssw_synth_frame_rgb8_dev(ctx.handle, w, h, 2, 0, B, fr.data_ptr()))
pl = torch.empty((B, h, w), dtype=torch.float32, device='cuda')
ctx.synchronize()
ctx.profile_begin()
for i in range(24):
    check(lib.ssw_lines_forward_dev(ctx.handle, 0, fr[i % B].data_ptr(), w, h, pl[i % B].data_ptr()))
p = ctx.profile_end()
print('variant $v (skip mask %d):' % ($v - 10), {k: round(v['ms'] / v['launches'] * 1e3, 2) for k, v in p.items()})
PY
done
