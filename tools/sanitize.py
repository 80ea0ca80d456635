"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool racecheck python tools/sanitize.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import ssw_oracle as so  # noqa: E402
import spread_spectrum_watermarking_b200 as wm  # noqa: E402

ctx = wm.Context(0)
rng = np.random.default_rng(0)
# fast pair plans (rows 640/1024, cols 1080/640), generic kernels (444, 37), single-line kernels (force via env in a 2nd run)
for (w, h) in [(640, 1080), (1024, 640), (640, 444), (52, 37), (1920, 64)]:
    frame = so.synth_frame(w, h, 3)
    mark = rng.standard_normal(200).astype(np.float32)
    out = wm.Writer.new(frame, ctx=ctx).mark_rgb8([mark, mark[::-1].copy()])
    e = wm.Reader.base(frame, ctx=ctx).extract(wm.Reader.derived(out, ctx=ctx), 200)
    s = wm.Tester.new(e, ctx=ctx).similarity(mark)
    print(w, h, 'similarity %.2f' % float(s.similarity), flush=True)
bank = wm.Bank.normal(3, 300, 200, ctx=ctx)
print('bank', bank.similarity(e).shape)
bank.close()
import ctypes
import torch
cfg = wm._lib.ssw_config(2, 0.1, 0)
fr = torch.from_numpy(np.stack([so.synth_frame(640, 1080, 3, i) for i in range(2)])).cuda()
mk = torch.from_numpy(rng.standard_normal((2, 200)).astype(np.float32)).cuda()
out = torch.empty_like(fr); ext = torch.empty((2, 200), device='cuda'); sim = torch.empty(2, device='cuda')
torch.cuda.synchronize()
wm._lib.check(wm.lib.ssw_embed_batch_rgb8_dev(ctx.handle, fr.data_ptr(), 640, 1080, 2, ctypes.byref(cfg), mk.data_ptr(), 200, out.data_ptr()))
wm._lib.check(wm.lib.ssw_extract_batch_rgb8_dev(ctx.handle, fr.data_ptr(), out.data_ptr(), 640, 1080, 2, ctypes.byref(cfg), 200,
                                                ext.data_ptr(), mk.data_ptr(), sim.data_ptr()))
ctx.synchronize()
print('fused', sim.cpu().numpy())
from spread_spectrum_watermarking_b200 import sharded
ops = sharded.CudaOps(0)
rows = torch.from_numpy(so.synth_frame(1024, 64, 5)).cuda()
torch.cuda.synchronize()
wr = sharded.ShardedWriter(rows, 1024, 64, cfg, ops, rank=0, world=1)
o = wr.mark_rgb8([rng.standard_normal(100).astype(np.float32)])
ops.synchronize()
print('sharded', o.shape)
del wr
ctx.close()
print('SANITIZE_RUN_OK')
