#!/usr/bin/env python
"""Time line of the persistent row / column pipelines (csrc/dct_pipe.cuh) on one frame: where the CTAs of the four line
kernels of an embed spend their time -- waiting for the first tile, computing, waiting for later tiles, draining the
stores.  Uses ssw_ctx_set_trace (clock64 stamps written by the kernels themselves; the overhead is a few stores per tile).

    python tools/pipe_trace.py [W H [batch]]  > profiles/rN_pipe_trace.txt        (on the GPU box)
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spread_spectrum_watermarking_b200 as wm  # noqa: E402
from spread_spectrum_watermarking_b200._lib import check, lib, ssw_config  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2160
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
N = 1000
NAMES = ['fwd_rows', 'fwd_cols', 'inv_cols', 'inv_rows']


def main():
    torch.cuda.set_device(0)
    ctx = wm.Context(0)
    frames = torch.empty((B, H, W, 3), dtype=torch.uint8, device='cuda')
    check(lib.ssw_synth_frame_rgb8_dev(ctx.handle, W, H, 9, 0, B, frames.data_ptr()))
    out = torch.empty_like(frames)
    marks = torch.randn((B, N), device='cuda')
    cfg = ssw_config(2, 0.1, 0)
    trace = torch.zeros((16, 1024, 64), dtype=torch.int64, device='cuda')

    def embed():
        check(lib.ssw_embed_batch_rgb8_dev(ctx.handle, frames.data_ptr(), W, H, B, ctypes.byref(cfg), marks.data_ptr(), N, out.data_ptr()))

    for _ in range(5):
        embed()
    ctx.synchronize()
    torch.cuda.synchronize()
    check(lib.ssw_ctx_set_trace(ctx.handle, trace.data_ptr()))
    embed()
    ctx.synchronize()
    check(lib.ssw_ctx_set_trace(ctx.handle, None))
    t = trace.cpu().numpy()
    clk_mhz = 1965.0   # SM clock under load on this pool (bench.py clocks line); stamps are SM cycles
    us = lambda cyc: cyc / clk_mhz
    print('# pipeline time lines, %dx%d x %d frame(s), one embed (fwd_rows, fwd_cols, [ordering], inv_cols, inv_rows)' % (W, H, B))
    print('# per kernel: CTAs, tiles per CTA; then medians over the CTAs, in us (SM cycles / %.0f MHz)' % clk_mhz)
    for li, name in enumerate(NAMES):
        blk = t[li]
        used = np.nonzero(blk[:, 1])[0]
        if len(used) == 0:
            print('%s: no trace (kernel did not take the pipeline path)' % name)
            continue
        b = blk[used]
        g0 = b[:, 0].min()
        nt = b[:, 5]
        start_skew = (b[:, 0] - g0) / 1e3
        life = (b[:, 4] - b[:, 0]) / 1e3
        span = (b[:, 4].max() - g0) / 1e3
        dep_wait = us(b[:, 2] - b[:, 1])
        print('\n== %s: %d CTAs on %d SMs, tiles per CTA %s' % (name, len(used), len(set(b[:, 6].tolist())),
              dict(zip(*np.unique(nt, return_counts=True)))))
        print('   kernel span (first CTA start -> last compute end, globaltimer) %.1f us; CTA start skew median %.1f max %.1f us; '
              'CTA life median %.1f max %.1f us' % (span, np.median(start_skew), start_skew.max(), np.median(life), life.max()))
        print('   dependency wait (griddepcontrol.wait incl. table staging) median %.2f us max %.2f' % (np.median(dep_wait), dep_wait.max()))
        maxj = int(min(nt.max(), 7))
        hdr = '   tile  n_cta  load->landed  wait_for_tile  compute  done->store_issued  store_read  next_load_after_done'
        print(hdr)
        for j in range(maxj):
            m = nt > j
            r = b[m]
            base = 8 + 8 * j
            issued, landed, done, st_iss, st_rd = (r[:, base + k] for k in range(5))
            prev_done = r[:, base - 8 + 2] if j > 0 else r[:, 2]
            # how long the compute warps sat at the tile's mbarrier: landed-stamp minus the moment they were ready for it
            wait_tile = us(landed - prev_done)
            load_lat = us(landed - issued)   # upper bound: the stamp is taken when the compute warps wake up
            comp = us(done - landed)
            d2s = np.where(st_iss > 0, us(st_iss - done), np.nan)
            srd = np.where(st_rd > 0, us(st_rd - st_iss), np.nan)
            print('   %4d  %5d  %12.2f  %13.2f  %7.2f  %18.2f  %10.2f' % (j, m.sum(), np.median(load_lat), np.median(wait_tile), np.median(comp),
                  np.nanmedian(d2s) if np.isfinite(d2s).any() else float('nan'), np.nanmedian(srd) if np.isfinite(srd).any() else float('nan')))
        # one CTA with the most tiles and the CTA that lived longest, in full
        for what, k in (('example', int(np.argmax(nt))), ('longest-lived', int(np.argmax(life)))):
            r = b[k]
            print('   %s CTA %d (SM %d, %d tiles), us since its start:' % (what, used[k], r[6], r[5]))
            print('      after dependency wait %.2f' % us(r[2] - r[1]))
            for j in range(int(min(r[5], 7))):
                base = 8 + 8 * j
                ev = ['load issued', 'landed', 'compute done', 'store issued', 'store read', 'A released', 'B landed', 'histogram done']
                print('      tile %d: ' % j + ', '.join('%s %.2f' % (ev[q], us(r[base + q] - r[1])) for q in range(8) if r[base + q] > 0))
            print('      compute warps finished %.2f' % us(r[3] - r[1]))
    ctx.close()


if __name__ == '__main__':
    main()
