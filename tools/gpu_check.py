"""Development aid: run the whole path on a GPU and print diagnostics for every stage without
stopping at the first mismatch (pytest -m gpu is the gate; this is the microscope).
Usage (on a GPU box):  python tools/gpu_check.py [quick]"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import ssw_oracle as so  # noqa: E402
import spread_spectrum_watermarking_b200 as wm  # noqa: E402

G = os.path.join(ROOT, 'tests', 'golden')
FAILS = []


def step(name):
    def deco(fn):
        t = time.time()
        try:
            fn()
            print('[ok  ] %-28s %.2fs' % (name, time.time() - t), flush=True)
        except Exception:
            FAILS.append(name)
            print('[FAIL] %s' % name, flush=True)
            traceback.print_exc()
        return fn
    return deco


ctx = wm.Context(0)
rng = np.random.default_rng(0)


@step('dct kat 3x3')
def _():
    a = np.array([1, 0, 0, 2, 0, 0, 0, 0, 3], np.float32)
    wm.dct2d.dct2_2d(wm.dct2d.Type.DCT2, 3, 3, a, ctx)
    print('   ', a)
    assert np.allclose(a, [24, 0, 12, -6.92820323, 12, -3.46410162, 0, -10.3923048, 0], atol=1e-4)


@step('dct sizes vs oracle')
def _():
    for (h, w) in [(1, 1), (5, 4), (4, 5), (7, 9), (37, 12), (444, 640), (64, 64), (1080, 1920), (2160, 3840), (100, 4099 // 4)]:
        a = rng.random((h, w)).astype(np.float32)
        f = a.copy().ravel()
        wm.dct2d.dct2_2d(0, w, h, f, ctx)
        ref = so.dct2_2d(a, so.DCT2)
        e1 = np.abs(f.reshape(h, w) - ref).max() / np.abs(ref).max()
        b = ref.astype(np.float32).ravel().copy()
        wm.dct2d.dct2_2d(2, w, h, b, ctx)
        e2 = np.abs(b.reshape(h, w) - a).max()
        o = a.copy().ravel()
        wm.dct2d.dct2_2d(1, w, h, o, ctx)
        e3 = np.abs(o.reshape(h, w) - so.dct2_2d(a, so.DCT2_ORTHO)).max()
        print('    %5dx%-5d fwd %.2e  inv %.2e  ortho %.2e' % (w, h, e1, e2, e3), flush=True)
        assert e1 < 1e-6 and e2 < 1e-5 and e3 < 1e-5


cat = np.load(os.path.join(G, 'cat_rgb8.npz'))['rgb']
gold = np.load(os.path.join(G, 'watermarked_with_1.npz'))['rgb']
marks = np.load(os.path.join(G, 'marks.npz'))
ora = np.load(os.path.join(G, 'cat_oracle.npz'))


@step('cat forward + topk')
def _():
    w = wm.Writer.new(cat, ctx=ctx)
    c = w.coefficient_image()
    ref, _, _ = so.forward(cat)
    print('    coeff max rel-to-max err %.3e  DC %.4f (ref %.4f)' % (np.abs(c - ref).max() / np.abs(ref).max(), c[0, 0], ref[0, 0]))
    idx = w.indices(1000)
    print('    idx match oracle: %.4f, first mismatch at %s' % ((idx == ora['top_idx']).mean(),
          np.flatnonzero(idx != ora['top_idx'])[:5]))
    mine = so.obtain_indices(c.ravel(), k=1000)
    print('    idx match own-coefficient ordering (must be 1.0): %.4f' % (idx == mine).mean())
    assert (idx == mine).all()


@step('cat golden embed')
def _():
    out = wm.Writer.new(cat, ctx=ctx).mark_rgb8([marks['seed_1']])
    d = np.abs(out.astype(int) - gold.astype(int))
    print('    vs golden PNG: %d of %d differ, max %d' % ((d > 0).sum(), d.size, d.max()))
    assert d.max() <= 1
    out32 = wm.Writer.new(cat, ctx=ctx).mark([marks['seed_1']])
    print('    rgb32f range', out32.min(), out32.max())


@step('cat extract + similarity')
def _():
    r = wm.Reader.base(cat, ctx=ctx)
    d = wm.Reader.derived(gold, ctx=ctx)
    e = r.extract(d, 1000)
    m = marks['seed_1']
    sim = wm.Tester.new(e, ctx=ctx).similarity(m)
    print('    sim %.4f (oracle 31.8876), max err %.4f mean err %.4f, vs oracle extract %.2e' % (
        float(sim.similarity), np.abs(e - m).max(), np.abs(e - m).mean(), np.abs(e - ora['extracted']).max()))
    assert float(sim.similarity) == float(so.similarity(e, m)), 'similarity must be bit-identical to the sequential loop'
    rs = wm.Tester.new(e, ctx=ctx).similarity(marks['seed_baaaaaad'])
    print('    random mark sim %.4f' % float(rs.similarity))


@step('multi-mark + options + orderings')
def _():
    f = so.synth_frame(200, 120, 11)
    ms = [rng.standard_normal(50).astype(np.float32), rng.standard_normal(30).astype(np.float32)]
    for method in (1, 2, 3):
        for ordering in (0, 1, 2):
            ins = {1: wm.Insertion.Option1, 2: wm.Insertion.Option2, 3: wm.Insertion.Option3}[method](0.1)
            w = wm.Writer.new(f, wm.WriteConfig(ins, ordering), ctx=ctx)
            c0 = w.coefficient_image().ravel()
            idx = w.indices(50)
            ref_idx = so.obtain_indices(c0, ordering, 200, 120, k=50)
            w.embed(ms)
            c1 = w.coefficient_image().ravel()
            ref = so.embed_watermark(c0, ref_idx, ms, method, 0.1)
            print('    method %d ordering %d idx ok %s, embed max diff %.3e' % (method, ordering, (idx == ref_idx).all(), np.abs(c1 - ref).max()))
            assert (idx == ref_idx).all()
            assert np.abs(c1 - ref).max() <= (2e-6 * np.abs(ref).max() if method == 3 else 0)


@step('general top-k path (big k, ties)')
def _():
    f = so.synth_frame(160, 96, 5)
    r = wm.Reader.base(f, ctx=ctx)
    c = r.coefficients()
    idx = r.indices()  # all w*h-1
    ref = so.obtain_indices(c)
    print('    full ordering match: %.5f' % (idx == ref).mean())
    assert (idx == ref).all()
    flat = np.full((64, 64, 3), 128, np.uint8)
    r2 = wm.Reader.base(flat, ctx=ctx)
    c2 = r2.coefficients()
    i2 = r2.indices(100)
    ref2 = so.obtain_indices(c2, k=100)
    print('    flat image (ties) match: %.4f, fallbacks %d' % ((i2 == ref2).mean(), ctx.last_topk_fallbacks()))
    assert (i2 == ref2).all()


@step('bank similarity + normal marks')
def _():
    bank = rng.standard_normal((1000, 1000)).astype(np.float32)
    e = rng.standard_normal((3, 1000)).astype(np.float32)
    b = wm.Bank(bank, ctx=ctx)
    s = b.similarity(e)
    ref = np.array([[so.similarity(e[i], bank[j]) for j in range(0, 1000, 97)] for i in range(3)])
    print('    bank sim max abs diff vs sequential oracle: %.3e' % np.abs(s[:, ::97] - ref).max())
    assert (s[:, ::97] == ref).all()
    m = wm.MarkBuf.generate_normal(1 << 20, seed=42, ctx=ctx).data()
    print('    normal: mean %.4f std %.4f min %.2f max %.2f' % (m.mean(), m.std(), m.min(), m.max()))
    assert abs(m.mean()) < 5e-3 and abs(m.std() - 1) < 5e-3


@step('synthetic frame generator')
def _():
    import torch
    w, h = 256, 144
    t = torch.empty((2, h, w, 3), dtype=torch.uint8, device='cuda')
    wm._lib.check(wm.lib.ssw_synth_frame_rgb8_dev(ctx.handle, w, h, 9, 3, 2, t.data_ptr()))
    ctx.synchronize()
    a = t.cpu().numpy()
    for i in range(2):
        assert (a[i] == so.synth_frame(w, h, 9, 3 + i)).all(), 'synth frame %d differs' % i


@step('fused batch embed/extract 1080p x4')
def _():
    import torch
    w, h, B, n = 1920, 1080, 4, 1000
    frames = torch.empty((B, h, w, 3), dtype=torch.uint8, device='cuda')
    wm._lib.check(wm.lib.ssw_synth_frame_rgb8_dev(ctx.handle, w, h, 3, 0, B, frames.data_ptr()))
    mk = torch.from_numpy(rng.standard_normal((B, n)).astype(np.float32)).cuda()
    out = torch.empty_like(frames)
    cfg = wm._lib.ssw_config(2, 0.1, 0)
    import ctypes
    torch.cuda.synchronize()
    wm._lib.check(wm.lib.ssw_embed_batch_rgb8_dev(ctx.handle, frames.data_ptr(), w, h, B, ctypes.byref(cfg), mk.data_ptr(), n, out.data_ptr()))
    ext = torch.empty((B, n), dtype=torch.float32, device='cuda')
    sim = torch.empty((B,), dtype=torch.float32, device='cuda')
    wm._lib.check(wm.lib.ssw_extract_batch_rgb8_dev(ctx.handle, frames.data_ptr(), out.data_ptr(), w, h, B, ctypes.byref(cfg), n, ext.data_ptr(), mk.data_ptr(), sim.data_ptr()))
    ctx.synchronize()
    print('    sims', sim.cpu().numpy(), 'fallbacks', ctx.last_topk_fallbacks())
    f0 = frames[0].cpu().numpy()
    ref_img, ref_idx, _ = so.embed(f0, [mk[0].cpu().numpy()])
    d = np.abs(out[0].cpu().numpy().astype(int) - ref_img.astype(int))
    print('    image 0 vs oracle: %d differ (max %d)' % ((d > 0).sum(), d.max()))
    assert d.max() <= 1
    # timing
    for name, fn in (('embed', lambda: wm.lib.ssw_embed_batch_rgb8_dev(ctx.handle, frames.data_ptr(), w, h, B, ctypes.byref(cfg), mk.data_ptr(), n, out.data_ptr())),
                     ('extract', lambda: wm.lib.ssw_extract_batch_rgb8_dev(ctx.handle, frames.data_ptr(), out.data_ptr(), w, h, B, ctypes.byref(cfg), n, ext.data_ptr(), mk.data_ptr(), sim.data_ptr()))):
        for _ in range(3):
            fn()
        ctx.synchronize()
        t0 = time.time()
        for _ in range(10):
            fn()
        ctx.synchronize()
        dt = (time.time() - t0) / 10
        print('    %s: %.3f ms per batch of %d -> %.0f Mpix/s' % (name, dt * 1e3, B, B * w * h / dt / 1e6))


if 'quick' not in sys.argv:
    @step('4K single frame timing')
    def _():
        import ctypes
        import torch
        w, h, n = 3840, 2160, 1000
        frames = torch.empty((1, h, w, 3), dtype=torch.uint8, device='cuda')
        wm._lib.check(wm.lib.ssw_synth_frame_rgb8_dev(ctx.handle, w, h, 2, 0, 1, frames.data_ptr()))
        mk = torch.from_numpy(rng.standard_normal((1, n)).astype(np.float32)).cuda()
        out = torch.empty_like(frames)
        cfg = wm._lib.ssw_config(2, 0.1, 0)
        plane = torch.empty((h, w), dtype=torch.float32, device='cuda')
        idx = torch.empty((n,), dtype=torch.int32, device='cuda')
        stages = {
            'forward': lambda: wm.lib.ssw_stage_forward_rgb8_dev(ctx.handle, frames.data_ptr(), w, h, 1, plane.data_ptr()),
            'topk': lambda: wm.lib.ssw_stage_topk_dev(ctx.handle, plane.data_ptr(), w, h, 1, 0, n, idx.data_ptr()),
            'inverse': lambda: wm.lib.ssw_stage_inverse_rgb8_dev(ctx.handle, plane.data_ptr(), frames.data_ptr(), w, h, 1, out.data_ptr()),
            'embed(all)': lambda: wm.lib.ssw_embed_batch_rgb8_dev(ctx.handle, frames.data_ptr(), w, h, 1, ctypes.byref(cfg), mk.data_ptr(), n, out.data_ptr()),
        }
        for name, fn in stages.items():
            for _ in range(3):
                wm._lib.check(fn())
            ctx.synchronize()
            t0 = time.time()
            for _ in range(20):
                fn()
            ctx.synchronize()
            dt = (time.time() - t0) / 20
            print('    %-10s %.1f us -> %.0f Mpix/s' % (name, dt * 1e6, w * h / dt / 1e6))
        f0 = frames[0].cpu().numpy()
        t0 = time.time()
        ref_img, ref_idx, _ = so.embed(f0, [mk[0].cpu().numpy()])
        print('    numpy oracle embed %.2fs' % (time.time() - t0))
        d = np.abs(out[0].cpu().numpy().astype(int) - ref_img.astype(int))
        print('    4K vs oracle: %d differ (max %d)' % ((d > 0).sum(), d.max()))
        assert d.max() <= 1

print('FAILED: %s' % FAILS if FAILS else 'ALL OK')
sys.exit(1 if FAILS else 0)
