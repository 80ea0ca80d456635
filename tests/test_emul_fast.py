"""The fast-path (compile-time planned) DCT kernel PHASES of csrc/dct_fast.cuh executed on the CPU
(tests/emul/fast_emul.cpp) against the oracle.  Exercises the exact index arithmetic, radix plans
and colour math of the device code in the GPU-less build container; test infrastructure only."""
import ctypes

import numpy as np
import pytest

from conftest import ptr

FAST = [3840, 2160, 1920, 1080, 640, 1280, 720, 2560, 1440, 7680, 4320, 1024, 2048, 4096, 8192, 16384]
f32 = ctypes.c_float


def test_u8_unit_is_exact(emul):
    assert emul.emul_fast_u8_unit_mismatches() == 0


def test_unit_to_u8_fast_is_exact(emul):
    rng = np.random.default_rng(0)
    k = np.arange(0, 256, dtype=np.float64)
    halves = ((k + 0.5) / 255.0).astype(np.float32)
    edge = np.concatenate([np.nextafter(halves, np.float32(0)), halves, np.nextafter(halves, np.float32(2))])
    tiny = np.array([0.5 / 255, np.nextafter(np.float32(0.5 / 255), np.float32(0)), 0.49999997 / 255, 0.0, -0.0, 1.0, -1.0, 2.0,
                     np.nan, np.inf, -np.inf, 1e-30, 0.0019607842], np.float32)
    v = np.concatenate([rng.random(200000).astype(np.float32) * 1.2 - 0.1, edge.astype(np.float32), tiny]).astype(np.float32)
    assert emul.emul_fast_unit_to_u8_mismatches(ptr(v), len(v)) == 0


def test_byte_unpack_is_exact(emul):
    """u8 -> float through PRMT + FADD (0x4B0000bb - 2^23) instead of I2F: every byte value, every byte lane"""
    emul.emul_fast_byte_unpack_mismatches.restype = ctypes.c_int
    assert emul.emul_fast_byte_unpack_mismatches() == 0


def test_pack_u8x4_is_exact_for_every_float(emul):
    """round(clamp(v)*255) through two round-toward-zero adds and a PRMT gather == the reference rounding
    (image::into_rgb8 after src/yiq.rs:139-147) for ALL 2^32 f32 bit patterns (NaNs, infinities, denormals)"""
    from concurrent.futures import ThreadPoolExecutor
    f = emul.emul_fast_pack_u8_mismatches
    f.restype = ctypes.c_longlong
    f.argtypes = [ctypes.c_uint, ctypes.c_uint]
    parts = 16
    with ThreadPoolExecutor(max_workers=8) as ex:
        bad = sum(ex.map(lambda p: f(p, parts), range(parts)))
    assert bad == 0


def test_plan_table(emul):
    for n in FAST:
        assert emul.emul_fast_has_plan(n) == 1
    for n in (444, 37, 1000, 32768):
        assert emul.emul_fast_has_plan(n) == 0


def dct1d_rows(so, a, kind):
    """reference 1-D pass along x for every row (scipy scaling, src/dct2d.rs:107-111)"""
    import scipy.fft
    a64 = a.astype(np.float64)
    if kind == 'fwd':
        return scipy.fft.dct(a64, type=2, axis=1)
    return 0.25 * scipy.fft.dct(a64, type=3, axis=1)


@pytest.mark.parametrize('n', FAST)
@pytest.mark.parametrize('h', [5, 8])
def test_row_passes_plane(emul, so, n, h):
    rng = np.random.default_rng(n + h)
    a = rng.random((2, h, n)).astype(np.float32)   # batch of 2
    out = np.zeros_like(a)
    assert emul.emul_fast_row_fwd(2, ptr(a), n, h, 2, ptr(out), f32(1.0), f32(1.0)) == 0
    ref = dct1d_rows(so, a.reshape(2 * h, n), 'fwd').reshape(a.shape)
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()
    back = np.zeros_like(a)
    c = ref.astype(np.float32)
    assert emul.emul_fast_row_inv(2, 2, ptr(c), None, n, h, 2, ptr(back), f32(2.0 / n)) == 0
    assert np.abs(back - a).max() <= 3e-6


@pytest.mark.parametrize('n', FAST)
@pytest.mark.parametrize('w', [8, 20])
def test_col_passes(emul, so, n, w):
    if emul.emul_fast_col_pairs(n) == 0:
        pytest.skip('column tile of %d-point lines does not fit shared memory: generic column kernel' % n)
    rng = np.random.default_rng(n + w)
    a = rng.random((2, n, w)).astype(np.float32)
    f = a.copy()
    assert emul.emul_fast_col(0, w, n, 2, ptr(f), f32(1.0), f32(1.0)) == 0
    import scipy.fft
    ref = scipy.fft.dct(a.astype(np.float64), type=2, axis=1)
    assert np.abs(f - ref).max() <= 2e-6 * np.abs(ref).max()
    b = ref.astype(np.float32)
    assert emul.emul_fast_col(1, w, n, 2, ptr(b), f32(2.0 / n), f32(1.0)) == 0
    assert np.abs(b - a).max() <= 3e-6


def test_fused_rgb8_frame_1080_rows(emul, so):
    """RGB8 -> luma -> row DCT and back through the recolouring store, on a 1920-wide strip"""
    w, h = 1920, 6
    rgb = so.synth_frame(w, h, seed=5)
    plane = np.zeros((h, w), np.float32)
    assert emul.emul_fast_row_fwd(0, ptr(rgb), w, h, 1, ptr(plane), f32(1.0), f32(1.0)) == 0
    y, _, _ = so.rgb32f_to_yiq(so.rgb8_to_rgb32f(rgb))
    ref = dct1d_rows(so, y, 'fwd')
    assert np.abs(plane - ref).max() <= 2e-6 * np.abs(ref).max()
    out = np.zeros_like(rgb)
    assert emul.emul_fast_row_inv(0, 0, ptr(plane.copy()), ptr(rgb), w, h, 1, ptr(out), f32(2.0 / w)) == 0
    assert np.abs(out.astype(int) - rgb.astype(int)).max() <= 1
    assert (out != rgb).mean() < 0.01


def test_full_frame_640x1080_matches_oracle(emul, so):
    """both passes of a frame whose width and height are planned lengths, odd tile counts included"""
    w, h = 640, 1080
    rgb = so.synth_frame(w, h, seed=9)
    plane = np.zeros((h, w), np.float32)
    assert emul.emul_fast_row_fwd(0, ptr(rgb), w, h, 1, ptr(plane), f32(1.0), f32(1.0)) == 0
    assert emul.emul_fast_col(0, w, h, 1, ptr(plane), f32(1.0), f32(1.0)) == 0
    ref, _, _ = so.forward(rgb)
    tol = 1e-5 * np.abs(ref) + 1e-7 * np.abs(ref).max()
    assert (np.abs(plane - ref) <= tol).all()
    out = np.zeros_like(rgb)
    assert emul.emul_fast_col(1, w, h, 1, ptr(plane), f32(1.0), f32(1.0)) == 0
    assert emul.emul_fast_row_inv(0, 0, ptr(plane), ptr(rgb), w, h, 1, ptr(out), f32(4.0 / (w * h))) == 0
    assert np.abs(out.astype(int) - rgb.astype(int)).max() <= 1


def test_rgb32f_rows_match_rgb8_rows(emul, so):
    """RGB32F pixels (u8/255 done by the caller, like into_rgb32f) give the same coefficients and the
    RGB32F destination is the clamped float image whose into_rgb8 equals the fused RGB8 store"""
    w, h = 1080, 4
    rgb = so.synth_frame(w, h, seed=6)
    rgbf = so.rgb8_to_rgb32f(rgb)
    p8 = np.zeros((h, w), np.float32); p32 = np.zeros((h, w), np.float32)
    assert emul.emul_fast_row_fwd(0, ptr(rgb), w, h, 1, ptr(p8), f32(1.0), f32(1.0)) == 0
    assert emul.emul_fast_row_fwd(1, ptr(rgbf), w, h, 1, ptr(p32), f32(1.0), f32(1.0)) == 0
    assert (p8 == p32).all()
    p8[:, 1:8] *= 1.6  # perturb so that the output differs from the input
    o8 = np.zeros_like(rgb); o32 = np.zeros_like(rgbf); o8b = np.zeros_like(rgb); o32b = np.zeros_like(rgbf)
    s = f32(2.0 / w)
    assert emul.emul_fast_row_inv(0, 0, ptr(p8.copy()), ptr(rgb), w, h, 1, ptr(o8), s) == 0
    assert emul.emul_fast_row_inv(1, 0, ptr(p8.copy()), ptr(rgb), w, h, 1, ptr(o32), s) == 0
    assert emul.emul_fast_row_inv(0, 1, ptr(p8.copy()), ptr(rgbf), w, h, 1, ptr(o8b), s) == 0
    assert emul.emul_fast_row_inv(1, 1, ptr(p8.copy()), ptr(rgbf), w, h, 1, ptr(o32b), s) == 0
    assert (so.rgb32f_to_rgb8(o32) == o8).all() and (o8b == o8).all() and (o32b == o32).all()
    assert o32.min() >= 0.0 and o32.max() <= 1.0 and (o8 != rgb).any()


@pytest.mark.parametrize('n', [1024, 4096, 32768])
def test_single_line_kernels(emul, so, n):
    """one real line through an n/2-point complex FFT (the form that carries 32768-point lines)"""
    lines = 3 if n < 32768 else 2
    rng = np.random.default_rng(n)
    a = rng.random((lines, n)).astype(np.float32)
    out = np.zeros_like(a)
    assert emul.emul_line1_fwd(2, ptr(a), n, lines, ptr(out), f32(1.0), f32(1.0)) == 0
    ref = dct1d_rows(so, a, 'fwd')
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()
    back = np.zeros_like(a)
    c = ref.astype(np.float32)
    assert emul.emul_line1_inv(2, ptr(c), None, n, lines, ptr(back), f32(2.0 / n)) == 0
    assert np.abs(back - a).max() <= 4e-6


def test_single_line_rgb8(emul, so):
    n, lines = 1024, 4
    rgb = so.synth_frame(n, lines, seed=8)
    plane = np.zeros((lines, n), np.float32)
    assert emul.emul_line1_fwd(0, ptr(rgb), n, lines, ptr(plane), f32(1.0), f32(1.0)) == 0
    y, _, _ = so.rgb32f_to_yiq(so.rgb8_to_rgb32f(rgb))
    ref = dct1d_rows(so, y, 'fwd')
    assert np.abs(plane - ref).max() <= 2e-6 * np.abs(ref).max()
    out = np.zeros_like(rgb)
    assert emul.emul_line1_inv(0, ptr(plane.copy()), ptr(rgb), n, lines, ptr(out), f32(2.0 / n)) == 0
    assert np.abs(out.astype(int) - rgb.astype(int)).max() <= 1 and (out != rgb).mean() < 0.01


@pytest.mark.parametrize('line1,n,seg,chunks,ranks', [(1, 1024, 128, 2, 4), (1, 4096, 512, 4, 2), (1, 1024, 1024, 1, 1), (0, 1920, 128, 1, 15),
                                                     (0, 640, 32, 4, 5)])
def test_segmented_source_lines(emul, so, line1, n, seg, chunks, ranks):
    """lines assembled from all-to-all blocks ([chunks][ranks][lines][seg]) are read in place"""
    lines = 5
    rng = np.random.default_rng(n + seg)
    a = rng.random((lines, n)).astype(np.float32)
    # sample m of line l: s = m // seg, g = s // chunks, c = s % chunks
    blocks = np.zeros((chunks, ranks, lines, seg), np.float32)
    for s_ in range(n // seg):
        g, c = divmod(s_, chunks)
        blocks[c, g] = a[:, s_ * seg:(s_ + 1) * seg]
    out = np.zeros_like(a)
    assert emul.emul_fwd_segmented(line1, ptr(blocks), n, lines, seg, chunks, ranks, ptr(out)) == 0
    ref = dct1d_rows(so, a, 'fwd')
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()
    # inverse: the same block layout holds coefficient lines; result = 0.25*dct3 * scale
    back = np.zeros_like(a)
    assert emul.emul_inv_segmented(line1, ptr(blocks), n, lines, seg, chunks, ranks, ptr(back), f32(2.0 / n)) == 0
    refi = dct1d_rows(so, a, 'inv') * np.float32(2.0 / n)
    assert np.abs(back - refi).max() <= 3e-6 * max(np.abs(refi).max(), 1e-30)



@pytest.mark.parametrize('n,h,per', [(3840, 7, 2), (1920, 10, 3), (640, 9, 8), (2160, 4, 1)])
def test_prefetching_row_kernel_is_bit_identical(emul, so, n, h, per):
    """cp.async-staged rows (several tiles per CTA) give exactly the coefficients of the direct loader"""
    rgb = so.synth_frame(n, h, seed=n + per, img=1)
    rgb2 = np.concatenate([rgb, so.synth_frame(n, h, seed=n, img=2)])   # batch of 2 images
    a, b = np.zeros((2 * h, n), np.float32), np.zeros((2 * h, n), np.float32)
    assert emul.emul_fast_row_fwd(0, ptr(rgb2), n, h, 2, ptr(a), f32(1.0), f32(1.0)) == 0
    assert emul.emul_fast_row_fwd_pf(ptr(rgb2), n, h, 2, ptr(b), per) == 0
    assert (a == b).all()


@pytest.mark.parametrize('n,variant', [(2160, 1), (2160, 2), (2160, 3), (1080, 1), (1920, 1), (2048, 1), (720, 1), (1440, 1)])
@pytest.mark.parametrize('inverse', [0, 1])
def test_col_pipe_matches_col_pass(emul, n, variant, inverse):
    """persistent TMA column pipelines (csrc/dct_pipe.cuh), phases run on the CPU with memcpy standing in for the tensor-map
    copies (parity-split sample side, natural coefficient side, zero fill / clipping of columns past the frame): the
    planes must be bit-identical to the one-CTA-per-tile column kernels, whose arithmetic they share"""
    w = 20   # 20 columns: tiles of 8 (4) columns, the last one partly outside the frame
    rng = np.random.default_rng(n + variant)
    a = (rng.random((2, n, w)).astype(np.float32) - 0.5) * 3.0
    ref = a.copy()
    got = a.copy()
    assert emul.emul_fast_col(inverse, w, n, 2, ptr(ref), f32(1.0 if inverse else 0.7), f32(1.0 if inverse else 1.3)) == 0
    assert emul.emul_col_pipe(variant, inverse, w, n, 2, ptr(got), f32(1.0 if inverse else 0.7), f32(1.0 if inverse else 1.3)) == 0
    assert np.abs(ref).max() > 1.0
    assert np.array_equal(ref, got)


@pytest.mark.parametrize('variant', [4, 5])   # 2 teams: one round per half tile; 4 teams: teams 0 and 1 take the half tile
@pytest.mark.parametrize('inverse', [0, 1])
def test_col_pipe_half_tiles_match_col_pass(emul, inverse, variant):
    """split schedule of the column pipelines (PipeArgs::half_tiles): 3 tiles on a 2-CTA grid = one round of whole tiles + the
    third tile as two half tiles (4 columns, one round over 2 line pairs, buffer rows of 4 floats) -- bit-identical planes"""
    w, n = 24, 2160
    rng = np.random.default_rng(77 + inverse)
    a = (rng.random((1, n, w)).astype(np.float32) - 0.5) * 3.0
    ref, got = a.copy(), a.copy()
    assert emul.emul_fast_col(inverse, w, n, 1, ptr(ref), f32(1.0 if inverse else 0.7), f32(1.0 if inverse else 1.3)) == 0
    assert emul.emul_col_pipe(variant, inverse, w, n, 1, ptr(got), f32(1.0 if inverse else 0.7), f32(1.0 if inverse else 1.3)) == 0
    assert np.abs(ref).max() > 1.0
    assert np.array_equal(ref, got)


@pytest.mark.parametrize('n,h', [(3840, 4), (1920, 8), (1080, 4), (2160, 2), (640, 8), (1280, 4), (720, 4), (2560, 2), (1440, 4)])
def test_row_pipe_matches_row_kernels(emul, so, n, h):
    """persistent bulk-copy row pipelines (csrc/dct_pipe.cuh RowPipe), phases run on the CPU with memcpy standing in for the
    bulk copies: coefficient planes and RGB8 output must be bit-identical to the one-CTA-per-tile row kernels"""
    rgb = np.concatenate([so.synth_frame(n, h, seed=n, img=1), so.synth_frame(n, h, seed=n + 1, img=2)])   # batch of 2 frames
    a, b = np.zeros((2 * h, n), np.float32), np.zeros((2 * h, n), np.float32)
    assert emul.emul_fast_row_fwd(0, ptr(rgb), n, h, 2, ptr(a), f32(0.7), f32(1.3)) == 0
    assert emul.emul_row_pipe(0, ptr(rgb), n, h, 2, ptr(b), None, f32(0.7), f32(1.3)) == 0
    assert np.abs(a).max() > 1.0 and np.array_equal(a, b)
    rng = np.random.default_rng(n)
    coef = ((rng.random((2 * h, n)).astype(np.float32) - 0.5) * 2.0)
    coef[:, 0] += 80.0
    o1, o2 = np.zeros_like(rgb), np.zeros_like(rgb)
    c1, c2, c3 = coef.copy(), coef.copy(), coef.copy()   # (named: a temporary would be freed before the call reads it)
    assert emul.emul_fast_row_inv(0, 0, ptr(c1), ptr(rgb), n, h, 2, ptr(o1), f32(2.0 / n)) == 0
    assert emul.emul_row_pipe(1, ptr(rgb), n, h, 2, ptr(c2), ptr(o2), f32(2.0 / n), f32(2.0 / n)) == 0
    assert len(np.unique(o1)) > 50 and np.array_equal(o1, o2)
    # in-place shape of the inverse pipeline (coefficient rows land in the FFT buffers, two-phase pre pass)
    o3 = np.zeros_like(rgb)
    rc = emul.emul_row_pipe(2, ptr(rgb), n, h, 2, ptr(c3), ptr(o3), f32(2.0 / n), f32(2.0 / n))
    assert rc in (0, -2)
    if n in (3840, 1920):
        assert rc == 0   # the shapes the in-place pipeline is built for
    if rc == 0:
        assert np.array_equal(o1, o3)
    # rows after a partial inverse column pass (RowPipeArgs::col_cut_img): the coefficients of the columns >= kcut come
    # without the gain of the skipped passes and get it as they are read.  With a power-of-two gain the scaling is exact:
    # plane / gain beyond kcut + the cut == the whole plane, byte for byte, in both shapes of the inverse pipeline
    kcut, gain = 24, 512.0
    scaled = coef.copy()
    scaled[:, kcut:] *= np.float32(1.0 / gain)
    for variant in (1, 2):
        o4, c4 = np.zeros_like(rgb), scaled.copy()
        rc = emul.emul_row_pipe_cut(variant, ptr(rgb), n, h, 2, ptr(c4), ptr(o4), f32(2.0 / n), f32(2.0 / n), kcut, f32(gain))
        assert rc in (0, -2) and (variant == 2 or rc == 0)
        if rc == 0:
            assert np.array_equal(o1, o4)
