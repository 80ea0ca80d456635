"""The oracle (numpy restatement AND the C restatement) against the reference's own unit-test
known answers.  Vectors are restated from the reference's tests (paths relative to the reference):
  src/dct2d.rs:229-524   scipy.fftpack known answers for DCT2 / DCT3 / DCT2Orthogonal
  src/yiq.rs:204-241     RGB <-> YIQ
  src/algorithm.rs:723-863  ordering, insert/extract options, single / multi mark embedding
"""
import ctypes

import numpy as np
import pytest
import scipy.fftpack

from conftest import ptr

TOL = 1e-4  # approx_equal tolerance used by the reference's DCT tests


def c_dct(coracle, kind, data):
    h, w = data.shape
    buf = np.ascontiguousarray(data, dtype=np.float32).copy()
    coracle.oracle_dct2_2d(ctypes.c_int(kind), ctypes.c_int(w), ctypes.c_int(h), ptr(buf))
    return buf


def both(so, coracle, data, kind):
    """the transform by both oracles: numpy f64, numpy f32, C f32"""
    k = {0: so.DCT2, 1: so.DCT2_ORTHO, 2: so.DCT3}[kind]
    a = np.asarray(data, dtype=np.float32)
    return [so.dct2_2d(a, k, np.float64), so.dct2_2d(a, k, np.float32), c_dct(coracle, kind, a)]


# ---------------------------------------------------------------------------- src/dct2d.rs:229-265
def test_simple_dct_against_scipy(so, coracle):
    x = np.array([[1.0, 0.0, 0.0]], np.float32)
    expected = np.array([2.0, 1.73205081, 1.0])  # scipy.fftpack.dct; rustdct alone gives half of it
    for r in both(so, coracle, x.T, 0)[:2] + [c_dct(coracle, 0, x.T)]:
        # a 1-wide, 3-high frame: the row pass (length 1) is x2, the column pass is the 1-D DCT
        assert np.allclose(np.ravel(r) / 2.0, expected, atol=TOL)


# ---------------------------------------------------------------------------- src/dct2d.rs:268-323
@pytest.mark.parametrize('inp,res', [
    ([1, 0, 0, 1, 0, 0, 0, 0, 1], [12, 3.46410162, 6.0, 0.0, 6.0, 0.0, 0.0, -3.46410162, 0.0]),
    ([1, 0, 0, 2, 0, 0, 0, 0, 3], [24, 0.0, 12.0, -6.92820323, 12.0, -3.46410162, 0.0, -10.3923048, 0.0]),
])
def test_2d_dct_against_scipy_3x3(so, coracle, inp, res):
    x = np.array(inp, np.float32).reshape(3, 3)
    for r in both(so, coracle, x, 0):
        assert np.allclose(np.ravel(r), res, atol=TOL)
        for back in both(so, coracle, r, 2):
            assert np.allclose(back, x, atol=TOL)


# ---------------------------------------------------------------------------- src/dct2d.rs:326-428
def test_2d_dct_against_scipy_larger(so, coracle):
    np.random.seed(0)
    x = np.random.rand(5, 4)  # 4 wide, 5 high -- the recipe in the reference's comment
    assert abs(x[0, 0] - 0.5488135039273248) < 1e-15 and abs(x[4, 3] - 0.8700121482468192) < 1e-15
    dct = scipy.fftpack.dct
    expected = dct(dct(x).T).T
    assert abs(expected[0, 0] - 46.524385961807795) < 1e-9 and abs(expected[4, 3] - 4.745483123369016) < 1e-9
    ortho = scipy.fftpack.dct(scipy.fftpack.dct(x, norm='ortho').T, norm='ortho').T
    assert abs(ortho[0, 0] - 2.600792240550979) < 1e-9
    for r in both(so, coracle, x, 0):
        assert np.allclose(r, expected, atol=TOL)
        for back in both(so, coracle, r, 2):
            assert np.allclose(back, x, atol=TOL)
    for r in both(so, coracle, x, 1):
        assert np.allclose(r, ortho, atol=TOL)


# ---------------------------------------------------------------------------- src/dct2d.rs:431-524
def test_ortho_dct_against_scipy(so, coracle):
    x = np.array([1, 0, 0, 2, 0, 0, 0, 0, 3], np.float32).reshape(3, 3)
    res = [2.0, 0.0, 1.4142135623730954, -0.816496580927726, 2.0, -0.5773502691896258, 0.0, -1.7320508075688774, 0.0]
    for r in both(so, coracle, x, 1):
        assert np.allclose(np.ravel(r), res, atol=TOL)
    x = np.array([1, 2, 3, 4, 2, 3, 5, 1, 0, 0, 3, 3], np.float32).reshape(3, 4)  # 4 wide, 3 high
    res = [7.794228634059947, -2.8232403410227764, -1.4433756729740645, 1.4818841531942584,
           1.414213562373095, 0.3826834323650898, 0.0, -0.9238795325112866,
           -1.224744871391589, -2.1336083871767086, 2.0412414523193156, -0.8837695307615787]
    for r in both(so, coracle, x, 1):
        assert np.allclose(np.ravel(r), res, atol=TOL)


def test_c_oracle_matches_numpy_oracle_on_awkward_sizes(so, coracle):
    rng = np.random.default_rng(3)
    for h, w in [(1, 1), (1, 7), (7, 1), (37, 12), (44, 64), (30, 45), (128, 96)]:
        x = rng.random((h, w)).astype(np.float32)
        ref = so.dct2_2d(x, so.DCT2)
        got = c_dct(coracle, 0, x)
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
        back = c_dct(coracle, 2, got)
        assert np.abs(back - x).max() < 1e-5


# ---------------------------------------------------------------------------- src/yiq.rs:204-241
@pytest.mark.parametrize('rgb,yiq', [
    ([1.0, 0.0, 0.0], [0.3, 0.6, 0.21]),
    ([0.0, 1.0, 0.0], [0.59, -0.28, -0.52]),
    ([0.0, 0.0, 1.0], [0.11, -0.32, 0.31]),
    ([0.5, 0.5, 1.0], [0.555, -0.16, 0.155]),
])
def test_yiq_to_rgb(so, rgb, yiq):
    y, i, q = so.rgb32f_to_yiq(np.array([[rgb]], np.float32))
    assert np.allclose([y[0, 0], i[0, 0], q[0, 0]], yiq, atol=1e-4)
    back = so.yiq_to_rgb32f(*(np.array([[v]], np.float32) for v in yiq))
    assert np.allclose(back[0, 0], rgb, atol=1e-4)


def test_yiq_to_rgb_image(so):
    img = np.zeros((5, 5, 3), np.float32)
    img[0, 0] = [0.1, 0.2, 0.3]; img[1, 0] = [0.11, 0, 0]; img[0, 1] = [0.21, 0, 0]
    img[4, 4] = [0.5, 0.3, 0.8]; img[0, 3] = [1.0, 0, 0]
    back = so.yiq_to_rgb32f(*so.rgb32f_to_yiq(img))
    assert back.shape == img.shape and np.allclose(back, img, atol=1e-3)


def test_u8_conversions(so):
    v = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, axis=2)
    f = so.rgb8_to_rgb32f(v)
    assert f.dtype == np.float32 and f[0, 255, 0] == 1.0 and f[0, 51, 0] == np.float32(51) / np.float32(255)
    assert (so.rgb32f_to_rgb8(f) == v).all()
    edge = np.array([[[-0.5, 0.5 / 255.0, 1.5]]], np.float32)   # clamp, half rounds away from zero
    assert so.rgb32f_to_rgb8(edge).ravel().tolist() == [0, 1, 255]


# ---------------------------------------------------------------------------- src/algorithm.rs:723-863
COEF = np.array([-3, 5.0, -8.0, 7.0, 1.0, 2.0], np.float32)


def test_indices(so, coracle):
    assert so.obtain_indices(COEF).tolist() == [2, 3, 1, 5, 4]
    out = np.zeros(5, np.uint64)
    coracle.oracle_obtain_indices(ptr(COEF), 6, 1, 0, ptr(out))
    assert out.tolist() == [2, 3, 1, 5, 4]


def test_indices_ties_and_total_cmp(so, coracle):
    c = np.array([9, 2, -2, 0.0, -0.0, 2, np.nan, np.inf, -np.inf, 1e-30], np.float32)
    # energies: 4 4 0 0 4 nan inf inf ~0(underflow) -> NaN first, infs by index, ties by index
    ref = [6, 7, 8, 1, 2, 5, 3, 4, 9]
    assert so.obtain_indices(c).tolist() == ref
    out = np.zeros(9, np.uint64)
    coracle.oracle_obtain_indices(ptr(c), 10, 1, 0, ptr(out))
    assert out.tolist() == ref


def test_orderings_agree_between_oracles(so, coracle):
    rng = np.random.default_rng(5)
    w, h = 13, 9
    c = rng.standard_normal(w * h).astype(np.float32)
    c[5] = c[17]  # a tie
    for ordering in (0, 1, 2):
        out = np.zeros(w * h - 1, np.uint64)
        coracle.oracle_obtain_indices(ptr(c), w, h, ordering, ptr(out))
        assert (out == so.obtain_indices(c, ordering, w, h)).all()


def test_insert_extract_functions(so):
    m = np.array([1.0, -0.5, 1.0, 0.5, 0.5, 0.1], np.float32)
    idx = np.arange(6)
    for method in (1, 2, 3):
        emb = so.embed_watermark(COEF, idx, [m], method, 0.1)
        ext = so.extract_watermark(np.append(COEF, 0), idx, np.append(emb, 0), 6, method, 0.1)
        assert np.allclose(ext, m, atol=1e-3)


def test_embedder_single(so):
    idx = so.obtain_indices(COEF)
    s = np.float32(0.1)
    one = np.float32(1.0)
    out = so.embed_watermark(COEF, idx, [np.array([1.0, -0.5, 1.0], np.float32)])
    expected = np.array([-3, np.float32(5) * (one + one * s), np.float32(-8) * (one + one * s),
                         np.float32(7) * (one - np.float32(0.5) * s), 1, 2], np.float32)
    assert (out == expected).all()  # assert_eq!, exact
    ext = so.extract_watermark(COEF, idx, out, 3)
    assert np.abs(ext - [1.0, -0.5, 1.0]).max() < 1e-6


def test_embedder_single_and_zero(so):
    idx = so.obtain_indices(COEF)
    a = so.embed_watermark(COEF, idx, [np.array([1.0, -0.5, 1.0], np.float32)])
    b = so.embed_watermark(COEF, idx, [np.array([1.0, -0.5, 1.0], np.float32), np.zeros(3, np.float32)])
    assert (a == b).all()


def test_embedder_multiple(so, coracle):
    idx = so.obtain_indices(COEF)
    m1 = np.array([1.0, -0.5, 1.0], np.float32)
    m2 = np.array([0.5, -0.5, -1.0], np.float32)
    out = so.embed_watermark(COEF, idx, [m1, m2])
    f = np.float32
    s = f(0.1)
    v2 = f(f(-8) + f(f(-8) * f(1 + 1 * s) - f(-8))) + f(f(-8) * f(1 + f(0.5) * s) - f(-8))
    v3 = f(f(7) + f(f(7) * f(1 + f(-0.5) * s) - f(7))) + f(f(7) * f(1 + f(-0.5) * s) - f(7))
    v1 = f(f(5) + f(f(5) * f(1 + 1 * s) - f(5))) + f(f(5) * f(1 + -1 * s) - f(5))
    assert (out == np.array([-3, v1, v2, v3, 1, 2], np.float32)).all()
    # the C restatement agrees bit for bit
    c = COEF.copy()
    idx64 = idx.astype(np.uint64)
    marks = (ctypes.c_void_p * 2)(m1.ctypes.data, m2.ctypes.data)
    lens = (ctypes.c_size_t * 2)(3, 3)
    coracle.oracle_embed_watermark(ptr(c), ptr(idx64), ctypes.c_size_t(5), marks, lens, ctypes.c_size_t(2), 2,
                                   ctypes.c_float(0.1), ctypes.c_size_t(6))
    assert (c == out).all()


def test_similarity(so, coracle):
    rng = np.random.default_rng(1)
    e = rng.standard_normal(1000).astype(np.float32)
    m = rng.standard_normal(1000).astype(np.float32)
    s = so.similarity(e, m)
    coracle.oracle_similarity.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    assert np.float32(coracle.oracle_similarity(ptr(e), ptr(m), 1000)) == s
    assert abs(float(s) - float(e @ m / np.sqrt(e @ e))) < 1e-3
    assert abs(float(so.similarity(m, m)) - np.sqrt(float(m @ m))) < 1e-2
