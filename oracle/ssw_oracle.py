"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

CPU restatement (numpy / scipy) of the embed / extract / similarity hot path of
iwanders/spread_spectrum_watermarking.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module, and only as the
checker.  Every function cites the reference lines it follows (paths relative to /root/reference).

The 1-D DCT arithmetic of the reference lives in the un-vendored crates rustdct 0.7.0 ->
rustfft 6.0.1 (Cargo.lock:584-603).  Their *definition and scaling* are pinned by the reference's
own scipy known-answer tests (src/dct2d.rs:229-524, re-expressed in tests/test_oracle_kat.py), so
the oracle evaluates the same definition with scipy.fft (pocketfft), in float64 by default
("the exact answer any correct f32 implementation must be close to") or in float32.

PARITY IS PINNED: tests/test_oracle_golden.py checks this module against the reference's golden
PNG (tests/single_simple.rs:36-43), its extraction/similarity thresholds (:61,70,79,90) and the
attack_crop similarity (tests/attack_crop.rs:93-94).
"""
import numpy as np
import scipy.fft

F32 = np.float32

# src/yiq.rs:157-159
RGB_TO_YIQ = np.array([[0.30, 0.59, 0.11],
                       [0.60, -0.28, -0.32],
                       [0.21, -0.52, 0.31]], dtype=F32)
# src/yiq.rs:163-165
YIQ_TO_RGB = np.array([[1.0, 0.948262, 0.624013],
                       [1.0, -0.276066, -0.639810],
                       [1.0, -1.105450, 1.729860]], dtype=F32)

ORDER_ENERGY, ORDER_ENERGY_ORTHO, ORDER_LEGACY = 0, 1, 2


# ------------------------------------------------------------------------------------------------
# colour space
# ------------------------------------------------------------------------------------------------
def rgb8_to_rgb32f(rgb8):
    """image 0.24.3 `into_rgb32f` (call site src/algorithm.rs:308,476): u8 as f32 / 255.0."""
    return rgb8.astype(F32) / F32(255.0)


def rgb32f_to_yiq(rgb):
    """src/yiq.rs:131-136,168-170,177-186: (m0*r + m1*g) + m2*b in f32, separate mul/add.
    Returns three separate planes (src/yiq.rs:58-62)."""
    rgb = rgb.astype(F32)
    r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    m = RGB_TO_YIQ
    planes = []
    for k in range(3):
        planes.append(((m[k, 0] * r).astype(F32) + (m[k, 1] * g).astype(F32)).astype(F32)
                      + (m[k, 2] * b).astype(F32))
    return [p.astype(F32) for p in planes]


def yiq_to_rgb32f(y, i, q):
    """src/yiq.rs:139-147,173-175,187-197: matrix product then clamp to [0,1]."""
    m = YIQ_TO_RGB
    out = np.empty(y.shape + (3,), dtype=F32)
    y = y.astype(F32); i = i.astype(F32); q = q.astype(F32)
    for k in range(3):
        v = ((m[k, 0] * y).astype(F32) + (m[k, 1] * i).astype(F32)).astype(F32) + (m[k, 2] * q).astype(F32)
        out[..., k] = np.clip(v.astype(F32), F32(0.0), F32(1.0))
    return out


def rgb32f_to_rgb8(rgb):
    """image 0.24.3 `into_rgb8` (call site tests/single_simple.rs:28):
    round(clamp(v,0,1)*255), f32, half away from zero."""
    v = (np.clip(rgb.astype(F32), F32(0), F32(1)) * F32(255.0)).astype(F32)
    return np.floor(v.astype(np.float64) + 0.5).astype(np.uint8)  # v >= 0: half away == floor(v+.5)


# ------------------------------------------------------------------------------------------------
# 2-D DCT driver -- src/dct2d.rs:83-219
# ------------------------------------------------------------------------------------------------
DCT2, DCT2_ORTHO, DCT3 = 'dct2', 'dct2_ortho', 'dct3'


def _dct1d(a, axis, kind, dtype):
    a = a.astype(dtype)
    n = a.shape[axis]
    if kind == DCT2:
        # rustdct dct2 (no factor 2) * T::two()  (src/dct2d.rs:107-108,166,202) == scipy norm=None
        return scipy.fft.dct(a, type=2, axis=axis)
    if kind == DCT2_ORTHO:
        # src/dct2d.rs:153-162,189-198: additionally sqrt(1/4N) for k=0, sqrt(1/2N) otherwise,
        # scale factors evaluated in f32.
        r = scipy.fft.dct(a, type=2, axis=axis)
        s0 = np.sqrt(F32(1.0) / (F32(4.0) * F32(n)))
        sn = np.sqrt(F32(1.0) / (F32(2.0) * F32(n)))
        sc = np.full(n, sn, dtype=dtype)
        sc[0] = s0
        shape = [1] * a.ndim
        shape[axis] = n
        return r * sc.reshape(shape)
    if kind == DCT3:
        # rustdct dct3 = x0/2 + sum_{n>=1} x_n cos(pi n (k+1/2)/N)  = scipy dct3(norm=None)/2,
        # times T::half() (src/dct2d.rs:109)
        return scipy.fft.dct(a, type=3, axis=axis) * dtype(0.25)
    raise ValueError(kind)


def dct2_2d(data, kind=DCT2, dtype=np.float64):
    """src/dct2d.rs:83-219.  `data` is [H][W] row-major.  Longest dimension first (:93-97); the
    DCT3 result is finally scaled by fl32(4)/fl32(W*H) (:213-217).  Returns `dtype`."""
    h, w = data.shape
    order = (1, 0) if w >= h else (0, 1)  # axis 1 == rows of length w
    out = data.astype(dtype)
    for ax in order:
        out = _dct1d(out, ax, kind, dtype).astype(dtype)
    if kind == DCT3:
        out = out * dtype(F32(4.0) / F32(w * h))
    return out.astype(dtype)


# ------------------------------------------------------------------------------------------------
# ordering -- src/algorithm.rs:200-280
# ------------------------------------------------------------------------------------------------
def _total_cmp_key(x):
    """f32::total_cmp as an unsigned sortable key."""
    b = np.ascontiguousarray(x, dtype=F32).view(np.uint32)
    return np.where(b >> 31, ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def _ortho_scaled(coeff, w, h):
    """src/algorithm.rs:240-266 (f32; scaling = 1.0 * s_row * s_col, then * value)."""
    s_k0_w = np.sqrt(F32(1.0) / (F32(4.0) * F32(w)))
    s_k0_h = np.sqrt(F32(1.0) / (F32(4.0) * F32(h)))
    s_w = np.sqrt(F32(1.0) / (F32(2.0) * F32(w)))
    s_h = np.sqrt(F32(1.0) / (F32(2.0) * F32(h)))
    idx = np.arange(coeff.size)
    first_row = idx < w
    first_col = (idx % w) == 0
    sc = np.where(first_row, s_k0_w, s_w).astype(F32)
    sc = (F32(1.0) * sc).astype(F32)
    sc = (sc * np.where(first_col, s_k0_h, s_h).astype(F32)).astype(F32)
    return (sc * coeff.astype(F32)).astype(F32)


def ordering_values(coeff_flat, ordering=ORDER_ENERGY, w=None, h=None):
    """The f32 value whose `total_cmp` the reference sorts by, descending (:214-232,235-280)."""
    c = np.asarray(coeff_flat, dtype=F32).ravel()
    if ordering == ORDER_ENERGY:
        return (c * c).astype(F32)
    v = _ortho_scaled(c, w, h)
    if ordering == ORDER_ENERGY_ORTHO:
        return (v * v).astype(F32)
    if ordering == ORDER_LEGACY:
        return v
    raise ValueError(ordering)


def obtain_indices(coeff_flat, ordering=ORDER_ENERGY, w=None, h=None, k=None):
    """src/algorithm.rs:200-210: indices 1..n-1, *stable* sort descending (ties keep ascending
    index).  `coeff_flat` is rounded to f32 first (the reference's coefficients are f32).
    Returns all n-1 indices, or the first k."""
    v = ordering_values(coeff_flat, ordering, w, h)
    key = _total_cmp_key(v)[1:]
    order = np.argsort(~key, kind='stable') + 1
    return order if k is None else order[:k]


# ------------------------------------------------------------------------------------------------
# embed / extract / similarity -- src/algorithm.rs:382-432, 543-593, 696-714
# ------------------------------------------------------------------------------------------------
def _insert(method, alpha, orig, w):
    orig = orig.astype(F32); w = w.astype(F32); alpha = F32(alpha)
    if method == 1:   # :414-416
        return (orig + (alpha * w).astype(F32)).astype(F32)
    if method == 2:   # :420-424
        return (orig * (F32(1.0) + (alpha * w).astype(F32)).astype(F32)).astype(F32)
    if method == 3:   # :428-432
        return (orig * np.exp((alpha * w).astype(F32)).astype(F32)).astype(F32)
    raise ValueError(method)


def embed_watermark(coeff_flat, indices, marks, method=2, alpha=0.1):
    """src/algorithm.rs:382-410.  coeff_flat: f32[n] (modified copy returned)."""
    c = np.array(coeff_flat, dtype=F32).ravel().copy()
    if len(marks) == 1:
        n = min(len(indices), len(marks[0]))
        idx = np.asarray(indices[:n])
        c[idx] = _insert(method, alpha, c[idx], np.asarray(marks[0][:n], dtype=F32))
    else:
        orig = c.copy()
        for m in marks:
            n = min(len(indices), len(m))
            idx = np.asarray(indices[:n])
            upd = _insert(method, alpha, orig[idx], np.asarray(m[:n], dtype=F32))
            change = (upd - orig[idx]).astype(F32)
            c[idx] = (c[idx] + change).astype(F32)
    return c


def extract_watermark(base_flat, indices, derived_flat, n, method=2, alpha=0.1):
    """src/algorithm.rs:543-593."""
    base_flat = np.asarray(base_flat, dtype=F32).ravel()
    derived_flat = np.asarray(derived_flat, dtype=F32).ravel()
    if derived_flat.size != base_flat.size:
        raise ValueError('Derived coefficient length not equal to base coefficient length.')
    if n >= base_flat.size:
        raise ValueError('Desired extraction length exceeds available coefficients.')
    idx = np.asarray(indices[:n])
    b = base_flat[idx]; d = derived_flat[idx]; alpha = F32(alpha)
    with np.errstate(divide='ignore', invalid='ignore'):
        if method == 1:
            return ((d - b).astype(F32) / alpha).astype(F32)
        if method == 2:
            return ((d - b).astype(F32) / (b * alpha).astype(F32)).astype(F32)
        if method == 3:
            return (np.log((d / b).astype(F32)).astype(F32) / alpha).astype(F32)
    raise ValueError(method)


def similarity(extracted, mark):
    """src/algorithm.rs:696-714: sequential f32 accumulation, nom/sqrt(den)."""
    e = np.asarray(extracted, dtype=F32); c = np.asarray(mark, dtype=F32)
    assert e.size == c.size
    nom = F32(0.0); den = F32(0.0)
    for a, b in zip(e, c):
        nom = F32(nom + F32(a * b))
        den = F32(den + F32(a * a))
    with np.errstate(divide='ignore', invalid='ignore'):
        return F32(nom / np.sqrt(den))


# ------------------------------------------------------------------------------------------------
# whole-path helpers (mirror Writer / Reader, src/algorithm.rs:286-379, 441-539)
# ------------------------------------------------------------------------------------------------
def forward(rgb8_or_f32, dtype=np.float64):
    """Writer::new / Reader::new_impl up to the coefficients: returns (coeff[H][W] f32-rounded,
    i plane, q plane)."""
    rgb = rgb8_to_rgb32f(rgb8_or_f32) if rgb8_or_f32.dtype == np.uint8 else rgb8_or_f32.astype(F32)
    y, i, q = rgb32f_to_yiq(rgb)
    c = dct2_2d(y, DCT2, dtype)
    return c, i, q


def embed(rgb, marks, method=2, alpha=0.1, ordering=ORDER_ENERGY, dtype=np.float64, to_rgb8=True):
    """Writer::new(img, cfg).mark(marks) [+ into_rgb8()].  Returns (image, indices, coeff_before)."""
    c, i, q = forward(rgb, dtype)
    h, w = c.shape
    c32 = c.astype(F32)
    kmax = max(len(m) for m in marks)
    idx = obtain_indices(c32.ravel(), ordering, w, h, k=min(kmax, w * h - 1))
    if dtype == np.float64:
        # keep the un-touched coefficients in f64 so the inverse is the "exact" answer
        c_mod = c.ravel().copy()
        c_mod[idx] = embed_watermark(c32.ravel(), idx, marks, method, alpha)[idx]
    else:
        c_mod = embed_watermark(c32.ravel(), idx, marks, method, alpha)
    y2 = dct2_2d(c_mod.reshape(h, w), DCT3, dtype).astype(F32)
    out = yiq_to_rgb32f(y2, i, q)
    return (rgb32f_to_rgb8(out) if to_rgb8 else out), idx, c32


def extract(base_rgb, derived_rgb, n, method=2, alpha=0.1, ordering=ORDER_ENERGY, dtype=np.float64):
    """Reader::base + Reader::derived + extract.  Returns (extracted f32[n], indices)."""
    cb, _, _ = forward(base_rgb, dtype)
    cd, _, _ = forward(derived_rgb, dtype)
    if cb.shape != cd.shape:
        raise ValueError('Derived coefficient length not equal to base coefficient length.')
    h, w = cb.shape
    idx = obtain_indices(cb.astype(F32).ravel(), ordering, w, h, k=min(n, w * h - 1))
    return extract_watermark(cb.astype(F32).ravel(), idx, cd.astype(F32).ravel(), n, method, alpha), idx


# ------------------------------------------------------------------------------------------------
# synthetic frames for the benchmark configs -- SURVEY.md section 8(d) generator (integer only)
# ------------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _sm64(x):
    with np.errstate(over='ignore'):
        z = (x + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_frame(w, h, seed, img=0):
    """Deterministic natural-image-like RGB8 frame [h][w][3] (value-noise octaves 2..8 + dither).
    Bit-identical to the device generator `ssw_synth_frame_rgb8` (include/ssw.h)."""
    with np.errstate(over='ignore'):
        base = np.uint64(seed) ^ (np.uint64(img) * np.uint64(0x9E3779B97F4A7C15))
    x = np.arange(w, dtype=np.uint64)[None, :]
    y = np.arange(h, dtype=np.uint64)[:, None]
    out = np.zeros((h, w, 3), dtype=np.uint8)
    for c in range(3):
        acc = np.zeros((h, w), dtype=np.uint64)
        for o in range(2, 9):
            s = np.uint64(1 << o)
            X = x >> np.uint64(o); Y = y >> np.uint64(o)
            fx = x & (s - np.uint64(1)); fy = y & (s - np.uint64(1))
            tag = base ^ (np.uint64(o) << np.uint64(58)) ^ (np.uint64(c) << np.uint64(56))

            def L(XX, YY):
                return _sm64(tag ^ (YY << np.uint64(28)) ^ XX) >> np.uint64(56)
            top = L(X, Y) * (s - fx) + L(X + np.uint64(1), Y) * fx
            bot = L(X, Y + np.uint64(1)) * (s - fx) + L(X + np.uint64(1), Y + np.uint64(1)) * fx
            v = (top * (s - fy) + bot * fy) >> np.uint64(2 * o)
            acc += v << np.uint64(o)
        noise = _sm64(base ^ np.uint64(0xABCDEF) ^ (np.uint64(c) << np.uint64(56)) ^ (y << np.uint64(28)) ^ x) >> np.uint64(61)
        val = (acc // np.uint64(508)).astype(np.int64) + noise.astype(np.int64) - 4
        out[..., c] = np.clip(val, 0, 255).astype(np.uint8)
    return out
