"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

Emulation of the JPEG decode the reference's golden test went through:
`image::open("tests/porcelain_cat_grey_background.jpg")` in /root/reference/tests/single_simple.rs:11-14
resolves (Cargo.lock: image 0.24.3 -> jpeg-decoder 0.2.6) to jpeg-decoder's x86-64 SSSE3 path.
That decode differs from libjpeg-turbo (PIL / cv2) in 3.3 % of bytes by 1..3 LSB, which is enough
to reorder ~35 % of the top-1000 coefficient ranks, so the golden PNG
(tests/watermarked_with_1.png, asserted pixel-exact at tests/single_simple.rs:36-43) can only be
reproduced from this decoder's output.  SURVEY.md Appendix B.3 is the recipe followed here:

  * baseline sequential Huffman decode, 4:2:0, no restart intervals;
  * dequantise + 8x8 IDCT in the 16-bit SSSE3 form (saturating adds, pmulhrsw), column pass,
    transpose, column pass, transpose, +((128<<6)+32) >> 6, unsigned saturate;
  * H2V2 "fancy" chroma upsampling (3:1 triangle filter, rows and columns);
  * YCbCr->RGB: 16-bit SSSE3 for the first (W/8 - 1)*8 pixels of each line, 20-bit scalar after.

Only used to (re)generate tests/golden/cat_rgb8.npz (sha256 of the pixels is pinned in
SURVEY.md Appendix C.1 and checked by oracle/make_golden.py).
"""
import numpy as np

ZIGZAG = [
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20,
    13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52,
    45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
]


def _i16(a):
    return np.asarray(a).astype(np.int64).astype(np.int16)  # wrap


def _adds(a, b):
    return np.clip(a.astype(np.int32) + b.astype(np.int32), -32768, 32767).astype(np.int16)


def _subs(a, b):
    return np.clip(a.astype(np.int32) - b.astype(np.int32), -32768, 32767).astype(np.int16)


def _mulhrs(a, c):
    # pmulhrsw: ((a*c >> 14) + 1) >> 1 on 16-bit lanes
    return ((((a.astype(np.int32) * np.int32(c)) >> 14) + 1) >> 1).astype(np.int16)


def _idct8(d):
    """One 1-D pass over d[0..8] (each an array of int16 lanes)."""
    p2, p3 = d[2], d[6]
    p1 = _mulhrs(_adds(p2, p3), 17734)
    t2 = _subs(_subs(p1, p3), _mulhrs(p3, 27779))
    t3 = _adds(p1, _mulhrs(p2, 25079))
    p2, p3 = d[0], d[4]
    t0 = _adds(p2, p3)
    t1 = _subs(p2, p3)
    x0 = _adds(t0, t3)
    x3 = _subs(t0, t3)
    x1 = _adds(t1, t2)
    x2 = _subs(t1, t2)
    t0, t1, t2, t3 = d[7], d[5], d[3], d[1]
    p3 = _adds(t0, t2)
    p4 = _adds(t1, t3)
    p1 = _adds(t0, t3)
    p2 = _adds(t1, t2)
    p5 = _adds(p3, p4)
    p5 = _adds(p5, _mulhrs(p5, 5763))
    t0 = _mulhrs(t0, 9786)
    t1 = _adds(_adds(t1, t1), _mulhrs(t1, 1741))
    t2 = _adds(_adds(t2, _adds(t2, t2)), _mulhrs(t2, 2383))
    t3 = _adds(t3, _mulhrs(t3, 16427))
    p1 = _subs(p5, _mulhrs(p1, 29490))
    p2 = _subs(_subs(_subs(p5, p2), p2), _mulhrs(p2, 18446))
    p3 = _subs(_mulhrs(p3, -31509), p3)
    p4 = _mulhrs(p4, -12785)
    t3 = _adds(_adds(t3, p1), p4)
    t2 = _adds(_adds(t2, p2), p3)
    t1 = _adds(_adds(t1, p2), p4)
    t0 = _adds(_adds(t0, p1), p3)
    return [_adds(x0, t3), _adds(x1, t2), _adds(x2, t1), _adds(x3, t0),
            _subs(x3, t0), _subs(x2, t1), _subs(x1, t2), _subs(x0, t3)]


def idct_blocks(coefs, qt):
    """coefs: [nblocks, 64] int (natural order), qt: [64] (natural order) -> [nblocks, 8, 8] uint8."""
    nb = coefs.shape[0]
    d = _i16((coefs.astype(np.int64) * qt.astype(np.int64)[None, :]))  # pmullw (wrap)
    d = _i16(d.astype(np.int64) << 3)                                    # psllw 3 (wrap)
    d = d.reshape(nb, 8, 8)
    rows = [d[:, i, :] for i in range(8)]
    rows = _idct8(rows)                                   # column pass (lanes = columns)
    m = np.stack(rows, axis=1).transpose(0, 2, 1)         # transpose
    rows = _idct8([m[:, i, :] for i in range(8)])
    m = np.stack(rows, axis=1).transpose(0, 2, 1)         # transpose back
    v = _adds(m, np.int16((128 << 6) + 32)) >> 6
    return np.clip(v, 0, 255).astype(np.uint8)


class _Bits:
    def __init__(self, data, pos):
        self.d = data
        self.p = pos
        self.acc = 0
        self.n = 0

    def _fill(self):
        b = self.d[self.p]
        self.p += 1
        if b == 0xFF:
            b2 = self.d[self.p]
            if b2 == 0:
                self.p += 1
            else:
                raise ValueError('marker inside scan: %x' % b2)
        self.acc = (self.acc << 8) | b
        self.n += 8

    def bit(self):
        if self.n == 0:
            self._fill()
        self.n -= 1
        return (self.acc >> self.n) & 1

    def bits(self, k):
        v = 0
        for _ in range(k):
            v = (v << 1) | self.bit()
        return v


def _build_huff(counts, symbols):
    table = {}
    code = 0
    k = 0
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            table[(length, code)] = symbols[k]
            k += 1
            code += 1
        code <<= 1
    return table


def _decode_sym(br, table):
    code = 0
    for length in range(1, 17):
        code = (code << 1) | br.bit()
        s = table.get((length, code))
        if s is not None:
            return s
    raise ValueError('bad huffman code')


def _extend(v, t):
    return v - ((1 << t) - 1) if t and v < (1 << (t - 1)) else v


def decode_planes(data):
    """Parse + entropy-decode + IDCT.  Returns (W, H, [Y, Cb, Cr] planes incl. MCU padding, comps)."""
    data = bytes(data)
    assert data[0:2] == b'\xff\xd8'
    pos = 2
    qts, dc_t, ac_t = {}, {}, {}
    comps = None
    W = H = None
    while True:
        assert data[pos] == 0xFF
        m = data[pos + 1]
        pos += 2
        if m == 0xD8 or (0xD0 <= m <= 0xD7):
            continue
        ln = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2:pos + ln]
        if m == 0xDB:
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                i += 1
                q = np.zeros(64, np.int64)
                for k in range(64):
                    if pq:
                        q[ZIGZAG[k]] = (seg[i] << 8) | seg[i + 1]
                        i += 2
                    else:
                        q[ZIGZAG[k]] = seg[i]
                        i += 1
                qts[tq] = q
        elif m == 0xC0:
            assert seg[0] == 8
            H = (seg[1] << 8) | seg[2]
            W = (seg[3] << 8) | seg[4]
            nc = seg[5]
            comps = []
            for c in range(nc):
                cid, hv, tq = seg[6 + 3 * c: 9 + 3 * c]
                comps.append({'id': cid, 'h': hv >> 4, 'v': hv & 15, 'tq': tq})
        elif m in (0xC1, 0xC2, 0xC3, 0xC9, 0xCA):
            raise ValueError('only baseline sequential JPEG is emulated')
        elif m == 0xC4:
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                counts = list(seg[i + 1:i + 17])
                n = sum(counts)
                syms = list(seg[i + 17:i + 17 + n])
                i += 17 + n
                (ac_t if tc else dc_t)[th] = _build_huff(counts, syms)
        elif m == 0xDD:
            assert ((seg[0] << 8) | seg[1]) == 0, 'restart intervals not emulated'
        elif m == 0xDA:
            ns = seg[0]
            for k in range(ns):
                cs, tt = seg[1 + 2 * k], seg[2 + 2 * k]
                for c in comps:
                    if c['id'] == cs:
                        c['td'], c['ta'] = tt >> 4, tt & 15
            pos += ln
            break
        pos += ln

    hmax = max(c['h'] for c in comps)
    vmax = max(c['v'] for c in comps)
    mcux = (W + 8 * hmax - 1) // (8 * hmax)
    mcuy = (H + 8 * vmax - 1) // (8 * vmax)
    for c in comps:
        c['bw'] = mcux * c['h']
        c['bh'] = mcuy * c['v']
        c['coef'] = np.zeros((c['bh'], c['bw'], 64), np.int64)
        c['pred'] = 0
        c['w'] = (W * c['h'] + hmax - 1) // hmax
        c['hgt'] = (H * c['v'] + vmax - 1) // vmax
    br = _Bits(data, pos)
    for my in range(mcuy):
        for mx in range(mcux):
            for c in comps:
                for by in range(c['v']):
                    for bx in range(c['h']):
                        blk = c['coef'][my * c['v'] + by, mx * c['h'] + bx]
                        t = _decode_sym(br, dc_t[c['td']])
                        diff = _extend(br.bits(t), t) if t else 0
                        c['pred'] += diff
                        blk[0] = c['pred']
                        k = 1
                        while k < 64:
                            rs = _decode_sym(br, ac_t[c['ta']])
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r == 15:
                                    k += 16
                                    continue
                                break
                            k += r
                            blk[ZIGZAG[k]] = _extend(br.bits(s), s)
                            k += 1
    planes = []
    for c in comps:
        px = idct_blocks(c['coef'].reshape(-1, 64), qts[c['tq']])
        px = px.reshape(c['bh'], c['bw'], 8, 8).transpose(0, 2, 1, 3).reshape(c['bh'] * 8, c['bw'] * 8)
        planes.append(px)
    return W, H, planes, comps


def upsample_h2v2(plane, in_w, in_h, out_w, out_h):
    """jpeg-decoder UpsamplerH2V2: triangle filter; far row = previous (even out row) / next (odd)."""
    p = plane.astype(np.int64)
    out = np.zeros((out_h, out_w), np.uint8)
    for row in range(out_h):
        near = row // 2
        far = near - 1 if row % 2 == 0 else near + 1
        far = min(max(far, 0), in_h - 1)
        t = 3 * p[near, :in_w] + p[far, :in_w]
        o = np.zeros(2 * in_w, np.int64)
        o[0] = (t[0] + 2) >> 2
        o[1] = (3 * t[0] + t[1] + 8) >> 4
        i = np.arange(1, in_w - 1)
        o[2 * i] = (3 * t[i] + t[i - 1] + 8) >> 4
        o[2 * i + 1] = (3 * t[i] + t[i + 1] + 8) >> 4
        o[2 * in_w - 2] = (3 * t[in_w - 1] + t[in_w - 2] + 8) >> 4
        o[2 * in_w - 1] = (t[in_w - 1] + 2) >> 2
        out[row] = o[:out_w].astype(np.uint8)
    return out


def ycbcr_to_rgb_line_mix(y, cb, cr, cb_const=11277):
    """One image worth of lines: SSSE3 16-bit for the first (W/8-1)*8 pixels, scalar 20-bit after."""
    H, W = y.shape
    out = np.zeros((H, W, 3), np.uint8)
    nv = max(W // 8 - 1, 0) * 8
    # --- SIMD part
    ys = _adds(_i16(y[:, :nv].astype(np.int64) << 6), np.int16(32))
    cbs = _subs(_i16(cb[:, :nv].astype(np.int64) << 6), np.int16(128 << 6))
    crs = _subs(_i16(cr[:, :nv].astype(np.int64) << 6), np.int16(128 << 6))
    cr_1402 = _adds(_mulhrs(crs, 13173), crs)
    cb_0344 = _mulhrs(cbs, cb_const)
    cr_0714 = _mulhrs(crs, 23401)
    cb_1772 = _adds(_mulhrs(cbs, 25297), cbs)
    r = _adds(ys, cr_1402) >> 6
    g = _subs(ys, _adds(cb_0344, cr_0714)) >> 6
    b = _adds(ys, cb_1772) >> 6
    out[:, :nv, 0] = np.clip(r, 0, 255)
    out[:, :nv, 1] = np.clip(g, 0, 255)
    out[:, :nv, 2] = np.clip(b, 0, 255)
    # --- scalar tail
    yy = (y[:, nv:].astype(np.int64) << 20) + (1 << 19)
    cbb = cb[:, nv:].astype(np.int64) - 128
    crr = cr[:, nv:].astype(np.int64) - 128
    out[:, nv:, 0] = np.clip((yy + 1470104 * crr) >> 20, 0, 255)
    out[:, nv:, 1] = np.clip((yy - 360857 * cbb - 748830 * crr) >> 20, 0, 255)
    out[:, nv:, 2] = np.clip((yy + 1858077 * cbb) >> 20, 0, 255)
    return out


def decode_rgb8(data):
    W, H, planes, comps = decode_planes(data)
    assert len(comps) == 3 and (comps[0]['h'], comps[0]['v']) == (2, 2) \
        and (comps[1]['h'], comps[1]['v']) == (1, 1), 'only 4:2:0 YCbCr is emulated'
    y = planes[0][:H, :W]
    cb = upsample_h2v2(planes[1], comps[1]['w'], comps[1]['hgt'], W, H)
    cr = upsample_h2v2(planes[2], comps[2]['w'], comps[2]['hgt'], W, H)
    return ycbcr_to_rgb_line_mix(y, cb, cr)


if __name__ == '__main__':
    import hashlib
    import sys
    img = decode_rgb8(open(sys.argv[1], 'rb').read())
    print(img.shape, hashlib.sha256(img.tobytes()).hexdigest())
