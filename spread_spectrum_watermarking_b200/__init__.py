"""spread_spectrum_watermarking_b200 -- host-side mirror of the reference crate's public API
(/root/reference/src/lib.rs:75-85) over the libssw C ABI (include/ssw.h).

Same names, argument meaning and error behaviour as the Rust crate, so the parity tests read like
the reference's own tests:

    import spread_spectrum_watermarking_b200 as wm
    mark = wm.MarkBuf.generate_normal(1000)
    res = wm.Writer.new(image, wm.WriteConfig.default()).mark([mark])      # Rgb32F, [h][w][3] f32
    reader = wm.Reader.base(image, wm.ReadConfig.default())
    derived = wm.Reader.derived(res)            # or the RGB8 image from Writer.mark_rgb8()
    extracted = reader.extract(derived, 1000)
    wm.Tester.new(extracted).similarity(mark).exceeds_sigma(6.0)

Images are numpy arrays standing in for `image::DynamicImage`: uint8 or float32, shape [h][w][3]
(or [h][w][4] -- alpha dropped -- or [h][w] luma, replicated), like `into_rgb32f()` would produce.
All arithmetic runs in the CUDA library; where the reference panics this raises `SswError`.
This module is plumbing only -- it never computes any part of the hot path on the CPU.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import SswError, lib, check, ssw_config

__all__ = [
    'Context', 'default_context', 'Insertion', 'Extraction', 'OrderingMethod', 'WriteConfig', 'ReadConfig',
    'Writer', 'Reader', 'ReaderDerived', 'Tester', 'Similarity', 'MarkBuf', 'Mark', 'Bank', 'dct2d', 'yiq',
    'SswError',
]


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


# ------------------------------------------------------------------------------------------------
# context (stands in for the DctPlanner each Writer/Reader owns, src/algorithm.rs:288,443)
# ------------------------------------------------------------------------------------------------
class Context:
    def __init__(self, device=0, stream=None):
        h = ctypes.c_void_p()
        if stream is None:
            check(lib.ssw_ctx_create(int(device), ctypes.byref(h)))
        else:
            check(lib.ssw_ctx_create_on_stream(int(device), ctypes.c_void_p(int(stream)), ctypes.byref(h)))
        self.handle = h
        self.device = int(device)

    def synchronize(self):
        check(lib.ssw_ctx_synchronize(self.handle))

    @property
    def stream(self):
        return lib.ssw_ctx_stream(self.handle)

    @property
    def launch_count(self):
        return int(lib.ssw_ctx_launch_count(self.handle))

    def profile_begin(self):
        check(lib.ssw_ctx_profile_begin(self.handle))

    def profile_end(self):
        """-> {kernel name: {'launches': n, 'ms': total}} for the launches since profile_begin()"""
        import json
        buf = ctypes.create_string_buffer(1 << 16)
        check(lib.ssw_ctx_profile_end(self.handle, buf, len(buf)))
        return json.loads(buf.value.decode())

    def set_tiling(self, row_pairs=0, col_pairs=0):
        check(lib.ssw_ctx_set_tiling(self.handle, int(row_pairs), int(col_pairs)))

    def last_topk_fallbacks(self):
        return int(lib.ssw_ctx_last_topk_fallbacks(self.handle))

    def close(self):
        if self.handle:
            lib.ssw_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def default_context():
    """One lazily created context on device 0 (raises SswError if there is no CUDA device)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ------------------------------------------------------------------------------------------------
# configuration -- src/algorithm.rs:66-171
# ------------------------------------------------------------------------------------------------
class _Method:
    def __init__(self, option, alpha=None, fn=None):
        self.option, self.alpha, self.fn = option, alpha, fn

    def __repr__(self):
        return '%s::Custom' % type(self).__name__ if self.option == 0 else \
            '%s::Option%d(%r)' % (type(self).__name__, self.option, self.alpha)

    @classmethod
    def Option1(cls, alpha):
        return cls(1, float(alpha))

    @classmethod
    def Option2(cls, alpha):
        return cls(2, float(alpha))

    @classmethod
    def Option3(cls, alpha):
        return cls(3, float(alpha))

    @classmethod
    def Custom(cls, fn):
        return cls(0, None, fn)


class Insertion(_Method):
    """src/algorithm.rs:68-78"""


class Extraction(_Method):
    """src/algorithm.rs:115-125"""


class OrderingMethod:
    """src/algorithm.rs:143-152"""
    Energy = 0
    EnergyOrthogonal = 1
    Legacy = 2

    class Custom:
        def __init__(self, fn):
            self.fn = fn


class WriteConfig:
    def __init__(self, insertion=None, ordering=OrderingMethod.Energy):
        self.insertion = insertion if insertion is not None else Insertion.Option2(0.1)
        self.ordering = ordering

    @classmethod
    def default(cls):
        """src/algorithm.rs:105-112"""
        return cls()


class ReadConfig:
    def __init__(self, extraction=None, ordering=OrderingMethod.Energy):
        self.extraction = extraction if extraction is not None else Extraction.Option2(0.1)
        self.ordering = ordering

    @classmethod
    def default(cls):
        """src/algorithm.rs:133-140"""
        return cls()


def _c_config(method, ordering):
    if method.option == 0:
        raise SswError(_lib.SSW_ERR_UNSUPPORTED,
                       'Custom insertion/extraction closures are host code; the device path has no CPU fallback')
    if isinstance(ordering, OrderingMethod.Custom):
        raise SswError(_lib.SSW_ERR_UNSUPPORTED,
                       'Custom ordering closures are host code; the device path has no CPU fallback')
    return ssw_config(int(method.option), float(method.alpha), int(ordering))


def _as_rgb(image):
    """DynamicImage -> contiguous [h][w][3] uint8 or float32 (into_rgb8 / into_rgb32f layout)."""
    a = np.asarray(image)
    if a.ndim == 2:
        a = np.repeat(a[:, :, None], 3, axis=2)
    if a.ndim != 3 or a.shape[2] not in (3, 4):
        raise SswError(_lib.SSW_ERR_INVALID, 'image must be [h][w], [h][w][3] or [h][w][4]')
    if a.shape[2] == 4:
        a = a[:, :, :3]
    if a.dtype == np.uint8:
        return np.ascontiguousarray(a)
    if a.dtype in (np.float32, np.float64):
        return np.ascontiguousarray(a, dtype=np.float32)
    raise SswError(_lib.SSW_ERR_INVALID, 'image dtype must be uint8 or float32')


# ------------------------------------------------------------------------------------------------
# marks -- src/algorithm.rs:596-666
# ------------------------------------------------------------------------------------------------
def _mark_data(m):
    """trait Mark: anything with .data() or AsRef<[f32]>."""
    if hasattr(m, 'data') and callable(m.data):
        m = m.data()
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).ravel())


class MarkBuf:
    def __init__(self, data=None):
        self._data = np.zeros(0, np.float32) if data is None else _mark_data(data).copy()

    @classmethod
    def new(cls):
        return cls()

    @classmethod
    def generate_normal(cls, length, seed=0, ctx=None):
        """src/algorithm.rs:619-626 (seed 0 = OS entropy, like thread_rng)."""
        ctx = ctx or default_context()
        out = np.empty(int(length), np.float32)
        check(lib.ssw_mark_generate_normal(ctx.handle, ctypes.c_uint64(seed), out.size, _ptr(out)))
        return cls(out)

    @classmethod
    def from_(cls, data):
        return cls(data)

    def data(self):
        return self._data

    def set_data(self, data):
        self._data = _mark_data(data).copy()

    def __len__(self):
        return self._data.size


Mark = MarkBuf


# ------------------------------------------------------------------------------------------------
# Writer -- src/algorithm.rs:286-433
# ------------------------------------------------------------------------------------------------
class Writer:
    def __init__(self, image, config=None, ctx=None):
        config = config or WriteConfig.default()
        self.ctx = ctx or default_context()
        img = _as_rgb(image)
        self.height, self.width = img.shape[:2]
        cfg = _c_config(config.insertion, config.ordering)
        h = ctypes.c_void_p()
        fn = lib.ssw_writer_new_rgb8 if img.dtype == np.uint8 else lib.ssw_writer_new_rgb32f
        check(fn(self.ctx.handle, _ptr(img), self.width, self.height, ctypes.byref(cfg), ctypes.byref(h)))
        self.handle = h

    @classmethod
    def new(cls, image, config=None, ctx=None):
        return cls(image, config, ctx)

    def coefficient_image(self):
        out = np.empty((self.height, self.width), np.float32)
        check(lib.ssw_writer_coefficients(self.handle, _ptr(out)))
        return out

    def indices(self, n):
        out = np.empty(int(n), np.uint64)
        check(lib.ssw_writer_indices(self.handle, _ptr(out), out.size))
        return out

    def embed(self, marks):
        datas = [_mark_data(m) for m in marks]
        n = len(datas)
        ptrs = (ctypes.c_void_p * max(n, 1))(*[d.ctypes.data for d in datas])
        lens = (ctypes.c_size_t * max(n, 1))(*[d.size for d in datas])
        check(lib.ssw_writer_embed(self.handle, ptrs, lens, n))

    def result(self):
        """-> Rgb32F image ([h][w][3] float32, clamped to [0,1]); consumes the writer."""
        out = np.empty((self.height, self.width, 3), np.float32)
        check(lib.ssw_writer_result_rgb32f(self.handle, _ptr(out)))
        return out

    def result_rgb8(self):
        """`result().into_rgb8()` with the quantisation fused into the last kernel."""
        out = np.empty((self.height, self.width, 3), np.uint8)
        check(lib.ssw_writer_result_rgb8(self.handle, _ptr(out)))
        return out

    def mark(self, marks):
        self.embed(marks)
        return self.result()

    def mark_rgb8(self, marks):
        self.embed(marks)
        return self.result_rgb8()

    def close(self):
        if getattr(self, 'handle', None):
            # an object that outlives its context (e.g. kept alive by a traceback) must not touch the destroyed context
            if getattr(getattr(self, 'ctx', None), 'handle', None):
                lib.ssw_writer_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# Reader / ReaderDerived -- src/algorithm.rs:435-594
# ------------------------------------------------------------------------------------------------
class Reader:
    def __init__(self, image, is_base, config=None, ctx=None):
        self.ctx = ctx or default_context()
        img = _as_rgb(image)
        self.height, self.width = img.shape[:2]
        self.is_base = bool(is_base)
        h = ctypes.c_void_p()
        u8 = img.dtype == np.uint8
        if is_base:
            config = config or ReadConfig.default()
            cfg = _c_config(config.extraction, config.ordering)
            fn = lib.ssw_reader_base_rgb8 if u8 else lib.ssw_reader_base_rgb32f
            check(fn(self.ctx.handle, _ptr(img), self.width, self.height, ctypes.byref(cfg), ctypes.byref(h)))
        else:
            fn = lib.ssw_reader_derived_rgb8 if u8 else lib.ssw_reader_derived_rgb32f
            check(fn(self.ctx.handle, _ptr(img), self.width, self.height, ctypes.byref(h)))
        self.handle = h

    @classmethod
    def base(cls, image, config=None, ctx=None):
        return cls(image, True, config, ctx)

    @classmethod
    def derived(cls, image, ctx=None):
        return ReaderDerived(image, ctx)

    def coefficients(self):
        out = np.empty(self.height * self.width, np.float32)
        check(lib.ssw_reader_coefficients(self.handle, _ptr(out)))
        return out

    def indices(self, n=None):
        """First n ordered indices (the reference returns all w*h-1; default here: all of them)."""
        n = self.width * self.height - 1 if n is None else int(n)
        out = np.empty(n, np.uint64)
        check(lib.ssw_reader_indices(self.handle, _ptr(out), out.size))
        return out

    def extract(self, derived, extracted):
        """`extracted` is a length or a float32 array to fill (the reference takes &mut [f32])."""
        d = derived.reader if isinstance(derived, ReaderDerived) else derived
        if isinstance(extracted, (int, np.integer)):
            extracted = np.empty(int(extracted), np.float32)
        if extracted.dtype != np.float32 or not extracted.flags.c_contiguous:
            raise SswError(_lib.SSW_ERR_INVALID, 'extracted must be a contiguous float32 array')
        check(lib.ssw_reader_extract(self.handle, d.handle, _ptr(extracted), extracted.size))
        return extracted

    def close(self):
        if getattr(self, 'handle', None):
            # an object that outlives its context (e.g. kept alive by a traceback) must not touch the destroyed context
            if getattr(getattr(self, 'ctx', None), 'handle', None):
                lib.ssw_reader_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ReaderDerived:
    def __init__(self, image, ctx=None):
        self.reader = Reader(image, False, None, ctx)

    @classmethod
    def new(cls, image, ctx=None):
        return cls(image, ctx)

    def coefficients(self):
        return self.reader.coefficients()


# ------------------------------------------------------------------------------------------------
# Tester / Similarity -- src/algorithm.rs:668-715
# ------------------------------------------------------------------------------------------------
class Similarity:
    def __init__(self, similarity):
        self.similarity = np.float32(similarity)

    def exceeds_sigma(self, n_sigma):
        return bool(self.similarity > np.float32(n_sigma))

    def __repr__(self):
        return 'Similarity { similarity: %r }' % float(self.similarity)


class Tester:
    def __init__(self, extracted_watermark, ctx=None):
        self.ctx = ctx or default_context()
        self.extracted = _mark_data(extracted_watermark)

    @classmethod
    def new(cls, extracted_watermark, ctx=None):
        return cls(extracted_watermark, ctx)

    def similarity(self, comparison_watermark):
        c = _mark_data(comparison_watermark)
        if c.size != self.extracted.size:  # assert_eq!, src/algorithm.rs:697-700
            raise SswError(_lib.SSW_ERR_INVALID, 'assertion failed: extracted and comparison lengths differ')
        out = ctypes.c_float()
        check(lib.ssw_similarity(self.ctx.handle, _ptr(self.extracted), _ptr(c), c.size, ctypes.byref(out)))
        return Similarity(out.value)

    def similarity_bank(self, bank):
        return bank.similarity(self.extracted[None, :])[0]


class Bank:
    """Device-resident bank of stored marks [n_marks][n] (README.md:62)."""

    def __init__(self, marks=None, ctx=None, normal=None):
        self.ctx = ctx or default_context()
        h = ctypes.c_void_p()
        if normal is not None:
            seed, n_marks, n = normal
            check(lib.ssw_bank_create_normal(self.ctx.handle, ctypes.c_uint64(seed), n_marks, n, ctypes.byref(h)))
            self.n_marks, self.n = int(n_marks), int(n)
        else:
            m = np.ascontiguousarray(np.asarray(marks, dtype=np.float32))
            if m.ndim != 2:
                raise SswError(_lib.SSW_ERR_INVALID, 'bank must be [n_marks][n]')
            self.n_marks, self.n = m.shape
            check(lib.ssw_bank_create(self.ctx.handle, _ptr(m), self.n_marks, self.n, ctypes.byref(h)))
        self.handle = h

    @classmethod
    def normal(cls, seed, n_marks, n, ctx=None):
        return cls(ctx=ctx, normal=(seed, n_marks, n))

    def row(self, index):
        out = np.empty(self.n, np.float32)
        check(lib.ssw_bank_row(self.handle, int(index), _ptr(out)))
        return out

    def similarity(self, extracted):
        e = np.ascontiguousarray(np.asarray(extracted, dtype=np.float32))
        if e.ndim == 1:
            e = e[None, :]
        if e.shape[1] != self.n:
            raise SswError(_lib.SSW_ERR_INVALID, 'assertion failed: extracted and bank mark lengths differ')
        out = np.empty((e.shape[0], self.n_marks), np.float32)
        check(lib.ssw_bank_similarity(self.handle, _ptr(e), e.shape[0], _ptr(out)))
        return out

    def close(self):
        if getattr(self, 'handle', None):
            # an object that outlives its context (e.g. kept alive by a traceback) must not touch the destroyed context
            if getattr(getattr(self, 'ctx', None), 'handle', None):
                lib.ssw_bank_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# dct2d / yiq modules -- src/dct2d.rs, src/yiq.rs
# ------------------------------------------------------------------------------------------------
class dct2d:
    class Type:
        DCT2 = 0
        DCT2Orthogonal = 1
        DCT3 = 2

    @staticmethod
    def dct2_2d(transform_type, width, height, data, ctx=None):
        """src/dct2d.rs:83: `data` (float32, width*height, row-major) is transformed in place."""
        ctx = ctx or default_context()
        if data.dtype != np.float32 or not data.flags.c_contiguous:
            raise SswError(_lib.SSW_ERR_INVALID, 'data must be a contiguous float32 array')
        if data.size != width * height:  # assert_eq!, src/dct2d.rs:90
            raise SswError(_lib.SSW_ERR_INVALID, 'assertion failed: data.len() == width * height')
        check(lib.ssw_dct2_2d(ctx.handle, int(transform_type), int(width), int(height), _ptr(data)))
        return data


class yiq:
    @staticmethod
    def rgb_to_yiq(rgb32f, ctx=None):
        """From<&Rgb32FImage> for YIQ32FImage (src/yiq.rs:177-186) -> (y, i, q) planes."""
        ctx = ctx or default_context()
        a = np.ascontiguousarray(rgb32f, dtype=np.float32)
        h, w = a.shape[:2]
        y, i, q = (np.empty((h, w), np.float32) for _ in range(3))
        check(lib.ssw_rgb32f_to_yiq(ctx.handle, _ptr(a), w, h, _ptr(y), _ptr(i), _ptr(q)))
        return y, i, q

    @staticmethod
    def yiq_to_rgb(y, i, q, ctx=None):
        """From<&YIQ32FImage> for Rgb32FImage (src/yiq.rs:187-197)."""
        ctx = ctx or default_context()
        y, i, q = (np.ascontiguousarray(p, dtype=np.float32) for p in (y, i, q))
        h, w = y.shape
        out = np.empty((h, w, 3), np.float32)
        check(lib.ssw_yiq_to_rgb32f(ctx.handle, _ptr(y), _ptr(i), _ptr(q), w, h, _ptr(out)))
        return out
