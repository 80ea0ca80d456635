"""ctypes binding of libssw.so (the C ABI declared in include/ssw.h).

There is no CPU fallback: if the CUDA library has not been built, importing this module raises.
Build it with `python -c "import __graft_entry__ as g; g.build()"` (or `make -C
spread_spectrum_watermarking_b200/csrc`).
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_size_t, c_uint8, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SSW_LIB') or os.path.join(HERE, 'csrc', 'libssw.so')   # SSW_LIB: another build of the same library (A/B, -DSSW_TRACE)

SSW_OK = 0
SSW_ERR_INVALID = -1
SSW_ERR_CUDA = -2
SSW_ERR_UNSUPPORTED = -3
SSW_ERR_STATE = -4


class SswError(RuntimeError):
    """Raised where the reference panics (or where CUDA fails). `.status` holds the ssw_status."""

    def __init__(self, status, message):
        super().__init__('libssw error %d: %s' % (status, message))
        self.status = status


class ssw_shard(ctypes.Structure):
    _fields_ = [('width', c_uint32), ('height', c_uint32), ('col0', c_uint32), ('ncols', c_uint32)]


class ssw_config(ctypes.Structure):
    _fields_ = [('method', c_int32), ('alpha', c_float), ('ordering', c_int32)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        'libssw.so is missing (%s). The sm_100a CUDA library must be built first: '
        'python -c "import __graft_entry__ as g; g.build()". There is no CPU fallback.' % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

_p = c_void_p
_pp = POINTER(c_void_p)
_f = POINTER(c_float)
_u8 = POINTER(c_uint8)
_cfg = POINTER(ssw_config)

# name -> (restype, argtypes); mirrors include/ssw.h one to one
SIGNATURES = {
    'ssw_last_error': (c_char_p, []),
    'ssw_version': (c_char_p, []),
    'ssw_ctx_create': (c_int, [c_int, _pp]),
    'ssw_ctx_create_on_stream': (c_int, [c_int, _p, _pp]),
    'ssw_ctx_destroy': (c_int, [_p]),
    'ssw_ctx_synchronize': (c_int, [_p]),
    'ssw_ctx_set_trace': (c_int, [_p, _p]),
    'ssw_ctx_stream': (c_void_p, [_p]),
    'ssw_ctx_launch_count': (c_uint64, [_p]),
    'ssw_ctx_profile_begin': (c_int, [_p]),
    'ssw_ctx_profile_end': (c_int, [_p, c_char_p, c_size_t]),
    'ssw_ctx_set_tiling': (c_int, [_p, c_int, c_int]),
    'ssw_host_alloc': (c_int, [c_size_t, _pp]),
    'ssw_host_free': (c_int, [_p]),
    'ssw_dct2_2d': (c_int, [_p, c_int, c_uint32, c_uint32, _p]),
    'ssw_dct2_2d_dev': (c_int, [_p, c_int, c_uint32, c_uint32, _p]),
    'ssw_rgb32f_to_yiq': (c_int, [_p, _p, c_uint32, c_uint32, _p, _p, _p]),
    'ssw_yiq_to_rgb32f': (c_int, [_p, _p, _p, _p, c_uint32, c_uint32, _p]),
    'ssw_writer_new_rgb8': (c_int, [_p, _p, c_uint32, c_uint32, _cfg, _pp]),
    'ssw_writer_new_rgb32f': (c_int, [_p, _p, c_uint32, c_uint32, _cfg, _pp]),
    'ssw_writer_new_rgb8_dev': (c_int, [_p, _p, c_uint32, c_uint32, _cfg, _pp]),
    'ssw_writer_embed': (c_int, [_p, POINTER(c_void_p), POINTER(c_size_t), c_size_t]),
    'ssw_writer_coefficients': (c_int, [_p, _p]),
    'ssw_writer_indices': (c_int, [_p, _p, c_size_t]),
    'ssw_writer_result_rgb8': (c_int, [_p, _p]),
    'ssw_writer_result_rgb32f': (c_int, [_p, _p]),
    'ssw_writer_result_rgb8_dev': (c_int, [_p, _p]),
    'ssw_writer_destroy': (c_int, [_p]),
    'ssw_reader_base_rgb8': (c_int, [_p, _p, c_uint32, c_uint32, _cfg, _pp]),
    'ssw_reader_base_rgb32f': (c_int, [_p, _p, c_uint32, c_uint32, _cfg, _pp]),
    'ssw_reader_derived_rgb8': (c_int, [_p, _p, c_uint32, c_uint32, _pp]),
    'ssw_reader_derived_rgb32f': (c_int, [_p, _p, c_uint32, c_uint32, _pp]),
    'ssw_reader_base_rgb8_dev': (c_int, [_p, _p, c_uint32, c_uint32, _cfg, _pp]),
    'ssw_reader_derived_rgb8_dev': (c_int, [_p, _p, c_uint32, c_uint32, _pp]),
    'ssw_reader_extract': (c_int, [_p, _p, _p, c_size_t]),
    'ssw_reader_extract_dev': (c_int, [_p, _p, _p, c_size_t]),
    'ssw_reader_coefficients': (c_int, [_p, _p]),
    'ssw_reader_indices': (c_int, [_p, _p, c_size_t]),
    'ssw_reader_destroy': (c_int, [_p]),
    'ssw_similarity': (c_int, [_p, _p, _p, c_size_t, _f]),
    'ssw_bank_create': (c_int, [_p, _p, c_size_t, c_size_t, _pp]),
    'ssw_bank_create_normal': (c_int, [_p, c_uint64, c_size_t, c_size_t, _pp]),
    'ssw_bank_row': (c_int, [_p, c_size_t, _p]),
    'ssw_bank_similarity': (c_int, [_p, _p, c_size_t, _p]),
    'ssw_bank_similarity_dev': (c_int, [_p, _p, c_size_t, _p]),
    'ssw_bank_destroy': (c_int, [_p]),
    'ssw_mark_generate_normal': (c_int, [_p, c_uint64, c_size_t, _p]),
    'ssw_embed_batch_rgb8_dev': (c_int, [_p, _p, c_uint32, c_uint32, c_uint32, _cfg, _p, c_size_t, _p]),
    'ssw_extract_batch_rgb8_dev': (c_int, [_p, _p, _p, c_uint32, c_uint32, c_uint32, _cfg, c_size_t, _p, _p, _p]),
    'ssw_embed_batch_rgb8': (c_int, [_p, _p, c_uint32, c_uint32, c_uint32, _cfg, _p, c_size_t, _p]),
    'ssw_extract_batch_rgb8': (c_int, [_p, _p, _p, c_uint32, c_uint32, c_uint32, _cfg, c_size_t, _p, _p, _p]),
    'ssw_embed_batch_rgb8_async': (c_int, [_p, _p, c_uint32, c_uint32, c_uint32, _cfg, _p, c_size_t, _p]),
    'ssw_extract_batch_rgb8_async': (c_int, [_p, _p, _p, c_uint32, c_uint32, c_uint32, _cfg, c_size_t, _p, _p, _p]),
    'ssw_ctx_marker': (c_int, [_p, POINTER(c_uint64)]),
    'ssw_ctx_wait_marker': (c_int, [_p, c_uint64]),
    'ssw_ctx_last_topk_fallbacks': (c_int, [_p]),
    'ssw_selftest_pack_u8': (c_int, [_p, POINTER(c_uint64)]),
    'ssw_synth_frame_rgb8_dev': (c_int, [_p, c_uint32, c_uint32, c_uint64, c_uint32, c_uint32, _p]),
    'ssw_synth_rows_rgb8_dev': (c_int, [_p, c_uint32, c_uint64, c_uint32, c_uint32, c_uint32, _p]),
    'ssw_stage_forward_rgb8_dev': (c_int, [_p, _p, c_uint32, c_uint32, c_uint32, _p]),
    'ssw_stage_topk_dev': (c_int, [_p, _p, c_uint32, c_uint32, c_uint32, c_int, c_size_t, _p]),
    'ssw_stage_inverse_rgb8_dev': (c_int, [_p, _p, _p, c_uint32, c_uint32, c_uint32, _p]),
    'ssw_lines_forward_dev': (c_int, [_p, c_int, _p, c_uint32, c_uint32, _p]),
    'ssw_lines_forward_seg_dev': (c_int, [_p, _p, c_uint32, c_uint32, c_uint32, c_uint32, c_uint32, _p]),
    'ssw_lines_inverse_seg_dev': (c_int, [_p, _p, c_uint32, c_uint32, c_uint32, c_uint32, c_uint32, c_float, c_int, _p, c_int, _p]),
    'ssw_lines_inverse_dev': (c_int, [_p, _p, c_uint32, c_uint32, c_float, c_int, _p, c_int, _p]),
    'ssw_transpose_dev': (c_int, [_p, _p, c_uint32, c_uint32, ctypes.c_int64, ctypes.c_int64, _p, ctypes.c_int64,
                                  ctypes.c_int64, c_uint32]),
    'ssw_shard_topk_bin_dev': (c_int, [_p, _p, POINTER(ssw_shard), c_int, c_size_t, _p]),
    'ssw_shard_topk_collect_dev': (c_int, [_p, _p, POINTER(ssw_shard), c_int, _p, _p, _p]),
    'ssw_shard_topk_merge_dev': (c_int, [_p, _p, _p, c_uint32, c_size_t, _p, _p]),
    'ssw_shard_embed_dev': (c_int, [_p, _p, POINTER(ssw_shard), _p, c_size_t, _p, c_size_t, c_size_t, _p, _cfg]),
    'ssw_shard_extract_dev': (c_int, [_p, _p, _p, POINTER(ssw_shard), _p, c_size_t, _cfg, _p]),
    'ssw_sharded_unique_id': (c_int, [_p]),
    'ssw_sharded_create': (c_int, [_p, _p, c_int, c_int, c_uint32, c_uint32, _pp]),
    'ssw_sharded_destroy': (c_int, [_p]),
    'ssw_sharded_embed_rgb8_dev': (c_int, [_p, _p, _cfg, _p, c_size_t, _p]),
    'ssw_sharded_extract_rgb8_dev': (c_int, [_p, _p, _p, _cfg, c_size_t, _p]),
    'ssw_sharded_indices': (c_int, [_p, _p, c_size_t]),
    'ssw_sharded_coefficients': (c_int, [_p, c_int, _p]),
    'ssw_sharded_overflow': (c_int, [_p, POINTER(c_int)]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def last_error():
    msg = lib.ssw_last_error()
    return msg.decode('utf-8', 'replace') if msg else ''


def check(status):
    if status != SSW_OK:
        raise SswError(status, last_error())
    return status
