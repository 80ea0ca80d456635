"""The drop-in boundary without a GPU: libssw.so loads, exports exactly what include/ssw.h declares,
the ctypes table mirrors the header, and the product path fails loudly when no CUDA device exists
(no CPU fallback).  No compute calls here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, ptr

HEADER = os.path.join(ROOT, 'include', 'ssw.h')


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ssw_[a-z0-9_]+)\s*\(', src)))


def test_header_compiles_as_c():
    """the boundary is plain C: no C++/torch types in the signatures"""
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-fsyntax-only', '-x', 'c', HEADER])


def test_library_exports_every_declared_symbol(wm):
    names = header_functions()
    assert len(names) > 40
    lib = ctypes.CDLL(wm._lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_table_mirrors_header(wm):
    assert sorted(wm._lib.SIGNATURES) == header_functions()


def test_exported_symbols_are_declared(wm):
    out = subprocess.check_output(['nm', '-D', '--defined-only', wm._lib.LIB_PATH], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if ' T ' in l and l.split()[-1].startswith('ssw_'))
    assert exported == header_functions()


def test_library_is_sm100a_only(wm):
    out = subprocess.run(['cuobjdump', '-lelf', wm._lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    archs = set(re.findall(r'sm_\d+a?', out.stdout))
    assert archs == {'sm_100a'}, archs


def test_version_and_error_strings(wm):
    assert b'sm_100a' in wm.lib.ssw_version()
    assert isinstance(wm._lib.last_error(), str)


def _no_device():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_device(), reason='a CUDA device is present')
def test_fails_loudly_without_a_gpu(wm):
    """no CPU fallback: every way into the product path raises when there is no device"""
    with pytest.raises(wm.SswError) as e:
        wm.Context(0)
    assert e.value.status == wm._lib.SSW_ERR_CUDA
    assert 'no CPU path' in str(e.value) or 'CUDA' in str(e.value)
    img = np.zeros((8, 8, 3), np.uint8)
    wm.__dict__['_default_ctx'] = None
    with pytest.raises(wm.SswError):
        wm.Writer.new(img)
    with pytest.raises(wm.SswError):
        wm.MarkBuf.generate_normal(10)


def test_null_arguments_are_rejected_not_crashed(wm):
    h = ctypes.c_void_p()
    cfg = wm._lib.ssw_config(2, 0.1, 0)
    assert wm.lib.ssw_ctx_create(0, None) == wm._lib.SSW_ERR_INVALID
    assert wm.lib.ssw_writer_new_rgb8(None, None, 4, 4, ctypes.byref(cfg), ctypes.byref(h)) == wm._lib.SSW_ERR_INVALID
    assert wm.lib.ssw_reader_extract(None, None, None, 10) == wm._lib.SSW_ERR_INVALID
    assert wm.lib.ssw_similarity(None, None, None, 0, None) == wm._lib.SSW_ERR_INVALID
    assert wm.lib.ssw_ctx_destroy(None) == 0 and wm.lib.ssw_writer_destroy(None) == 0
    assert 'NULL' in wm._lib.last_error() or wm._lib.last_error()


def test_host_side_config_mirror(wm):
    """WriteConfig/ReadConfig defaults (src/algorithm.rs:105-112,133-140) and Custom rejection"""
    wc, rc = wm.WriteConfig.default(), wm.ReadConfig.default()
    assert (wc.insertion.option, wc.insertion.alpha, wc.ordering) == (2, 0.1, wm.OrderingMethod.Energy)
    assert (rc.extraction.option, rc.extraction.alpha, rc.ordering) == (2, 0.1, wm.OrderingMethod.Energy)
    c = wm._c_config(wm.Insertion.Option3(0.25), wm.OrderingMethod.Legacy)
    assert (c.method, c.alpha, c.ordering) == (3, 0.25, 2)
    with pytest.raises(wm.SswError) as e:
        wm._c_config(wm.Insertion.Custom(lambda i, o, w: o), wm.OrderingMethod.Energy)
    assert e.value.status == wm._lib.SSW_ERR_UNSUPPORTED
    with pytest.raises(wm.SswError):
        wm._c_config(wm.Insertion.Option2(0.1), wm.OrderingMethod.Custom(lambda a, b: 0))


def test_host_side_image_and_mark_adapters(wm):
    """DynamicImage stand-ins: luma / RGBA / f64 inputs become contiguous RGB8 or RGB32F"""
    assert wm._as_rgb(np.zeros((4, 5), np.uint8)).shape == (4, 5, 3)
    assert wm._as_rgb(np.zeros((4, 5, 4), np.uint8)).shape == (4, 5, 3)
    assert wm._as_rgb(np.zeros((4, 5, 3), np.float64)).dtype == np.float32
    with pytest.raises(wm.SswError):
        wm._as_rgb(np.zeros((4, 5, 2), np.uint8))
    with pytest.raises(wm.SswError):
        wm._as_rgb(np.zeros((4, 5, 3), np.int32))
    m = wm.MarkBuf.from_([1.0, 2.0, 3.0])
    assert len(m) == 3 and m.data().dtype == np.float32
    m.set_data(np.arange(5))
    assert m.data().tolist() == [0, 1, 2, 3, 4]
    assert wm._mark_data([0.5, 1.5]).tolist() == [0.5, 1.5]       # impl Mark for AsRef<[f32]>
    assert wm._mark_data(m) is m.data() or (wm._mark_data(m) == m.data()).all()
    s = wm.Similarity(6.5)
    assert s.exceeds_sigma(6.0) and not s.exceeds_sigma(6.5)      # strict >, src/algorithm.rs:677-679


def test_every_pdl_launched_kernel_waits_before_touching_memory():
    """launch_pdl (programmatic stream serialization) lets a grid start while its predecessor drains, so every kernel
    that goes through it MUST begin with pdl_enter() / pdl_wait() (csrc/pdl.cuh).  Static check of the sources:
    each kernel named at a launch_pdl call site has the wait as its first statement."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'spread_spectrum_watermarking_b200', 'csrc')
    src = {f: open(os.path.join(root, f)).read() for f in os.listdir(root) if f.endswith(('.cu', '.cuh', '.h'))}
    api = src['ssw_api.cu']
    launched = set(re.findall(r'launch_pdl\(c,\s*([A-Za-z_0-9]+)\s*,', api))
    launched.discard('kernel')   # the template wrappers fast_kernel / fast_kernel_pf, checked below
    assert len(launched) >= 10, launched
    all_src = '\n'.join(src.values())
    for k in sorted(launched):
        m = re.search(r'\b' + k + r'\s*\([^{;]*\)\s*\{\s*(?://[^\n]*\n\s*)*(\w+)\(\);', all_src)
        assert m, 'definition of %s not found' % k
        assert m.group(1) in ('pdl_enter', 'pdl_wait'), '%s starts with %s' % (k, m.group(1))
    for wrapper in ('fast_kernel', 'fast_kernel_pf'):
        body = src['dct_fast.cuh'].split('fast::' + wrapper)[0] if False else src['dct_fast.cuh']
        i = body.index(wrapper + '(const __grid_constant__ FastArgs a)')
        head = body[i:i + 600]
        assert 'pdl_wait();' in head and head.index('pdl_wait();') < head.index('phase<'), wrapper
    # and nothing else is launched with the attribute: the generic fallback kernels keep <<< >>>
    assert 'cudaLaunchAttributeProgrammaticStreamSerialization' not in all_src.replace(api, '')
    assert api.count('cudaLaunchAttributeProgrammaticStreamSerialization') == 1
