"""Shared fixtures.  `-m "not gpu"` covers the oracle, the host logic and the C-ABI surface; `-m gpu`
are the parity tests proper (CUDA path through the C ABI vs. the oracle).  Nothing here reads
/root/reference: the reference's fixtures travel as tests/golden/*.npz (see oracle/make_golden.py)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (sm_100a); run on the B200 box')


def _make(path, target):
    if not os.path.exists(os.path.join(path, target)):
        subprocess.check_call(['make', '-C', path, target])


@pytest.fixture(scope='session')
def so():
    """the numpy oracle (checker)"""
    import ssw_oracle
    return ssw_oracle


@pytest.fixture(scope='session')
def coracle():
    """the C restatement (checker + CPU baseline), oracle/liboracle.so"""
    path = os.path.join(ROOT, 'oracle')
    _make(path, 'liboracle.so')
    lib = ctypes.CDLL(os.path.join(path, 'liboracle.so'))
    lib.oracle_similarity.restype = ctypes.c_float
    return lib


@pytest.fixture(scope='session')
def emul():
    """CPU emulation of the CUDA kernel bodies (tests/emul)"""
    path = os.path.join(ROOT, 'tests', 'emul')
    _make(path, 'libssw_emul.so')
    return ctypes.CDLL(os.path.join(path, 'libssw_emul.so'))


@pytest.fixture(scope='session')
def golden():
    cat = np.load(os.path.join(GOLDEN, 'cat_rgb8.npz'))['rgb']
    gold = np.load(os.path.join(GOLDEN, 'watermarked_with_1.npz'))['rgb']
    marks = dict(np.load(os.path.join(GOLDEN, 'marks.npz')))
    ora = dict(np.load(os.path.join(GOLDEN, 'cat_oracle.npz')))
    return {'cat': cat, 'gold': gold, 'marks': marks, 'oracle': ora}


@pytest.fixture(scope='session')
def wm():
    """the product package (ctypes over libssw.so); importing it needs the built library"""
    import spread_spectrum_watermarking_b200 as m
    return m


@pytest.fixture(scope='session')
def ctx(wm):
    """one CUDA context for the gpu-marked tests; fails loudly if there is no device"""
    c = wm.Context(0)
    yield c
    c.close()


def ptr(a):
    # data_as keeps a reference to the array, so ptr(x.copy()) stays valid for the duration of the call it is passed to
    # (a bare c_void_p of the address would let the temporary be freed -- and its memory reused -- before the callee reads it)
    return a.ctypes.data_as(ctypes.c_void_p)
