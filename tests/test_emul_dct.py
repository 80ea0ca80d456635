"""The CUDA DCT kernel *bodies* executed on the CPU (tests/emul: one std::thread per CUDA thread, a
std::barrier for __syncthreads) against the oracle.  This exercises the exact index arithmetic of
the device code in the GPU-less build container; it is test infrastructure, not a fallback."""
import ctypes

import numpy as np
import pytest

from conftest import ptr


def plan(emul, n):
    rad = (ctypes.c_int * 16)()
    tp, npad = ctypes.c_int(), ctypes.c_int()
    ns = emul.emul_plan(n, rad, ctypes.byref(tp), ctypes.byref(npad))
    return list(rad[:max(ns, 0)]), tp.value, npad.value, ns


@pytest.mark.parametrize('n', [1, 2, 3, 5, 7, 37, 64, 444, 640, 1080, 1920, 2160, 3840, 4096, 16384])
def test_plan_factorisation(emul, n):
    rad, tp, npad, ns = plan(emul, n)
    assert ns >= 0
    assert int(np.prod(rad)) == n if rad else n == 1
    assert tp % 32 == 0 and 32 <= tp <= 1024
    assert npad >= n


@pytest.mark.parametrize('r', [2, 3, 4, 5, 6, 8, 9, 10, 12, 15, 16])
def test_register_dfts(emul, r):
    rng = np.random.default_rng(r)
    x = (rng.standard_normal(r) + 1j * rng.standard_normal(r)).astype(np.complex64)
    buf = x.copy()
    emul.emul_dft(r, ptr(buf))
    ref = np.fft.fft(x.astype(np.complex128))
    assert np.abs(buf - ref).max() < 3e-6 * np.abs(ref).max()


@pytest.mark.parametrize('w,h,pr,pc,g', [
    (1, 1, 1, 1, 1), (4, 5, 1, 1, 1), (5, 4, 2, 1, 1), (9, 7, 1, 2, 1), (12, 37, 2, 3, 2),
    (30, 16, 3, 4, 1), (64, 48, 2, 4, 2), (74, 20, 1, 4, 1),
])
def test_dct2d_bodies_against_oracle(emul, so, w, h, pr, pc, g):
    rng = np.random.default_rng(w * 100 + h)
    a = rng.random((h, w)).astype(np.float32)
    for kind, okind in ((0, so.DCT2), (1, so.DCT2_ORTHO)):
        f = a.copy()
        assert emul.emul_dct2d(kind, w, h, ptr(f), pr, pc, g) == 0
        ref = so.dct2_2d(a, okind)
        assert np.abs(f - ref).max() <= 2e-6 * max(np.abs(ref).max(), 1e-30), (kind, np.abs(f - ref).max())
    c = so.dct2_2d(a, so.DCT2).astype(np.float32)
    b = c.copy()
    assert emul.emul_dct2d(2, w, h, ptr(b), pr, pc, g) == 0
    assert np.abs(b - a).max() < 2e-6


def test_fused_rgb8_forward_and_inverse_bodies(emul, so):
    w, h = 40, 24
    rgb = so.synth_frame(w, h, seed=3)
    plane = np.zeros((h, w), np.float32)
    assert emul.emul_rgb8_forward(w, h, ptr(rgb), ptr(plane), 2, 4, 1) == 0
    ref, _, _ = so.forward(rgb)
    assert np.abs(plane - ref).max() <= 2e-6 * np.abs(ref).max()
    out = np.zeros_like(rgb)
    assert emul.emul_rgb8_inverse(w, h, ptr(plane.copy()), ptr(rgb), ptr(out), 2, 4, 1) == 0
    # DCT3(DCT2(Y)) + the original I,Q gives back the original pixels (+-1 LSB from f32 rounding)
    assert np.abs(out.astype(int) - rgb.astype(int)).max() <= 1
