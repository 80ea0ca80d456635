#!/usr/bin/env python
"""Measured answer to "would a dense-matrix DCT on the tensor cores win?" (BASELINE.json north_star: "used only where ncu
shows it wins at the stated FP32 tolerance").

One row pass of the forward transform of a batch of 1080p luma planes, written as the GEMM  X[B*1080][1920] @ C[1920][1920]
(C = the scaled DCT-II matrix), run by cuBLAS through torch.matmul -- the best case for the dense form, a library kernel on
the 5th-generation tensor cores:
    tf32      one TF32 GEMM (10-bit mantissa inputs)                      -> accuracy far outside the 1e-5 coefficient bound
    3xtf32    X_hi C_hi + X_hi C_lo + X_lo C_hi (split operands)           -> FP32-grade accuracy, three GEMMs
    fp32      cuBLAS SGEMM on the FP32 pipes
against the FFT-form row pass of libssw (fwd_rows, same frames, RGB8 -> coefficients) timed with the same CUDA events.
Prints one JSON line; `python tools/tc_dct_probe.py > profiles/rN_tensor_core_dct.json` on the GPU box."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spread_spectrum_watermarking_b200 as wm  # noqa: E402
from spread_spectrum_watermarking_b200._lib import check, lib  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def tf32_round(t):
    """round-to-nearest to the 10-bit mantissa of TF32 (what the tensor core keeps of an f32 operand)"""
    i = t.view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def main():
    B, H, W = 64, 1080, 1920
    torch.backends.cuda.matmul.allow_tf32 = False
    n = np.arange(W)
    C64 = 2.0 * np.cos(np.pi * np.outer(2 * n + 1, n) / (2.0 * W))          # X @ C = 2 * sum_n x_n cos(pi k (2n+1) / 2N)
    C = torch.from_numpy(C64.astype(np.float32)).cuda()
    rng = np.random.default_rng(0)
    # natural-image-like luma rows: smooth + noise, in [0, 1]
    base = np.cumsum(rng.standard_normal((B * H, W)) * 0.01, axis=1)
    X64 = np.clip(0.5 + base + 0.02 * rng.standard_normal((B * H, W)), 0, 1)
    X = torch.from_numpy(X64.astype(np.float32)).cuda()
    ref = torch.from_numpy(X64[:H].astype(np.float32).astype(np.float64) @ C64).cuda()   # f64 reference of the first frame
    tol = lambda out: float(((out[:H].double() - ref).abs() / (1e-5 * ref.abs() + 1e-7 * ref.abs().max())).max())
    res = {}
    out = torch.empty_like(X)
    # fp32 SGEMM
    res['fp32_us'] = timed(lambda: torch.matmul(X, C, out=out))
    res['fp32_err_over_tol'] = tol(out)
    # single TF32 GEMM
    torch.backends.cuda.matmul.allow_tf32 = True
    res['tf32_us'] = timed(lambda: torch.matmul(X, C, out=out))
    res['tf32_err_over_tol'] = tol(out)
    # 3 x TF32 split
    Xh, Ch = tf32_round(X), tf32_round(C)
    Xl, Cl = X - Xh, C - Ch
    tmp = torch.empty_like(X)

    def three():
        torch.matmul(Xh, Ch, out=out)
        torch.matmul(Xh, Cl, out=tmp); out.add_(tmp)
        torch.matmul(Xl, Ch, out=tmp); out.add_(tmp)
    res['tf32x3_us'] = timed(three)
    res['tf32x3_err_over_tol'] = tol(out)
    torch.backends.cuda.matmul.allow_tf32 = False
    # the FFT-form row pass of libssw on the same number of pixels (RGB8 frames -> coefficient rows)
    ctx = wm.Context(0)
    frames = torch.empty((B, H, W, 3), dtype=torch.uint8, device='cuda')
    check(lib.ssw_synth_frame_rgb8_dev(ctx.handle, W, H, 3, 0, B, frames.data_ptr()))
    plane = torch.empty((B * H, W), dtype=torch.float32, device='cuda')
    ctx.synchronize()
    s = torch.cuda.ExternalStream(ctx.stream)

    def rows():
        check(lib.ssw_lines_forward_dev(ctx.handle, 0, frames.data_ptr(), W, B * H, plane.data_ptr()))
    with torch.cuda.stream(s):
        for _ in range(3):
            rows()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(20):
            rows()
        e1.record(s)
    ctx.synchronize()
    res['fft_rows_us'] = e0.elapsed_time(e1) / 20 * 1e3
    flops = 2.0 * B * H * W * W
    res.update(shape=[B * H, W, W], gemm_flop=flops, tf32_tflops=flops / res['tf32_us'] / 1e6, tf32x3_tflops_effective=flops / res['tf32x3_us'] / 1e6,
               note='err_over_tol = max |err| / (1e-5 |c| + 1e-7 max|C|): <= 1 passes the coefficient bound of the parity tests')
    ctx.close()
    print(json.dumps(res))


if __name__ == '__main__':
    main()
