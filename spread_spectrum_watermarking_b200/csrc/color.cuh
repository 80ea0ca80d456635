// RGB <-> YIQ arithmetic, bit-faithful to the reference (no FMA contraction, same association).
//
//   /root/reference/src/yiq.rs:157-159  RGB_TO_YIQ_MATRIX
//   /root/reference/src/yiq.rs:163-165  YIQ_TO_RGB_MATRIX
//   /root/reference/src/yiq.rs:131-147  Matrix3x3::product / product_clamp: (m0*v0 + m1*v1) + m2*v2
//   image 0.24.3 into_rgb32f (u8 as f32 / 255.0) and into_rgb8 (round(clamp(v,0,1)*255)), call sites
//   /root/reference/src/algorithm.rs:308,476 and /root/reference/tests/single_simple.rs:28
#pragma once
#include "dct_core.cuh"

#if defined(__CUDA_ARCH__)
#define SSW_FMUL(a, b) __fmul_rn((a), (b))
#define SSW_FADD(a, b) __fadd_rn((a), (b))
#define SSW_FDIV(a, b) __fdiv_rn((a), (b))
#else
// host emulation is compiled with -ffp-contract=off
#define SSW_FMUL(a, b) ((a) * (b))
#define SSW_FADD(a, b) ((a) + (b))
#define SSW_FDIV(a, b) ((a) / (b))
#endif

namespace ssw {

SSW_HD float u8_to_unit(unsigned v) { return SSW_FDIV((float)v, 255.0f); }

SSW_HD float mat3(float m0, float m1, float m2, float a, float b, float c) {
    return SSW_FADD(SSW_FADD(SSW_FMUL(m0, a), SSW_FMUL(m1, b)), SSW_FMUL(m2, c));
}

SSW_HD float rgb_to_y(float r, float g, float b) { return mat3(0.30f, 0.59f, 0.11f, r, g, b); }
SSW_HD float rgb_to_i(float r, float g, float b) { return mat3(0.60f, -0.28f, -0.32f, r, g, b); }
SSW_HD float rgb_to_q(float r, float g, float b) { return mat3(0.21f, -0.52f, 0.31f, r, g, b); }

SSW_HD float clamp01(float v) {
    // f32::clamp: NaN stays NaN
    if (v < 0.0f) return 0.0f;
    if (v > 1.0f) return 1.0f;
    return v;
}

SSW_HD void yiq_to_rgb(float y, float i, float q, float& r, float& g, float& b) {
    r = clamp01(mat3(1.0f, 0.948262f, 0.624013f, y, i, q));
    g = clamp01(mat3(1.0f, -0.276066f, -0.639810f, y, i, q));
    b = clamp01(mat3(1.0f, -1.105450f, 1.729860f, y, i, q));
}

SSW_HD unsigned unit_to_u8(float v) {
    // round half away from zero of clamp(v,0,1)*255; NaN -> 0 (Rust `as`-style saturating cast)
    float s = SSW_FMUL(clamp01(v), 255.0f);
    if (!(s == s)) return 0u;
    return (unsigned)(int)roundf(s);  // roundf == f32::round (half away from zero), s in [0,255]
}

}  // namespace ssw
