// Compile-time planned line kernels of the full-frame separable DCT ("fast path").
//
// Same arithmetic as dct_kernels.cuh (two real lines packed into one complex Stockham FFT, Makhoul
// reordering, twiddle post/pre pass -- replaces /root/reference/src/dct2d.rs:129-206 and the rustdct
// calls at :141-145,181-185, with the colour conversion of /root/reference/src/yiq.rs:177-197 fused
// into the row passes), but the radix plan, the thread shape and every shared-memory offset are
// template constants:
//   * no padding for lengths with an odd first radix (its stride-R first-stage stores are conflict
//     free), `a + a/16` padding for powers of two;
//   * butterfly operands are addressed as base + immediate, twiddles as one coalesced __ldg each;
//   * 128-bit global accesses: 4 pixels (12 B of RGB8, or a float4 of the plane) per thread and row.
// The generic kernels stay as the fallback for every other length (e.g. 444 = 4*3*37).
//
// Every kernel is written as a sequence of barrier-separated PHASES over a per-thread register
// state, `K::phase<PH>(args, smem, tile, tid, state)`.  The __global__ wrapper runs the phases with
// __syncthreads() between them; tests/emul runs the very same phase functions on the CPU, one
// thread after the other, so the index arithmetic is verified without a GPU.
#pragma once
#include "dct_kernels.cuh"
#if defined(__CUDACC__)
#include "pdl.cuh"
#endif

#if defined(__CUDA_ARCH__)
#define SSW_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define SSW_FADD_RZ(a, b) __fadd_rz((a), (b))
#define SSW_PRMT(a, b, sel) __byte_perm((a), (b), (sel))
#define SSW_F2U(f) __float_as_uint(f)
#define SSW_U2F(u) __uint_as_float(u)
#else
#include <cmath>
#include <cstring>
#define SSW_FMA(a, b, c) std::fmaf((a), (b), (c))
// host stand-ins (tests/emul): a + b is exact in double for the operands used here; truncate to f32
inline float ssw_host_fadd_rz(float a, float b) {
    const double d = (double)a + (double)b;
    float f = (float)d;
    if (std::fabs((double)f) > std::fabs(d)) f = std::nextafterf(f, 0.0f);
    return f;
}
inline unsigned ssw_host_prmt(unsigned a, unsigned b, unsigned sel) {   // PRMT, default mode, selectors 0..7
    const unsigned long long v = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 255u) << (8 * i);
    return r;
}
inline unsigned ssw_host_f2u(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float ssw_host_u2f(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
#define SSW_FADD_RZ(a, b) ssw_host_fadd_rz((a), (b))
#define SSW_PRMT(a, b, sel) ssw_host_prmt((a), (b), (sel))
#define SSW_F2U(f) ssw_host_f2u(f)
#define SSW_U2F(u) ssw_host_u2f(u)
#endif

namespace ssw {
namespace fast {

struct FastArgs {
    int w, h;            // frame size
    const void* src;     // row_fwd: pixels / plane; row_inv (RGB8): original pixels for I,Q
    float* plane;        // coefficient plane [h][w]
    void* dst;           // row_inv destination
    long long src_stride, plane_stride, dst_stride;  // per-image strides in pixels
    int tiles_per_image;
    int total_tiles, tiles_per_cta;   // prefetching kernels: a CTA works through tiles_per_cta consecutive tiles
    float scale0, scalen;  // forward: factors for k == 0 / k > 0; inverse: scale0 = output scale
    const cplx* tw;        // stage twiddles of the plan, layout per stage [r-1][k]
    const cplx* t4;        // exp(-i*pi*k/(2N)), k < N
    // segmented f32 source lines (sharded frames: a line is the concatenation of the blocks received from
    // the all-to-all, buffer layout [chunks][ranks][lines][seg_len]); seg_shift < 0: contiguous lines.
    // sample m of line l lives at (((c*ranks + g)*lines + l) << seg_shift) + (m & (seg_len-1)) with
    // s = m >> seg_shift, g = s >> chunk_shift, c = s & (chunks-1)      (seg_len and chunks: powers of two)
    int seg_shift, chunk_shift, seg_ranks, seg_lines;
    int pdl_late;   // 0: release the dependent grid at the first instruction; 1: before the last phase (see pdl.cuh)
    int dbg_skip;   // tuning builds only (-DSSW_TUNE): 1 = no global loads, 2 = no global stores, 4 = no FFT stages
    float neg_zero; // -0.0f, deliberately a RUN-TIME value: fma.f32x2(m, c, -0) is the separately rounded product m*c that
                    // the reference's un-fused colour arithmetic needs, in a form ptxas cannot contract with the following add
                    // (it does contract mul.rn.f32x2 + add.rn.f32x2 into FFMA2, even under --fmad=false)
};

// position of 4 consecutive samples [m, m+4) of line `line` inside a segmented source
SSW_HD long long seg_index(const FastArgs& a, int line, int m) {
    const int s = m >> a.seg_shift, off = m & ((1 << a.seg_shift) - 1);
    const int g = s >> a.chunk_shift, c = s & ((1 << a.chunk_shift) - 1);
    return (((long long)(c * a.seg_ranks + g) * a.seg_lines + line) << a.seg_shift) + off;
}

// ------------------------------------------------------------------------------------------------
// plan: N = R0*R1*R2*R3 (unused trailing radices = 1), T threads cooperate on one line pair
// ------------------------------------------------------------------------------------------------
template <int N_, int T_, int R0_, int R1_, int R2_ = 1, int R3_ = 1>
struct Plan {
    static constexpr int N = N_, T = T_;
    static constexpr int NST = R3_ > 1 ? 4 : (R2_ > 1 ? 3 : 2);
    static_assert(R0_ * R1_ * R2_ * R3_ == N_, "radices must multiply to N");
    static_assert(N_ % 4 == 0, "fast path needs N % 4 == 0 (4 pixels per thread)");
    static_assert(T_ % 32 == 0, "whole warps per team");
    static constexpr int radix(int s) { return s == 0 ? R0_ : s == 1 ? R1_ : s == 2 ? R2_ : R3_; }
    static constexpr int ns(int s) { int p = 1; for (int i = 0; i < s; ++i) p *= radix(i); return p; }
    static constexpr int tw_off(int s) { int o = 0; for (int i = 1; i < s; ++i) o += (radix(i) - 1) * ns(i); return o; }
    static constexpr int TW_TOTAL = tw_off(NST);
    static constexpr int KEY = ((N_ * 31 + R0_) * 31 + R1_) * 31 + R2_;  // twiddle-cache key (plans of one N may differ in radices)
    static constexpr bool PAD = (R0_ % 2) == 0;
    static SSW_HD int idx(int a) { return PAD ? a + (a >> 4) : a; }
    static constexpr int LINE = PAD ? N_ + (N_ >> 4) : N_;
    // pitch between the line pairs of one CTA (float2 units): even (16-byte alignment), == 4 mod 16
    static constexpr int PITCH = LINE + ((4 - LINE % 16) + 16) % 16;
};

template <class P, int S>
struct StageInfo {
    static constexpr int R = P::radix(S), NS = P::ns(S), NB = P::N / R;
    // blocks of 15 butterflies: give each block its own half-warp so no access straddles two blocks
    static constexpr bool MAP16 = (NS == 15) && (NB > 15);
    static constexpr int SLOTS = MAP16 ? (NB / 15) * 16 : NB;
    static constexpr int ITER = (SLOTS + P::T - 1) / P::T;
    static constexpr bool GUARD = (ITER * P::T != SLOTS);
    static constexpr int TW = P::tw_off(S);
};

template <class P, int S = 0>
struct MaxRegs {
    static constexpr int here = (S < P::NST) ? StageInfo<P, S>::ITER * StageInfo<P, S>::R : 0;
    static constexpr int rest = MaxRegs<P, S + 1>::value;
    static constexpr int value = here > rest ? here : rest;
};
template <class P>
struct MaxRegs<P, 4> { static constexpr int value = 0; };

template <class P, int S>
SSW_HD bool stage_map(int t, int it, int& j, int& k) {
    using I = StageInfo<P, S>;
    const int slot = t + it * P::T;
    if constexpr (I::MAP16) {
        const int blk = slot >> 4;
        k = slot & 15;
        j = blk * 15 + k;
        return (k < 15) && (!I::GUARD || slot < I::SLOTS);
    } else {
        j = slot;
        k = (I::NS == 1) ? 0 : ((I::NS >= I::NB) ? j : j % I::NS);
        return !I::GUARD || slot < I::SLOTS;
    }
}

// padded position of (base + off): when the constant offset is a multiple of 16 (or the layout is not padded)
// the padding of the sum splits into idx(base) + a compile-time constant, so butterfly operands stay
// addressed as base + immediate; otherwise (first power-of-two stage: off = r < 16 on a base that is a
// multiple of 16) the low bits cannot carry either.
template <class P, int OFF, bool BASE16>
SSW_HD int idx_off(int base_idx, int base) {
    if constexpr (!P::PAD) return base_idx + OFF;
    else if constexpr (OFF % 16 == 0) return base_idx + OFF + OFF / 16;
    else if constexpr (BASE16 && OFF < 16) return base_idx + OFF;
    else return P::idx(base + OFF);
}

template <class P, int S>
SSW_HD void stage_load(const cplx* s, int t, cplx* v) {
    using I = StageInfo<P, S>;
#pragma unroll
    for (int it = 0; it < I::ITER; ++it) {
        int j, k;
        if (stage_map<P, S>(t, it, j, k)) {
            const int bj = P::idx(j);
            static_for<I::R>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                v[it * I::R + r] = s[idx_off<P, r * I::NB, false>(bj, j)];
            });
        }
    }
}

// TWS: the twiddle table lives in shared memory (plain loads; __ldg is for global memory only)
template <class P, int S, bool TWS = false>
SSW_HD void stage_store(cplx* s, const cplx* tw, int t, cplx* v) {
    using I = StageInfo<P, S>;
#pragma unroll
    for (int it = 0; it < I::ITER; ++it) {
        int j, k;
        if (stage_map<P, S>(t, it, j, k)) {
            cplx* x = v + it * I::R;
            if constexpr (I::NS > 1) {
                const cplx* twk = tw + I::TW + k;
#pragma unroll
                for (int r = 1; r < I::R; ++r) x[r] = cmul(x[r], TWS ? twk[(r - 1) * I::NS] : SSW_LDG(twk + (r - 1) * I::NS));
            }
            Dft<I::R>::run(x);
            const int j0 = (j - k) * I::R + k;
            const int b0 = P::idx(j0);
            // first stage (NS == 1): j0 = j*R; with R == 16 it is a multiple of 16 and r < 16 cannot carry
            static_for<I::R>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                s[idx_off<P, r * I::NS, (I::NS == 1 && I::R == 16)>(b0, j0)] = x[r];
            });
        }
    }
}

// phases 1 .. 2*NST of every kernel: the FFT of the team's line pair
template <class P, int PH, bool TWS = false>
SSW_HD void fft_phase(cplx* s, const cplx* tw, int t, cplx* v) {
    constexpr int S = (PH - 1) / 2;
    if constexpr (((PH - 1) & 1) == 0) stage_load<P, S>(s, t, v);
    else stage_store<P, S, TWS>(s, tw, t, v);
}

// ------------------------------------------------------------------------------------------------
// colour helpers (bit-identical to color.cuh; the division by 255 is replaced by an exact
// two-term constant and one fma -- verified for all 256 inputs by the CPU test-suite)
// ------------------------------------------------------------------------------------------------
#ifndef SSW_U8_MAGIC
#define SSW_U8_MAGIC 1
#endif
// x (an integer 0..255 held exactly in a float) -> x/255, correctly rounded
SSW_HD float unit_of(float x) {
    // 1/255 = c_hi + c_lo to ~2^-50: fma(x, c_hi, fl(x*c_lo)) is the correctly rounded x/255 for x = 0..255
    const float c_hi = 0.0039215688593685626983642578125f;  // fl32(1/255)
    const float c_lo = -2.31917579870781060424633324146270751953125e-10f;  // fl32(1/255 - c_hi)
    return SSW_FMA(x, c_hi, SSW_FMUL(x, c_lo));
}
SSW_HD float u8_unit(unsigned v) { return unit_of((float)v); }

// byte j of w -> float, exactly.  Device: one PRMT builds the bits of 2^23 + b (0x4B0000bb) and one FADD removes
// the 2^23 -- both on the full-rate ALU / FMA pipes, instead of an I2F.U8 on the conversion unit.
template <int J>
SSW_HD float byte_to_float(unsigned w) {
#if SSW_U8_MAGIC
    return SSW_FADD(SSW_U2F(SSW_PRMT(w, 0x4B000000u, 0x7540u + J)), -8388608.0f);
#else
    return (float)((w >> (8 * J)) & 255u);
#endif
}

// 12 bytes (4 RGB8 pixels) -> the 12 channel values as u8/255
SSW_HD void unpack4_unit(unsigned w0, unsigned w1, unsigned w2, float* c) {
    c[0] = unit_of(byte_to_float<0>(w0)); c[1] = unit_of(byte_to_float<1>(w0));
    c[2] = unit_of(byte_to_float<2>(w0)); c[3] = unit_of(byte_to_float<3>(w0));
    c[4] = unit_of(byte_to_float<0>(w1)); c[5] = unit_of(byte_to_float<1>(w1));
    c[6] = unit_of(byte_to_float<2>(w1)); c[7] = unit_of(byte_to_float<3>(w1));
    c[8] = unit_of(byte_to_float<0>(w2)); c[9] = unit_of(byte_to_float<1>(w2));
    c[10] = unit_of(byte_to_float<2>(w2)); c[11] = unit_of(byte_to_float<3>(w2));
}

struct f4 { float a, b, c, d; };  // 16-byte vector for host + device
// streaming 128-bit / 32-bit global loads: read-only path, do not allocate in L1 (the twiddle tables live there)
SSW_HD f4 ld4(const float* p) {
#if defined(__CUDA_ARCH__)
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return f4{v.x, v.y, v.z, v.w};
#else
    return f4{p[0], p[1], p[2], p[3]};
#endif
}
SSW_HD unsigned ldw(const unsigned* p) {
#if defined(__CUDA_ARCH__)
    unsigned v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}
SSW_HD void st4(float* p, float a, float b, float c, float d) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
#else
    p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#endif
}

// round(clamp(v,0,1)*255) with ties away from zero, NaN -> 0: bit-identical to unit_to_u8(clamp01(v)).
// s + 0.5 is exact for every f32 s in [0,255] except the largest float below 0.5, where round-to-
// nearest would give 1.0; adding with round-toward-zero fixes exactly that case.
SSW_HD unsigned unit_to_u8_fast(float v) {
#if defined(__CUDA_ARCH__)
    const float s = __fmul_rn(fminf(fmaxf(v, 0.0f), 1.0f), 255.0f);
    return (unsigned)__float2int_rz(__fadd_rz(s, 0.5f));
#else
    float c = v;
    if (!(c == c)) return 0u;
    c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
    const float s = c * 255.0f;
    const double d = (double)s + 0.5;  // exact
    return (unsigned)(long long)d;     // truncation == round-toward-zero add followed by F2I.TRUNC
#endif
}

// four unit values -> four RGB8 bytes packed into one word (byte i = unit_to_u8_fast(o[i])).
// Device: floor(t) of t = RZ(s + 0.5) in [0, 256) is read off the low mantissa byte of RZ(t + 2^23), and three
// PRMTs gather the four bytes -- no F2I on the conversion unit, no shift/or chain.
SSW_HD unsigned pack_u8x4(const float* o) {
#if SSW_U8_MAGIC
    unsigned m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#if defined(__CUDA_ARCH__)
        const float c = fminf(fmaxf(o[i], 0.0f), 1.0f);   // NaN -> 0
#else
        const float c = !(o[i] == o[i]) ? 0.0f : (o[i] < 0.0f ? 0.0f : (o[i] > 1.0f ? 1.0f : o[i]));
#endif
        m[i] = SSW_F2U(SSW_FADD_RZ(SSW_FADD_RZ(SSW_FMUL(c, 255.0f), 0.5f), 8388608.0f));
    }
    return SSW_PRMT(SSW_PRMT(m[0], m[1], 0x0040u), SSW_PRMT(m[2], m[3], 0x0040u), 0x5410u);
#else
    return unit_to_u8_fast(o[0]) | (unit_to_u8_fast(o[1]) << 8) | (unit_to_u8_fast(o[2]) << 16) | (unit_to_u8_fast(o[3]) << 24);
#endif
}

// base pointer of image `img` inside a batch of pixel type TYPE (stride in pixels)
template <int TYPE>
SSW_HD const void* image_base(const void* p, long long img, long long stride) {
    if constexpr (TYPE == PIX_RGB8) return (const unsigned char*)p + 3 * img * stride;
    else if constexpr (TYPE == PIX_RGB32F) return (const float*)p + 3 * img * stride;
    else return (const float*)p + img * stride;
}

// 4 pixels of one row -> luma
template <int SRC>
SSW_HD void load_luma4(const void* src, long long pix4 /* index of the first of 4 pixels */, float* y) {
    if constexpr (SRC == PIX_RGB8) {
        const unsigned* p = (const unsigned*)((const unsigned char*)src + 3 * pix4);
        float c[12];
        unpack4_unit(ldw(p), ldw(p + 1), ldw(p + 2), c);
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = rgb_to_y(c[3 * i], c[3 * i + 1], c[3 * i + 2]);
    } else if constexpr (SRC == PIX_RGB32F) {
        const float* p = (const float*)src + 3 * pix4;
        const f4 v0 = ld4(p), v1 = ld4(p + 4), v2 = ld4(p + 8);
        y[0] = rgb_to_y(v0.a, v0.b, v0.c); y[1] = rgb_to_y(v0.d, v1.a, v1.b);
        y[2] = rgb_to_y(v1.c, v1.d, v2.a); y[3] = rgb_to_y(v2.b, v2.c, v2.d);
    } else {
        const f4 v = ld4((const float*)src + pix4);
        y[0] = v.a; y[1] = v.b; y[2] = v.c; y[3] = v.d;
    }
}

// 4 pixels of one row: new luma y[4] (+ chroma of the original pixels) -> destination
template <int DST, int SRC>
SSW_HD void store_pix4(const void* src, void* dst, long long pix4, const float* y) {
    if constexpr (DST == PIX_PLANE) {
        st4((float*)dst + pix4, y[0], y[1], y[2], y[3]);
    } else {
        float c[12];
        if constexpr (SRC == PIX_RGB8) {
            const unsigned* p = (const unsigned*)((const unsigned char*)src + 3 * pix4);
            unpack4_unit(ldw(p), ldw(p + 1), ldw(p + 2), c);
        } else {
            const float* p = (const float*)src + 3 * pix4;
            const f4 v0 = ld4(p), v1 = ld4(p + 4), v2 = ld4(p + 8);
            c[0] = v0.a; c[1] = v0.b; c[2] = v0.c; c[3] = v0.d; c[4] = v1.a; c[5] = v1.b;
            c[6] = v1.c; c[7] = v1.d; c[8] = v2.a; c[9] = v2.b; c[10] = v2.c; c[11] = v2.d;
        }
        float o[12];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float ci = rgb_to_i(c[3 * i], c[3 * i + 1], c[3 * i + 2]);
            const float cq = rgb_to_q(c[3 * i], c[3 * i + 1], c[3 * i + 2]);
            // src/yiq.rs:163-165 rows of YIQ_TO_RGB_MATRIX, (m0*y + m1*i) + m2*q, m0 == 1
            o[3 * i] = SSW_FADD(SSW_FADD(y[i], SSW_FMUL(0.948262f, ci)), SSW_FMUL(0.624013f, cq));
            o[3 * i + 1] = SSW_FADD(SSW_FADD(y[i], SSW_FMUL(-0.276066f, ci)), SSW_FMUL(-0.639810f, cq));
            o[3 * i + 2] = SSW_FADD(SSW_FADD(y[i], SSW_FMUL(-1.105450f, ci)), SSW_FMUL(1.729860f, cq));
        }
        if constexpr (DST == PIX_RGB8) {
            unsigned* d = (unsigned*)((unsigned char*)dst + 3 * pix4);
            d[0] = pack_u8x4(o);
            d[1] = pack_u8x4(o + 4);
            d[2] = pack_u8x4(o + 8);
        } else {
            float* d = (float*)dst + 3 * pix4;
#pragma unroll
            for (int i = 0; i < 12; ++i) o[i] = clamp01(o[i]);
            st4(d, o[0], o[1], o[2], o[3]);
            st4(d + 4, o[4], o[5], o[6], o[7]);
            st4(d + 8, o[8], o[9], o[10], o[11]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Two rows at once (device): lane x = row A, lane y = row B of the CTA's row pair, packed FP32 (FADD2 / FMUL2 /
// FFMA2).  Every lane performs exactly the IEEE operations of the scalar helpers above -- separately rounded
// products (fma(m, c, -0) == rn(m*c)), the same association -- so the results are bit-identical to them; the host
// forms (tests/emul) simply call the scalar helpers per row.
// ------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
SSW_HD float2 prod2(float m, float2 c, float nz) { return __ffma2_rn(make_float2(m, m), c, make_float2(nz, nz)); }   // rn(m*c) per lane
SSW_HD float2 mat3x2(float m0, float m1, float m2, float2 a, float2 b, float2 c, float nz) {
    return __fadd2_rn(__fadd2_rn(prod2(m0, a, nz), prod2(m1, b, nz)), prod2(m2, c, nz));
}
SSW_HD float2 unit_of2(float2 x) {
    const float c_hi = 0.0039215688593685626983642578125f, c_lo = -2.31917579870781060424633324146270751953125e-10f;
    return __ffma2_rn(x, make_float2(c_hi, c_hi), __fmul2_rn(x, make_float2(c_lo, c_lo)));
}
template <int J>
SSW_HD float2 byte_to_float2(unsigned wa, unsigned wb) {
    return __fadd2_rn(make_float2(__uint_as_float(__byte_perm(wa, 0x4B000000u, 0x7540u + J)),
                                  __uint_as_float(__byte_perm(wb, 0x4B000000u, 0x7540u + J))), make_float2(-8388608.0f, -8388608.0f));
}
SSW_HD void unpack4_unit2(const unsigned* wa, const unsigned* wb, float2* c) {
    c[0] = unit_of2(byte_to_float2<0>(wa[0], wb[0])); c[1] = unit_of2(byte_to_float2<1>(wa[0], wb[0]));
    c[2] = unit_of2(byte_to_float2<2>(wa[0], wb[0])); c[3] = unit_of2(byte_to_float2<3>(wa[0], wb[0]));
    c[4] = unit_of2(byte_to_float2<0>(wa[1], wb[1])); c[5] = unit_of2(byte_to_float2<1>(wa[1], wb[1]));
    c[6] = unit_of2(byte_to_float2<2>(wa[1], wb[1])); c[7] = unit_of2(byte_to_float2<3>(wa[1], wb[1]));
    c[8] = unit_of2(byte_to_float2<0>(wa[2], wb[2])); c[9] = unit_of2(byte_to_float2<1>(wa[2], wb[2]));
    c[10] = unit_of2(byte_to_float2<2>(wa[2], wb[2])); c[11] = unit_of2(byte_to_float2<3>(wa[2], wb[2]));
}
// four (A,B) pairs of unit values -> one RGB8 word per row
#ifndef SSW_PACK_F2I
#define SSW_PACK_F2I 1
#endif
// float -> u8, truncating, saturating to [0, 255], NaN -> 0 (cvt.rzi.u8.f32 clamps by definition; SASS: F2IP.U8.F32.TRUNC)
SSW_HD unsigned sat_u8_rz(float t) { unsigned r; asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(r) : "f"(t)); return r; }
SSW_HD void pack_u8x4x2(const float2* o, float nz, unsigned& wa, unsigned& wb) {
    unsigned ma[4], mb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#if SSW_PACK_F2I
        // round(clamp(v, 0, 1) * 255), ties away from zero, == sat_u8(trunc(RZ(rn(255 v) + 0.5))) for EVERY float v: inside
        // [0, 1] the forms coincide (see unit_to_u8_fast), below 0 / above 1 / NaN the saturating conversion gives the 0 /
        // 255 / 0 of the clamp (rn(255 v) is monotone, so v < 0 cannot reach 0.5 and v > 1 cannot fall below 255).  The
        // clamp (4 FMNMX per pair) and the second magic add leave the FP32 / ALU pipes, which bound the inverse row pass;
        // the conversion unit is otherwise idle there.  Checked against unit_to_u8(clamp01(v)) over all 2^32 bit patterns
        // on the device (ssw_selftest_pack_u8, tests/test_gpu_parity.py).
        const float2 t = __fadd2_rz(prod2(255.0f, o[i], nz), make_float2(0.5f, 0.5f));
        ma[i] = sat_u8_rz(t.x); mb[i] = sat_u8_rz(t.y);
#else
        const float2 c = make_float2(fminf(fmaxf(o[i].x, 0.0f), 1.0f), fminf(fmaxf(o[i].y, 0.0f), 1.0f));   // NaN -> 0
        // the product must be rounded to f32 BEFORE the +0.5 (round half away from zero of the f32 value): prod2, not
        // FMUL2 -- ptxas would contract FMUL2 + FADD2.RZ into one FFMA2.RZ
        const float2 m = __fadd2_rz(__fadd2_rz(prod2(255.0f, c, nz), make_float2(0.5f, 0.5f)), make_float2(8388608.0f, 8388608.0f));
        ma[i] = __float_as_uint(m.x); mb[i] = __float_as_uint(m.y);
#endif
    }
    wa = __byte_perm(__byte_perm(ma[0], ma[1], 0x0040u), __byte_perm(ma[2], ma[3], 0x0040u), 0x5410u);
    wb = __byte_perm(__byte_perm(mb[0], mb[1], 0x0040u), __byte_perm(mb[2], mb[3], 0x0040u), 0x5410u);
}
#endif

// 12 bytes (4 RGB8 pixels) of row A and of row B, as three words each -> luma pairs y2[i] = (A_i, B_i)
SSW_HD void luma4x2_words(const unsigned* wa, const unsigned* wb, float nz, cplx* y2) {
#if defined(__CUDA_ARCH__)
    float2 c[12];
    unpack4_unit2(wa, wb, c);
#pragma unroll
    for (int i = 0; i < 4; ++i) y2[i] = mat3x2(0.30f, 0.59f, 0.11f, c[3 * i], c[3 * i + 1], c[3 * i + 2], nz);
#else
    (void)nz;
    float ca[12], cb[12];
    unpack4_unit(wa[0], wa[1], wa[2], ca);
    unpack4_unit(wb[0], wb[1], wb[2], cb);
    for (int i = 0; i < 4; ++i)
        y2[i] = mk(rgb_to_y(ca[3 * i], ca[3 * i + 1], ca[3 * i + 2]), rgb_to_y(cb[3 * i], cb[3 * i + 1], cb[3 * i + 2]));
#endif
}

// 4 pixels of row A (pixel index pa4) and of row B (pb4) -> luma pairs y2[i] = (A_i, B_i).  Rows that do not exist
// (hb false: odd frame height) read row A again and are zeroed.
template <int SRC>
SSW_HD void load_luma4x2(const void* src, long long pa4, long long pb4, bool hb, float nz, cplx* y2) {
#if defined(__CUDA_ARCH__)
    if constexpr (SRC == PIX_RGB8) {
        const unsigned* qa = (const unsigned*)((const unsigned char*)src + 3 * pa4);
        const unsigned* qb = (const unsigned*)((const unsigned char*)src + 3 * (hb ? pb4 : pa4));
        const unsigned wa[3] = {ldw(qa), ldw(qa + 1), ldw(qa + 2)}, wb[3] = {ldw(qb), ldw(qb + 1), ldw(qb + 2)};
        luma4x2_words(wa, wb, nz, y2);
        if (!hb) {
#pragma unroll
            for (int i = 0; i < 4; ++i) y2[i].y = 0.f;
        }
        return;
    }
#endif
    (void)nz;
    float ya[4], yb[4] = {0.f, 0.f, 0.f, 0.f};
    load_luma4<SRC>(src, pa4, ya);
    if (hb) load_luma4<SRC>(src, pb4, yb);
#pragma unroll
    for (int i = 0; i < 4; ++i) y2[i] = mk(ya[i], yb[i]);
}

// new luma pairs y2[i] = (A_i, B_i) of 4 pixels of rows A and B + the 12 original bytes of each row (chroma) -> the 12
// output bytes of each row, as words.  oa / ob may alias wa / wb (the pipelines convert in place).
SSW_HD void rgb8_out4x2_words(const unsigned* wa, const unsigned* wb, float nz, const cplx* y2, unsigned* oa, unsigned* ob) {
#if defined(__CUDA_ARCH__)
    float2 c[12], o[12];
    unpack4_unit2(wa, wb, c);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 ci = mat3x2(0.60f, -0.28f, -0.32f, c[3 * i], c[3 * i + 1], c[3 * i + 2], nz);
        const float2 cq = mat3x2(0.21f, -0.52f, 0.31f, c[3 * i], c[3 * i + 1], c[3 * i + 2], nz);
        // src/yiq.rs:163-165 rows of YIQ_TO_RGB_MATRIX, (m0*y + m1*i) + m2*q, m0 == 1
        o[3 * i] = __fadd2_rn(__fadd2_rn(y2[i], prod2(0.948262f, ci, nz)), prod2(0.624013f, cq, nz));
        o[3 * i + 1] = __fadd2_rn(__fadd2_rn(y2[i], prod2(-0.276066f, ci, nz)), prod2(-0.639810f, cq, nz));
        o[3 * i + 2] = __fadd2_rn(__fadd2_rn(y2[i], prod2(-1.105450f, ci, nz)), prod2(1.729860f, cq, nz));
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) pack_u8x4x2(o + 4 * j, nz, oa[j], ob[j]);
#else
    (void)nz;
    for (int row = 0; row < 2; ++row) {
        const unsigned* wsrc = row ? wb : wa;
        float c[12], o[12];
        unpack4_unit(wsrc[0], wsrc[1], wsrc[2], c);
        for (int i = 0; i < 4; ++i) {
            const float y = row ? y2[i].y : y2[i].x;
            const float ci = rgb_to_i(c[3 * i], c[3 * i + 1], c[3 * i + 2]);
            const float cq = rgb_to_q(c[3 * i], c[3 * i + 1], c[3 * i + 2]);
            o[3 * i] = SSW_FADD(SSW_FADD(y, SSW_FMUL(0.948262f, ci)), SSW_FMUL(0.624013f, cq));
            o[3 * i + 1] = SSW_FADD(SSW_FADD(y, SSW_FMUL(-0.276066f, ci)), SSW_FMUL(-0.639810f, cq));
            o[3 * i + 2] = SSW_FADD(SSW_FADD(y, SSW_FMUL(-1.105450f, ci)), SSW_FMUL(1.729860f, cq));
        }
        unsigned* od = row ? ob : oa;
        od[0] = pack_u8x4(o); od[1] = pack_u8x4(o + 4); od[2] = pack_u8x4(o + 8);
    }
#endif
}

// new luma pairs y2[i] = (A_i, B_i) of 4 pixels of rows A and B (+ chroma of the original pixels) -> destination
template <int DST, int SRC>
SSW_HD void store_pix4x2(const void* src, void* dst, long long pa4, long long pb4, bool hb, float nz, const cplx* y2) {
#if defined(__CUDA_ARCH__)
    if constexpr (DST == PIX_RGB8 && SRC == PIX_RGB8) {
        const unsigned* qa = (const unsigned*)((const unsigned char*)src + 3 * pa4);
        const unsigned* qb = (const unsigned*)((const unsigned char*)src + 3 * (hb ? pb4 : pa4));
        const unsigned wa[3] = {ldw(qa), ldw(qa + 1), ldw(qa + 2)}, wb[3] = {ldw(qb), ldw(qb + 1), ldw(qb + 2)};
        unsigned ua[3], ub[3];
        rgb8_out4x2_words(wa, wb, nz, y2, ua, ub);
        unsigned* da = (unsigned*)((unsigned char*)dst + 3 * pa4);
        unsigned* db = (unsigned*)((unsigned char*)dst + 3 * pb4);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            da[j] = ua[j];
            if (hb) db[j] = ub[j];
        }
        return;
    }
#endif
    (void)nz;
    float ya[4], yb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ya[i] = y2[i].x; yb[i] = y2[i].y; }
    store_pix4<DST, SRC>(src, dst, pa4, ya);
    if (hb) store_pix4<DST, SRC>(src, dst, pb4, yb);
}

// FFT-input positions of 4 consecutive samples m = 4u..4u+3 (Makhoul): 2u, N-1-2u, 2u+1, N-2-2u.
// scatter (a_i, b_i) = sample i of line A / line B
template <class P>
SSW_HD void put4(cplx* s, int u, const float* a, const float* b) {
    if constexpr (!P::PAD) {
        st4(&s[2 * u].x, a[0], b[0], a[2], b[2]);
        st4(&s[P::N - 2 - 2 * u].x, a[3], b[3], a[1], b[1]);
    } else {
        s[P::idx(2 * u)] = mk(a[0], b[0]);
        s[P::idx(2 * u + 1)] = mk(a[2], b[2]);
        s[P::idx(P::N - 2 - 2 * u)] = mk(a[3], b[3]);
        s[P::idx(P::N - 1 - 2 * u)] = mk(a[1], b[1]);
    }
}
template <class P>
SSW_HD void put4x2(cplx* s, int u, const cplx* y2) {   // y2[i] = (sample i of line A, of line B)
    if constexpr (!P::PAD) {
        st4(&s[2 * u].x, y2[0].x, y2[0].y, y2[2].x, y2[2].y);
        st4(&s[P::N - 2 - 2 * u].x, y2[3].x, y2[3].y, y2[1].x, y2[1].y);
    } else {
        s[P::idx(2 * u)] = y2[0];
        s[P::idx(2 * u + 1)] = y2[2];
        s[P::idx(P::N - 2 - 2 * u)] = y2[3];
        s[P::idx(P::N - 1 - 2 * u)] = y2[1];
    }
}
template <class P>
SSW_HD void get4(const cplx* s, int u, cplx* f) {  // f[i] = FFT value belonging to sample 4u+i
    if constexpr (!P::PAD) {
        const float* lo = &s[2 * u].x;
        const float* hi = &s[P::N - 2 - 2 * u].x;
        f[0] = mk(lo[0], lo[1]); f[2] = mk(lo[2], lo[3]); f[3] = mk(hi[0], hi[1]); f[1] = mk(hi[2], hi[3]);
    } else {
        f[0] = s[P::idx(2 * u)]; f[2] = s[P::idx(2 * u + 1)];
        f[3] = s[P::idx(P::N - 2 - 2 * u)]; f[1] = s[P::idx(P::N - 1 - 2 * u)];
    }
}

template <class P>
struct ThreadState { cplx v[MaxRegs<P>::value]; };

// ------------------------------------------------------------------------------------------------
// forward row pass: pixels / plane rows -> luma -> DCT-II along x -> coefficient plane
// tile = G row pairs; team g (T threads) owns rows 2*(tile*G+g), +1
// ------------------------------------------------------------------------------------------------
template <class P_, int G_, int SRC_>
struct RowFwd {
    using P = P_;
    static constexpr int G = G_, SRC = SRC_, THREADS = G_ * P_::T, NPH = 2 + 2 * P_::NST;
    static constexpr int SMEM = G_ * P_::PITCH * (int)sizeof(cplx);
    static constexpr int MINB = 0;
    using Thread = ThreadState<P_>;
    static int tiles_per_image(int w, int h) { (void)w; return ((h + 1) / 2 + G - 1) / G; }

    template <int PH>
    static SSW_HD void phase(const FastArgs& a, cplx* smem, int tile, int tid, Thread& th) {
        constexpr int N = P::N, T = P::T;
        const int g = tid / T, t = tid - g * T;
        cplx* s = smem + g * P::PITCH;
        const int img = tile / a.tiles_per_image;
        const int ra = 2 * ((tile - img * a.tiles_per_image) * G + g), rb = ra + 1;
        if constexpr (PH == 0) {
            const bool ha = ra < a.h, hb = rb < a.h;
            const void* src = image_base<SRC>(a.src, img, a.src_stride);
            const long long rowa = (long long)ra * N;
#pragma unroll
            for (int it = 0; it < (N / 4 + T - 1) / T; ++it) {
                const int u = t + it * T;
                if (u < N / 4) {
                    cplx y2[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) y2[i] = mk(0.f, 0.f);
#ifdef SSW_TUNE
                    if (a.dbg_skip & 1) { y2[0].x = (float)u; y2[1].y = 1.f; } else
#endif
                    if (SRC == PIX_PLANE && a.seg_shift >= 0) {
                        if (ha) load_luma4x2<SRC>(src, seg_index(a, ra, 4 * u), seg_index(a, hb ? rb : ra, 4 * u), hb, a.neg_zero, y2);
                    } else {
                        if (ha) load_luma4x2<SRC>(src, rowa + 4 * u, rowa + N + 4 * u, hb, a.neg_zero, y2);
                    }
                    put4x2<P>(s, u, y2);
                }
            }
        } else if constexpr (PH < NPH - 1) {
#ifdef SSW_TUNE
            if (a.dbg_skip & 4) return;
#endif
            fft_phase<P, PH>(s, a.tw, t, th.v);
        } else {
            if (ra >= a.h) return;
#ifdef SSW_TUNE
            if (a.dbg_skip & 2) { if (s[P::idx(t)].x == 123.456f) a.plane[t] = 1.f; return; }
#endif
            const bool hb = rb < a.h;
            float* oa = a.plane + img * a.plane_stride + (long long)ra * N;
            float* ob = oa + N;
#pragma unroll 4
            for (int k = t; k <= N / 2; k += T) {
                const int kr = k ? N - k : 0;
                float xa, xb, ya, yb;
                dct2_post(s[P::idx(k)], s[P::idx(kr)], SSW_LDG(&a.t4[k]), xa, xb, ya, yb);
                const float sk = k ? a.scalen : a.scale0;
                oa[k] = xa * sk;
                if (hb) ob[k] = xb * sk;
                if (k && kr != k) {
                    oa[kr] = ya * a.scalen;
                    if (hb) ob[kr] = yb * a.scalen;
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// inverse row pass: plane rows -> DCT-III along x -> scale -> plane | Y' + I,Q(original pixels) -> RGB
// ------------------------------------------------------------------------------------------------
template <class P_, int G_, int DST_, int SRC_>
struct RowInv {
    using P = P_;
    static constexpr int G = G_, DST = DST_, SRC = SRC_, THREADS = G_ * P_::T, NPH = 2 + 2 * P_::NST;
    static constexpr int SMEM = G_ * P_::PITCH * (int)sizeof(cplx);
    static constexpr int MINB = 0;
    using Thread = ThreadState<P_>;
    static int tiles_per_image(int w, int h) { (void)w; return ((h + 1) / 2 + G - 1) / G; }

    template <int PH>
    static SSW_HD void phase(const FastArgs& a, cplx* smem, int tile, int tid, Thread& th) {
        constexpr int N = P::N, T = P::T;
        const int g = tid / T, t = tid - g * T;
        cplx* s = smem + g * P::PITCH;
        const int img = tile / a.tiles_per_image;
        const int ra = 2 * ((tile - img * a.tiles_per_image) * G + g), rb = ra + 1;
        const bool ha = ra < a.h, hb = rb < a.h;
        if constexpr (PH == 0) {
            const float* ia = a.plane + img * a.plane_stride + (long long)ra * N;
            const float* ib = ia + N;
            const bool seg = a.seg_shift >= 0;   // coefficient lines held as all-to-all blocks (sharded frames)
#pragma unroll 4
            for (int k = t; k <= N / 2; k += T) {
                const int kr = k ? N - k : 0;
                float pa = 0.f, pb = 0.f, qa = 0.f, qb = 0.f;
                if (!seg) {
                    pa = ha ? ia[k] : 0.f; pb = hb ? ib[k] : 0.f;
                    qa = (k && ha) ? ia[kr] : 0.f; qb = (k && hb) ? ib[kr] : 0.f;
                } else {
                    if (ha) { pa = a.plane[seg_index(a, ra, k)]; if (k) qa = a.plane[seg_index(a, ra, kr)]; }
                    if (hb) { pb = a.plane[seg_index(a, rb, k)]; if (k) qb = a.plane[seg_index(a, rb, kr)]; }
                }
                cplx zk, zr;
                dct3_pre(pa, pb, qa, qb, SSW_LDG(&a.t4[k]), zk, zr);
                s[P::idx(k)] = zk;
                if (k && kr != k) s[P::idx(kr)] = zr;
            }
        } else if constexpr (PH < NPH - 1) {
            fft_phase<P, PH>(s, a.tw, t, th.v);
        } else {
            if (!ha) return;
            const void* src = (DST == PIX_PLANE) ? nullptr : image_base<SRC>(a.src, img, a.src_stride);
            void* dst = const_cast<void*>(image_base<DST>(a.dst, img, a.dst_stride));
            const long long row = (long long)ra * N;
#pragma unroll
            for (int it = 0; it < (N / 4 + T - 1) / T; ++it) {
                const int u = t + it * T;
                if (u < N / 4) {
                    cplx f[4];
                    get4<P>(s, u, f);
                    cplx y2[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) y2[i] = cmul_lanes_x(f[i], a.scale0, -a.scale0, a.neg_zero);   // feeds the colour adds: no contraction
                    store_pix4x2<DST, SRC>(src, dst, row + 4 * u, row + N + 4 * u, hb, a.neg_zero, y2);
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// column passes, in place: tile = G column pairs = 2G adjacent columns (G even: 128-bit accesses);
// a float2 of two adjacent columns in one row *is* one complex sample of the packed line pair.
// TEAMS teams of T threads transform the G pairs in G/TEAMS rounds.
// ------------------------------------------------------------------------------------------------
template <class P_, int G_, int TEAMS_, bool INVERSE_, int MINB_ = 0>
struct ColPass {
    using P = P_;
    static_assert(G_ % 2 == 0 && G_ % TEAMS_ == 0, "column tiles: even number of pairs, whole rounds");
    static constexpr int G = G_, H = G_ / 2, TEAMS = TEAMS_, ROUNDS = G_ / TEAMS_, THREADS = TEAMS_ * P_::T;
    static constexpr int NPH = 2 + 2 * P_::NST * ROUNDS;
    static constexpr int SMEM = G_ * P_::PITCH * (int)sizeof(cplx);
    static constexpr int MINB = MINB_;
    using Thread = ThreadState<P_>;
    static int tiles_per_image(int w, int h) { (void)h; return (w / 2 + G - 1) / G; }

    template <int PH>
    static SSW_HD void phase(const FastArgs& a, cplx* smem, int tile, int tid, Thread& th) {
        constexpr int N = P::N, T = P::T;
        const int w = a.w;
        const int img = tile / a.tiles_per_image;
        const int c0 = (tile - img * a.tiles_per_image) * 2 * G;
        float* plane = a.plane + img * a.plane_stride;
        if constexpr (PH > 0 && PH < NPH - 1) {
            constexpr int RD = (PH - 1) / (2 * P::NST), SUB = (PH - 1) % (2 * P::NST) + 1;
            const int g = tid / T, t = tid - g * T;
            fft_phase<P, SUB>(smem + (RD * TEAMS + g) * P::PITCH, a.tw, t, th.v);
        } else if constexpr ((PH == 0) != INVERSE_) {
            // sample-domain side: forward load (PH == 0) or inverse store (PH == NPH-1)
#pragma unroll 6
            for (int it = 0; it < (N * H + THREADS - 1) / THREADS; ++it) {
                const int e = tid + it * THREADS;
                const int r = e / H, q = e - r * H;
                const int c = c0 + 4 * q;
                if (e < N * H && c < w) {
                    float* gp = plane + (long long)r * w + c;
                    cplx* s0 = smem + (2 * q) * P::PITCH + P::idx(makhoul(r, N));
                    if constexpr (!INVERSE_) {
                        const f4 v = ld4(gp);
                        s0[0] = mk(v.a, v.b);
                        s0[P::PITCH] = mk(v.c, v.d);
                    } else {
                        const cplx f0 = s0[0], f1 = s0[P::PITCH];
                        st4(gp, f0.x * a.scale0, -f0.y * a.scale0, f1.x * a.scale0, -f1.y * a.scale0);
                    }
                }
            }
            if (!INVERSE_ && c0 + 2 * G > w) {
                // columns beyond the frame (last tile): keep the FFT input finite
                for (int e = tid; e < N * H; e += THREADS) {
                    const int r = e / H, q = e - r * H;
                    if (c0 + 4 * q >= w) {
                        cplx* s0 = smem + (2 * q) * P::PITCH + P::idx(r);
                        s0[0] = mk(0.f, 0.f);
                        s0[P::PITCH] = mk(0.f, 0.f);
                    }
                }
            }
        } else {
            // coefficient-domain side: forward store (post pass) or inverse load (pre pass)
#pragma unroll(P::PAD ? 2 : 4)
            for (int e = tid; e < (N / 2 + 1) * H; e += THREADS) {
                const int k = e / H, q = e - k * H;
                const int c = c0 + 4 * q;
                const int kr = k ? N - k : 0;
                cplx* s0 = smem + (2 * q) * P::PITCH;
                cplx* s1 = s0 + P::PITCH;
                const cplx tw = SSW_LDG(&a.t4[k]);
                if constexpr (!INVERSE_) {
                    if (c >= w) continue;
                    float xa0, xb0, ya0, yb0, xa1, xb1, ya1, yb1;
                    dct2_post(s0[P::idx(k)], s0[P::idx(kr)], tw, xa0, xb0, ya0, yb0);
                    dct2_post(s1[P::idx(k)], s1[P::idx(kr)], tw, xa1, xb1, ya1, yb1);
                    const float sk = k ? a.scalen : a.scale0;
                    st4(plane + (long long)k * w + c, xa0 * sk, xb0 * sk, xa1 * sk, xb1 * sk);
                    if (k && kr != k)
                        st4(plane + (long long)kr * w + c, ya0 * a.scalen, yb0 * a.scalen, ya1 * a.scalen, yb1 * a.scalen);
                } else {
                    f4 pv = f4{0.f, 0.f, 0.f, 0.f}, qv = f4{0.f, 0.f, 0.f, 0.f};
                    if (c < w) {
                        pv = ld4(plane + (long long)k * w + c);
                        if (k) qv = ld4(plane + (long long)kr * w + c);
                    }
                    cplx zk, zr;
                    dct3_pre(pv.a, pv.b, qv.a, qv.b, tw, zk, zr);
                    s0[P::idx(k)] = zk;
                    if (k && kr != k) s0[P::idx(kr)] = zr;
                    dct3_pre(pv.c, pv.d, qv.c, qv.d, tw, zk, zr);
                    s1[P::idx(k)] = zk;
                    if (k && kr != k) s1[P::idx(kr)] = zr;
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// single-line kernels: one real line of length N = 2M per CTA through an M-point complex FFT (real-FFT
// split).  A packed line PAIR of N complex samples needs 8N bytes of shared memory, which stops at
// N ~ 28000; this form needs 4N bytes and carries the 32768-point lines of the gigapixel frames.
//   forward : v = Makhoul(x);  z[m] = v[2m] + i v[2m+1];  Z = FFT_M(z);  with P = (Z_k + conj Z_{M-k})/2,
//             Q = W_N^k (-i)(Z_k - conj Z_{M-k})/2:  V_k = P + Q,  V_{M-k} = conj(P - Q);
//             X_k = 2 Re(t_k V_k),  X_{N-k} = -2 Im(t_k V_k)            (t_k = exp(-i pi k / 2N))
//   inverse : V_k = conj(t_k)(X_k - i X_{N-k})/2;  A_k = (V_k + V_{k+M})/2 + i conj(W_N^k)(V_k - V_{k+M})/2;
//             (y[2m], y[2m+1]) = conj(FFT_M(conj A))[m], un-permuted   (= 0.25 * scipy dct type 3)
// P_ is the plan of the M-point FFT.  Verified against the oracle by tests/test_emul_fast.py.
// ------------------------------------------------------------------------------------------------
SSW_HD cplx cconj(cplx a) { return mk(a.x, -a.y); }
SSW_HD cplx cscale(cplx a, float f) { return cscale2(a, f); }

template <class P>
SSW_HD cplx line1_w(const cplx* t4, int k) {  // W_N^k = exp(-2 pi i k / N) = t4[4k], N = 2*P::N, k <= N/4
    return (4 * k == 2 * P::N) ? mk(0.f, -1.f) : SSW_LDG(&t4[4 * k]);
}

template <class P_, int SRC_>
struct Line1Fwd {
    using P = P_;
    static constexpr int M = P_::N, N = 2 * P_::N, SRC = SRC_, THREADS = P_::T, NPH = 2 + 2 * P_::NST;
    static constexpr int SMEM = P_::PITCH * (int)sizeof(cplx);
    static constexpr int MINB = 0;
    using Thread = ThreadState<P_>;
    static int tiles_per_image(int w, int h) { (void)w; return h; }

    template <int PH>
    static SSW_HD void phase(const FastArgs& a, cplx* s, int tile, int t, Thread& th) {
        constexpr int T = P::T;
        const int img = tile / a.tiles_per_image;
        const int row = tile - img * a.tiles_per_image;
        if constexpr (PH == 0) {
            const void* src = image_base<SRC>(a.src, img, a.src_stride);
#pragma unroll 2
            for (int u = t; u < N / 4; u += T) {
                float y[4];
                load_luma4<SRC>(src, (SRC == PIX_PLANE && a.seg_shift >= 0) ? seg_index(a, row, 4 * u) : (long long)row * N + 4 * u, y);
                s[P::idx(u)] = mk(y[0], y[2]);
                s[P::idx(M - 1 - u)] = mk(y[3], y[1]);
            }
        } else if constexpr (PH < NPH - 1) {
            fft_phase<P, PH>(s, a.tw, t, th.v);
        } else {
            float* o = a.plane + img * a.plane_stride + (long long)row * N;
#pragma unroll 2
            for (int k = t; k <= M / 2; k += T) {
                const cplx zk = s[P::idx(k)], zr = cconj(s[P::idx(k ? M - k : 0)]);
                const cplx p = cscale(cadd(zk, zr), 0.5f);
                const cplx q = cmul(line1_w<P>(a.t4, k), cscale(mul_mi(csub(zk, zr)), 0.5f));
                const cplx u1 = cmul(SSW_LDG(&a.t4[k]), cadd(p, q));
                const cplx u2 = cmul(SSW_LDG(&a.t4[M - k]), cconj(csub(p, q)));
                o[k] = 2.f * u1.x * (k ? a.scalen : a.scale0);
                o[M - k] = 2.f * u2.x * a.scalen;
                if (k) {
                    o[N - k] = -2.f * u1.y * a.scalen;
                    o[M + k] = -2.f * u2.y * a.scalen;
                }
            }
        }
    }
};

template <class P_, int DST_, int SRC_>
struct Line1Inv {
    using P = P_;
    static constexpr int M = P_::N, N = 2 * P_::N, DST = DST_, SRC = SRC_, THREADS = P_::T, NPH = 2 + 2 * P_::NST;
    static constexpr int SMEM = P_::PITCH * (int)sizeof(cplx);
    static constexpr int MINB = 0;
    using Thread = ThreadState<P_>;
    static int tiles_per_image(int w, int h) { (void)w; return h; }

    template <int PH>
    static SSW_HD void phase(const FastArgs& a, cplx* s, int tile, int t, Thread& th) {
        constexpr int T = P::T;
        const int img = tile / a.tiles_per_image;
        const int row = tile - img * a.tiles_per_image;
        if constexpr (PH == 0) {
            const float* x = a.plane + img * a.plane_stride + (long long)row * N;
            const bool seg = a.seg_shift >= 0;   // coefficient line held as all-to-all blocks (sharded frames)
            auto X = [&](int j) -> float { return seg ? a.plane[seg_index(a, row, j)] : x[j]; };
#pragma unroll 2
            for (int k = t; k <= M / 2; k += T) {
                if (k == 0) {
                    const float v0 = 0.5f * X(0), vm = 0.70710678118654752440f * X(M);
                    s[P::idx(0)] = mk(0.5f * (v0 + vm), -0.5f * (v0 - vm));  // conj(A_0)
                    continue;
                }
                const float xk = X(k), xmk = X(M - k), xpk = X(M + k), xnk = X(N - k);
                const cplx tk = SSW_LDG(&a.t4[k]), tm = SSW_LDG(&a.t4[M - k]);
                const cplx tkM = mul_mi(cconj(tm));   // t_{M+k} = -i conj(t_{M-k})
                const cplx tnk = mul_mi(cconj(tk));   // t_{N-k} = -i conj(t_k)
                const cplx w = line1_w<P>(a.t4, k);
                // V_j = conj(t_j) (X_j - i X_{N-j}) / 2
                const cplx vk = cscale(cmul(cconj(tk), mk(xk, -xnk)), 0.5f);
                const cplx vkM = cscale(cmul(cconj(tkM), mk(xpk, -xmk)), 0.5f);
                const cplx vm = cscale(cmul(cconj(tm), mk(xmk, -xpk)), 0.5f);
                const cplx vnk = cscale(cmul(cconj(tnk), mk(xnk, -xk)), 0.5f);
                // A_k = (V_k + V_{k+M})/2 + i conj(W)(V_k - V_{k+M})/2 ;  A_{M-k}: W^{M-k} = -conj(W)
                const cplx ak = cadd(cscale(cadd(vk, vkM), 0.5f), mul_pi(cscale(cmul(cconj(w), csub(vk, vkM)), 0.5f)));
                const cplx am = cadd(cscale(cadd(vm, vnk), 0.5f), mul_pi(cscale(cmul(mk(-w.x, -w.y), csub(vm, vnk)), 0.5f)));
                s[P::idx(k)] = cconj(ak);
                s[P::idx(M - k)] = cconj(am);
            }
        } else if constexpr (PH < NPH - 1) {
            fft_phase<P, PH>(s, a.tw, t, th.v);
        } else {
            const void* src = (DST == PIX_PLANE) ? nullptr : image_base<SRC>(a.src, img, a.src_stride);
            void* dst = const_cast<void*>(image_base<DST>(a.dst, img, a.dst_stride));
#pragma unroll 2
            for (int u = t; u < N / 4; u += T) {
                const cplx f = s[P::idx(u)], g = s[P::idx(M - 1 - u)];
                float y[4];
                y[0] = f.x * a.scale0; y[2] = -f.y * a.scale0; y[3] = g.x * a.scale0; y[1] = -g.y * a.scale0;
                store_pix4<DST, SRC>(src, dst, (long long)row * N + 4 * u, y);
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// forward row pass with asynchronous prefetch (RGB8 rows): a CTA works through `tiles_per_cta` consecutive
// tiles; while the FFT of tile i runs, the raw bytes of tile i+1 stream into a staging buffer with
// cp.async (LDGSTS, no registers, no scoreboard stall), so only the first tile of a CTA waits for HBM.
// Shared memory: G line-pair buffers + G * 2 rows * 3N bytes of staging.
// ------------------------------------------------------------------------------------------------
SSW_HD void async_copy16(void* smem_dst, const void* gmem_src) {
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
#else
    __builtin_memcpy(smem_dst, gmem_src, 16);
#endif
}
SSW_HD void async_commit_wait_all() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#endif
}

template <class P_, int G_>
struct RowFwdPF {
    using P = P_;
    static_assert((3 * P_::N) % 16 == 0, "rows must be whole 16-byte chunks");
    static constexpr int G = G_, SRC = PIX_RGB8, THREADS = G_ * P_::T, NPH = 2 + 2 * P_::NST;
    static constexpr int ROW_BYTES = 3 * P_::N;
    static constexpr int FFT_BYTES = G_ * P_::PITCH * (int)sizeof(cplx);
    static constexpr int SMEM = FFT_BYTES + G_ * 2 * ROW_BYTES;
    static constexpr int MINB = 0;
    static constexpr bool PREFETCH = true;
    using Thread = ThreadState<P_>;
    static int tiles_per_image(int w, int h) { (void)w; return ((h + 1) / 2 + G - 1) / G; }

    // issue the copies of tile `tile` into the staging buffer (all threads of the CTA)
    static SSW_HD void prefetch(const FastArgs& a, cplx* smem, int tile, int tid) {
        constexpr int N = P::N, CH = ROW_BYTES / 16;
        const int img = tile / a.tiles_per_image;
        const int row0 = 2 * ((tile - img * a.tiles_per_image) * G);   // first of the 2G rows of the tile
        const unsigned char* src = (const unsigned char*)a.src + 3 * (img * a.src_stride + (long long)row0 * N);
        unsigned char* stage = (unsigned char*)smem + FFT_BYTES;
        for (int e = tid; e < 2 * G * CH; e += THREADS) {
            const int r = e / CH, c = e - r * CH;
            if (row0 + r < a.h) async_copy16(stage + r * ROW_BYTES + 16 * c, src + (long long)r * ROW_BYTES + 16 * c);
        }
    }

    template <int PH>
    static SSW_HD void phase(const FastArgs& a, cplx* smem, int tile, int tid, Thread& th) {
        constexpr int N = P::N, T = P::T;
        const int g = tid / T, t = tid - g * T;
        cplx* s = smem + g * P::PITCH;
        const int img = tile / a.tiles_per_image;
        const int ra = 2 * ((tile - img * a.tiles_per_image) * G + g), rb = ra + 1;
        if constexpr (PH == 0) {
            const bool ha = ra < a.h, hb = rb < a.h;
            const unsigned* sa = (const unsigned*)((const unsigned char*)smem + FFT_BYTES + (2 * g) * ROW_BYTES);
            const unsigned* sb = (const unsigned*)((const unsigned char*)sa + ROW_BYTES);
#pragma unroll
            for (int it = 0; it < (N / 4 + T - 1) / T; ++it) {
                const int u = t + it * T;
                if (u < N / 4) {
                    float ya[4] = {0.f, 0.f, 0.f, 0.f}, yb[4] = {0.f, 0.f, 0.f, 0.f};
                    float c[12];
                    if (ha) {
                        unpack4_unit(sa[3 * u], sa[3 * u + 1], sa[3 * u + 2], c);
#pragma unroll
                        for (int i = 0; i < 4; ++i) ya[i] = rgb_to_y(c[3 * i], c[3 * i + 1], c[3 * i + 2]);
                    }
                    if (hb) {
                        unpack4_unit(sb[3 * u], sb[3 * u + 1], sb[3 * u + 2], c);
#pragma unroll
                        for (int i = 0; i < 4; ++i) yb[i] = rgb_to_y(c[3 * i], c[3 * i + 1], c[3 * i + 2]);
                    }
                    put4<P>(s, u, ya, yb);
                }
            }
        } else {
            RowFwd<P_, G_, PIX_RGB8>::template phase<PH>(a, smem, tile, tid, th);   // FFT stages + post pass are shared
        }
    }
};

#if defined(__CUDACC__)
template <class K>
constexpr int min_blocks() {
    if (K::MINB > 0) return K::MINB;
    // aim at 1024 resident threads per SM (64 registers each), bounded by shared memory
    int by_threads = 1024 / K::THREADS;
    int by_smem = (227 * 1024) / (K::SMEM + 1024);
    int m = by_threads < by_smem ? by_threads : by_smem;
    return m < 1 ? 1 : m;
}

template <class K>
__global__ void __launch_bounds__(K::THREADS, min_blocks<K>()) fast_kernel(const __grid_constant__ FastArgs a) {
    // one CTA per tile.  (A persistent one-wave grid striding over the tiles was measured: no faster, and the
    // loop state cost registers in the widest kernels.)
    extern __shared__ __align__(16) unsigned char fast_smem[];
    typename K::Thread th;
    if (!a.pdl_late) pdl_trigger();
    pdl_wait();
    static_for<K::NPH>([&](auto ph) {
        constexpr int p = decltype(ph)::value;
        if constexpr (p == K::NPH - 1) { if (a.pdl_late) pdl_trigger(); }   // only the output phase is left
        K::template phase<p>(a, (cplx*)fast_smem, blockIdx.x, threadIdx.x, th);
        if constexpr (p + 1 < K::NPH) __syncthreads();
    });
}

// prefetching kernels: CTA b works through tiles [b*per, (b+1)*per)
template <class K>
__global__ void __launch_bounds__(K::THREADS, min_blocks<K>()) fast_kernel_pf(const __grid_constant__ FastArgs a) {
    extern __shared__ __align__(16) unsigned char fast_smem[];
    typename K::Thread th;
    if (!a.pdl_late) pdl_trigger();
    pdl_wait();
    int tile = blockIdx.x * a.tiles_per_cta;
    const int end = min(tile + a.tiles_per_cta, a.total_tiles);
    if (tile < end) K::prefetch(a, (cplx*)fast_smem, tile, threadIdx.x);
    for (; tile < end; ++tile) {
        async_commit_wait_all();
        __syncthreads();                                   // staging of `tile` has landed for everybody
        K::template phase<0>(a, (cplx*)fast_smem, tile, threadIdx.x, th);
        __syncthreads();                                   // staging consumed
        if (tile + 1 < end) K::prefetch(a, (cplx*)fast_smem, tile + 1, threadIdx.x);
        static_for<K::NPH - 1>([&](auto ph) {
            constexpr int p = decltype(ph)::value + 1;
            if constexpr (p == K::NPH - 1) { if (a.pdl_late && tile + 1 == end) pdl_trigger(); }
            K::template phase<p>(a, (cplx*)fast_smem, tile, threadIdx.x, th);
            __syncthreads();
        });
    }
}
#endif

// host: stage twiddles of a plan, layout per stage s >= 1: [r-1][k], k < ns(s); exp(-2*pi*i*k*r/(ns*R))
template <class P>
inline void make_stage_twiddles(float* out /* 2*P::TW_TOTAL floats */) {
    int o = 0;
    for (int s = 1; s < P::NST; ++s) {
        const int R = P::radix(s), ns = P::ns(s);
        for (int r = 1; r < R; ++r)
            for (int k = 0; k < ns; ++k) {
                const double ang = -2.0 * M_PI * (double)k * (double)r / ((double)ns * (double)R);
                out[2 * o] = (float)std::cos(ang);
                out[2 * o + 1] = (float)std::sin(ang);
                ++o;
            }
    }
}

// the planned lengths: 4K / 1080p frames in both orientations (+ the reference fixture's width)
using Plan3840 = Plan<3840, 256, 15, 16, 16>;
using Plan2160 = Plan<2160, 192, 15, 12, 12>;
using Plan1920 = Plan<1920, 128, 15, 16, 8>;
using Plan1080 = Plan<1080, 96, 15, 6, 12>;
using Plan640 = Plan<640, 64, 5, 8, 16>;
#ifdef SSW_TUNE
using Plan3840b = Plan<3840, 480, 5, 8, 8, 12>;   // small radices, more threads per line pair
using Plan3840c = Plan<3840, 384, 15, 8, 8, 4>;
using Plan3840d = Plan<3840, 128, 15, 16, 16>;   // two butterflies per thread (more ILP, half the warps)
#endif
// other common video formats: 720p, 1440p, 8K in both orientations
using Plan1280 = Plan<1280, 96, 5, 16, 16>;
using Plan720 = Plan<720, 96, 5, 9, 16>;
using Plan2560 = Plan<2560, 320, 5, 8, 8, 8>;
using Plan1440 = Plan<1440, 160, 9, 10, 16>;
using Plan7680 = Plan<7680, 512, 15, 16, 16, 2>;
using Plan4320 = Plan<4320, 384, 15, 12, 12, 2>;
// powers of two (first radix even -> padded layout)
using Plan1024 = Plan<1024, 64, 16, 16, 4>;
using Plan2048 = Plan<2048, 128, 16, 16, 8>;
using Plan4096 = Plan<4096, 256, 16, 16, 16>;
using Plan8192 = Plan<8192, 512, 16, 16, 16, 2>;
using Plan16384 = Plan<16384, 1024, 16, 16, 16, 4>;
// plans of the M = N/2 point FFT of the single-line kernels (lines of 1024, 4096, 32768 samples)
using PlanL512 = Plan<512, 64, 8, 8, 8>;
using PlanL2048 = Plan<2048, 128, 16, 16, 8>;
using PlanL16384 = Plan<16384, 1024, 16, 16, 16, 4>;

}  // namespace fast
}  // namespace ssw
