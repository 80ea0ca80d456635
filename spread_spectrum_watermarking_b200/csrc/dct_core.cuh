// Shared-memory mixed-radix Stockham FFT + the DCT-II / DCT-III pre/post passes.
//
// Everything in this header is `__host__ __device__` and takes the thread id explicitly so that
// the exact same index arithmetic can be executed on the CPU by tests/emul/dct_emul.cpp (the
// build container has no GPU).  The __global__ wrappers live in dct_kernels.cuh.
//
// Replaces the arithmetic of rustdct 0.7.0 `process_dct2_with_scratch` / `process_dct3_with_scratch`
// as called from /root/reference/src/dct2d.rs:141-145,181-185, including the driver's scalings
// (x2 per forward pass :107-108,166,202; x0.5 per inverse pass :109).
#pragma once
#include "dct_plan.h"

#if defined(__CUDACC__)
#define SSW_HD __host__ __device__ __forceinline__
typedef float2 cplx;
#else
#define SSW_HD inline
struct cplx { float x, y; };
#endif

#if defined(__CUDA_ARCH__)
#define SSW_LDG(p) __ldg(p)
#else
#define SSW_LDG(p) (*(p))
#endif

namespace ssw {

struct DctPlanDev {
    int n, npad, tp, nstages;
    DctStage stages[kMaxStages];
    unsigned ns_magic[kMaxStages];  // floor(2^32/ns)+1 : j/ns == umulhi(j, magic) for j,ns < 2^16
    const cplx* stage_tw;
    const cplx* wn;
    const cplx* t4;
};

SSW_HD cplx mk(float x, float y) { cplx c; c.x = x; c.y = y; return c; }
// Complex arithmetic.  On the device every operation is written with the packed FP32 instructions of sm_100
// (add/mul/fma.rn.f32x2 -> FADD2 / FMUL2 / FFMA2): a complex number is one 64-bit register pair, ptxas folds the
// lane swaps / per-lane negations / scalar broadcasts below into operand modifiers (R.F32x2.LO_HI.NP, R.F32),
// so a complex add is ONE instruction and a complex multiply TWO (instead of 2 and 4).  Same flop rate as the
// scalar forms, half the issue slots -- the line kernels are issue-bound, not FMA-pipe-bound.
// The host forms (tests/emul, built with -ffp-contract=off) round every product separately; the device forms
// round a.x*b first and fuse the second product -- both are within the 1e-5 coefficient tolerance of the oracle.
#if defined(__CUDA_ARCH__)
#define SSW_PACKED 1
SSW_HD cplx cadd(cplx a, cplx b) { return __fadd2_rn(a, b); }
SSW_HD cplx csub(cplx a, cplx b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
SSW_HD cplx cbc(float v) { return make_float2(v, v); }                    // broadcast
SSW_HD cplx cswap_np(cplx a) { return make_float2(-a.y, a.x); }           // a * (+i)
SSW_HD cplx cswap_pn(cplx a) { return make_float2(a.y, -a.x); }           // a * (-i)
SSW_HD cplx cscale2(cplx a, float f) { return __fmul2_rn(a, cbc(f)); }
SSW_HD cplx cfma_s(cplx a, float f, cplx c) { return __ffma2_rn(a, cbc(f), c); }   // a*f + c
SSW_HD cplx cmul_lanes(cplx a, float fx, float fy) { return __fmul2_rn(a, make_float2(fx, fy)); }   // (a.x*fx, a.y*fy)
// the same with the products rounded on their own whatever follows (nz = -0.0f at run time: ptxas contracts FMUL2 + FADD2)
SSW_HD cplx cmul_lanes_x(cplx a, float fx, float fy, float nz) { return __ffma2_rn(a, make_float2(fx, fy), make_float2(nz, nz)); }
// a * b = b.x * a + b.y * (i a)
SSW_HD cplx cmul(cplx a, cplx b) { return __ffma2_rn(cswap_np(a), cbc(b.y), __fmul2_rn(a, cbc(b.x))); }
SSW_HD cplx mul_mi(cplx a) { return cswap_pn(a); }  // a * (-i)
SSW_HD cplx mul_pi(cplx a) { return cswap_np(a); }  // a * (+i)
#else
SSW_HD cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
SSW_HD cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
SSW_HD cplx cscale2(cplx a, float f) { return mk(a.x * f, a.y * f); }
SSW_HD cplx cfma_s(cplx a, float f, cplx c) { return mk(a.x * f + c.x, a.y * f + c.y); }
SSW_HD cplx cmul_lanes(cplx a, float fx, float fy) { return mk(a.x * fx, a.y * fy); }
SSW_HD cplx cmul_lanes_x(cplx a, float fx, float fy, float nz) { (void)nz; return mk(a.x * fx, a.y * fy); }
SSW_HD cplx cmul(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
SSW_HD cplx mul_mi(cplx a) { return mk(a.y, -a.x); }  // a * (-i)
SSW_HD cplx mul_pi(cplx a) { return mk(-a.y, a.x); }  // a * (+i)
#endif
SSW_HD int padi(int a) { return a + (a >> 5); }

SSW_HD unsigned fastdiv(unsigned j, unsigned magic) {
#if defined(__CUDA_ARCH__)
    return __umulhi(j, magic);
#else
    return (unsigned)(((unsigned long long)j * magic) >> 32);
#endif
}

// Makhoul reordering: line position m -> FFT input index
SSW_HD int makhoul(int m, int n) { return (m & 1) ? (n - 1 - (m >> 1)) : (m >> 1); }

// ------------------------------------------------------------------------------------------------
// compile-time twiddles exp(-2*pi*i*M/R) for the in-register composite butterflies
// ------------------------------------------------------------------------------------------------
template <int R, int M>
SSW_HD cplx tw_mul(cplx v) {
    constexpr int m = ((M % R) + R) % R;
    if constexpr (m == 0) return v;
    else if constexpr (4 * m == R) return mul_mi(v);
    else if constexpr (2 * m == R) return mk(-v.x, -v.y);
    else if constexpr (4 * m == 3 * R) return mul_pi(v);
    else if constexpr (8 * m == R) { const float h = 0.70710678118654752440f; return cscale2(cadd(v, mul_mi(v)), h); }        // (1-i)/sqrt2
    else if constexpr (8 * m == 3 * R) { const float h = 0.70710678118654752440f; return cscale2(csub(mul_mi(v), v), h); }   // (-1-i)/sqrt2
    else if constexpr (R == 16 && m == 1) return cmul(v, mk(0.92387953251128675613f, -0.38268343236508977173f));
    else if constexpr (R == 16 && m == 3) return cmul(v, mk(0.38268343236508977173f, -0.92387953251128675613f));
    else if constexpr (R == 16 && m == 9) return cmul(v, mk(-0.92387953251128675613f, 0.38268343236508977173f));
    else if constexpr (R == 9 && m == 1) return cmul(v, mk(0.76604444311897803520f, -0.64278760968653932632f));
    else if constexpr (R == 9 && m == 2) return cmul(v, mk(0.17364817766693034885f, -0.98480775301220805937f));
    else if constexpr (R == 9 && m == 4) return cmul(v, mk(-0.93969262078590838405f, -0.34202014332566873304f));
    else { static_assert(R < 0, "twiddle constant not tabulated"); return v; }
}

// ------------------------------------------------------------------------------------------------
// in-register DFTs, natural order in and out, forward sign exp(-2*pi*i*nk/R)
// ------------------------------------------------------------------------------------------------
template <int R> struct Dft;

template <> struct Dft<2> {
    static SSW_HD void run(cplx* x) { cplx a = x[0], b = x[1]; x[0] = cadd(a, b); x[1] = csub(a, b); }
};
template <> struct Dft<3> {
    static SSW_HD void run(cplx* x) {
        const float s = 0.86602540378443864676f;
        cplx t = cadd(x[1], x[2]);
        cplx d = csub(x[1], x[2]);
        cplx m1 = cfma_s(t, -0.5f, x[0]);
        cplx e = cscale2(mul_mi(d), s);  // (-i*s)*d
        x[0] = cadd(x[0], t);
        x[1] = cadd(m1, e);
        x[2] = csub(m1, e);
    }
};
template <> struct Dft<4> {
    static SSW_HD void run(cplx* x) {
        cplx a = cadd(x[0], x[2]), b = csub(x[0], x[2]);
        cplx c = cadd(x[1], x[3]), d = mul_mi(csub(x[1], x[3]));
        x[0] = cadd(a, c); x[2] = csub(a, c);
        x[1] = cadd(b, d); x[3] = csub(b, d);
    }
};
template <> struct Dft<5> {
    static SSW_HD void run(cplx* x) {
        const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
        const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
        cplx a1 = cadd(x[1], x[4]), a2 = cadd(x[2], x[3]);
        cplx b1 = csub(x[1], x[4]), b2 = csub(x[2], x[3]);
        cplx r1 = cfma_s(a2, c2, cfma_s(a1, c1, x[0]));
        cplx r2 = cfma_s(a2, c1, cfma_s(a1, c2, x[0]));
        cplx i1 = cfma_s(b2, s2, cscale2(b1, s1));
        cplx i2 = cfma_s(b2, -s1, cscale2(b1, s2));
        x[0] = cadd(x[0], cadd(a1, a2));
        x[1] = cadd(r1, mul_mi(i1));
        x[4] = csub(r1, mul_mi(i1));
        x[2] = cadd(r2, mul_mi(i2));
        x[3] = csub(r2, mul_mi(i2));
    }
};

template <int N> struct IC { static constexpr int value = N; };
template <int N, int I = 0, class F>
SSW_HD void static_for(F&& f) {
    if constexpr (I < N) { f(IC<I>{}); static_for<N, I + 1>(f); }
}

// Cooley-Tukey A x B (twiddled); n = B*n1 + n2, k = k1 + A*k2
template <int A, int B> struct DftCT {
    static SSW_HD void run(cplx* x) {
        constexpr int R = A * B;
        cplx y[R];
        static_for<B>([&](auto n2c) {
            constexpr int n2 = decltype(n2c)::value;
            cplx t[A];
#pragma unroll
            for (int n1 = 0; n1 < A; ++n1) t[n1] = x[B * n1 + n2];
            Dft<A>::run(t);
            static_for<A>([&](auto k1c) {
                constexpr int k1 = decltype(k1c)::value;
                y[n2 * A + k1] = tw_mul<R, n2 * k1>(t[k1]);
            });
        });
#pragma unroll
        for (int k1 = 0; k1 < A; ++k1) {
            cplx t[B];
#pragma unroll
            for (int n2 = 0; n2 < B; ++n2) t[n2] = y[n2 * A + k1];
            Dft<B>::run(t);
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) x[k1 + A * k2] = t[k2];
        }
    }
};

SSW_HD constexpr int modinv(int a, int m) {
    for (int i = 1; i < m; ++i) if ((a * i) % m == 1) return i;
    return 1;
}

// Good-Thomas prime-factor A x B (gcd(A,B)=1, no twiddles):
//   input  n = (B*n1 + A*n2) mod R,  output k = (B*inv(B,A)*k1 + A*inv(A,B)*k2) mod R
template <int A, int B> struct DftPFA {
    static SSW_HD void run(cplx* x) {
        constexpr int R = A * B;
        constexpr int EA = B * modinv(B % A, A), EB = A * modinv(A % B, B);
        cplx y[R];
#pragma unroll
        for (int n2 = 0; n2 < B; ++n2) {
            cplx t[A];
#pragma unroll
            for (int n1 = 0; n1 < A; ++n1) t[n1] = x[(B * n1 + A * n2) % R];
            Dft<A>::run(t);
#pragma unroll
            for (int k1 = 0; k1 < A; ++k1) y[n2 * A + k1] = t[k1];
        }
#pragma unroll
        for (int k1 = 0; k1 < A; ++k1) {
            cplx t[B];
#pragma unroll
            for (int n2 = 0; n2 < B; ++n2) t[n2] = y[n2 * A + k1];
            Dft<B>::run(t);
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) x[(EA * k1 + EB * k2) % R] = t[k2];
        }
    }
};

template <> struct Dft<6> { static SSW_HD void run(cplx* x) { DftPFA<3, 2>::run(x); } };
template <> struct Dft<8> { static SSW_HD void run(cplx* x) { DftCT<4, 2>::run(x); } };
template <> struct Dft<9> { static SSW_HD void run(cplx* x) { DftCT<3, 3>::run(x); } };
template <> struct Dft<10> { static SSW_HD void run(cplx* x) { DftPFA<5, 2>::run(x); } };
template <> struct Dft<12> { static SSW_HD void run(cplx* x) { DftPFA<4, 3>::run(x); } };
template <> struct Dft<15> { static SSW_HD void run(cplx* x) { DftPFA<5, 3>::run(x); } };
template <> struct Dft<16> { static SSW_HD void run(cplx* x) { DftCT<4, 4>::run(x); } };

// ------------------------------------------------------------------------------------------------
// one Stockham stage, in place: phase A loads the butterfly inputs into registers, (barrier),
// phase B twiddles, transforms and stores.  `tpr` = rank of this thread inside the line pair's team.
// ------------------------------------------------------------------------------------------------
template <int R> struct StageRegs { static constexpr int MAXIT = (R >= 9) ? 1 : 16 / R; cplx v[MAXIT * R]; };

template <int R>
SSW_HD void stage_load(const cplx* s, int n, int tpr, int tp, StageRegs<R>& rg) {
    const int nb = n / R;
#pragma unroll
    for (int it = 0; it < StageRegs<R>::MAXIT; ++it) {
        const int j = tpr + it * tp;
        if (j < nb) {
#pragma unroll
            for (int r = 0; r < R; ++r) rg.v[it * R + r] = s[padi(j + r * nb)];
        }
    }
}

template <int R>
SSW_HD void stage_store(cplx* s, int n, int ns, unsigned ns_magic, const cplx* tw, int tpr, int tp, StageRegs<R>& rg) {
    const int nb = n / R;
#pragma unroll
    for (int it = 0; it < StageRegs<R>::MAXIT; ++it) {
        const int j = tpr + it * tp;
        if (j < nb) {
            cplx* x = &rg.v[it * R];
            int k = 0;
            if (ns > 1) {
                const int blk = (int)fastdiv((unsigned)j, ns_magic);
                k = j - blk * ns;
#pragma unroll
                for (int r = 1; r < R; ++r) x[r] = cmul(x[r], SSW_LDG(&tw[(r - 1) * ns + k]));
            }
            Dft<R>::run(x);
            const int j0 = (j - k) * R + k;
#pragma unroll
            for (int r = 0; r < R; ++r) s[padi(j0 + r * ns)] = x[r];
        }
    }
}

// generic prime-radix stage: each thread owns up to kGenericOutMax output elements
struct GenericRegs { cplx acc[kGenericOutMax]; };

SSW_HD void gstage_compute(const cplx* s, int n, int p, int ns, const cplx* wn, int tpr, int tp, GenericRegs& rg) {
    const int nb = n / p;
    const int span = ns * p;
    const int f1 = n / span;
#pragma unroll 1
    for (int o = 0; o < kGenericOutMax; ++o) {
        const int e = tpr + o * tp;
        if (e < n) {
            const int blk = e / span, rem = e - blk * span;
            const int m = rem / ns, k = rem - m * ns;
            const int j = blk * ns + k;
            const int step = (int)(((long long)k * f1 + (long long)m * nb) % n);
            int idx = 0;
            cplx a = mk(0.f, 0.f);
            for (int q = 0; q < p; ++q) {
                a = cadd(a, cmul(s[padi(j + q * nb)], SSW_LDG(&wn[idx])));
                idx += step;
                if (idx >= n) idx -= n;
            }
            rg.acc[o] = a;
        }
    }
}

SSW_HD void gstage_store(cplx* s, int n, int tpr, int tp, const GenericRegs& rg) {
#pragma unroll 1
    for (int o = 0; o < kGenericOutMax; ++o) {
        const int e = tpr + o * tp;
        if (e < n) s[padi(e)] = rg.acc[o];
    }
}

// ------------------------------------------------------------------------------------------------
// DCT-II post pass: from Z = FFT(perm(A) + i perm(B)) to the two spectra (scipy scaling, i.e. the
// reference's rustdct result x2).  Handles k and n-k together.
//   out: xa_k, xb_k (position k) and xa_r, xb_r (position n-k; only valid if 0 < k and k != n-k)
// ------------------------------------------------------------------------------------------------
SSW_HD void dct2_post(cplx zk, cplx zr, cplx t, float& xa_k, float& xb_k, float& xa_r, float& xb_r) {
    // with S = Z[k] + conj(Z[n-k]) = (sr, si), D = Z[k] - conj(Z[n-k]) = (dr, di):
    //   P = Z[k] + Z[n-k] = (sr, di),  Q = Z[k] - Z[n-k] = (dr, si)
    //   (xa_k, xb_k) = (Re(t S), Im(t D))                     =  t.x P + t.y (i Q)
    //   (xa_r, xb_r) = (Re(t' conj S), Im(t' (-conj D)))      = -t.y P + t.x (i Q),   t' = t_{n-k} = -i conj(t_k)
    const cplx p = cadd(zk, zr), iq = mul_pi(csub(zk, zr));
    const cplx a = cfma_s(iq, t.y, cscale2(p, t.x));
    const cplx b = cfma_s(iq, t.x, cscale2(p, -t.y));
    xa_k = a.x; xb_k = a.y; xa_r = b.x; xb_r = b.y;
}

// DCT-III pre pass: from the two coefficient lines to conj(Z) (so that a *forward* FFT yields
// conj of the inverse), including the reference's 0.5 per pass (and the 1/2 of the Hermitian split).
//   pa,pb = lines A,B at k;  qa,qb = lines A,B at n-k (0 when k == 0)
//   zk -> index k, zr -> index n-k
SSW_HD void dct3_pre(float pa, float pb, float qa, float qb, cplx t, cplx& zk, cplx& zr) {
    // conjZ[k]   = 1/4 t_k     ((pa+qb) - i (pb-qa)) = t_k     * 1/4 (conj(p) + swap(q))
    // conjZ[n-k] = 1/4 t_{n-k} ((qa+pb) - i (qb-pa)) = t_{n-k} * 1/4 (conj(q) + swap(p)),  t_{n-k} = (-t.y, -t.x)
    const cplx u = cscale2(cadd(mk(pa, -pb), mk(qb, qa)), 0.25f);
    const cplx v = cscale2(cadd(mk(qa, -qb), mk(pb, pa)), 0.25f);
    zk = cmul(u, t);
    zr = cmul(v, mk(-t.y, -t.x));
}

}  // namespace ssw
