// Line-tile kernels of the full-frame separable DCT (forward = DCT-II, inverse = DCT-III).
//
// Replaces /root/reference/src/dct2d.rs:129-206 (per-line gather -> rustdct -> scaled scatter),
// with the colour conversion of /root/reference/src/yiq.rs:177-197 fused into the row passes:
//
//   row_fwd   : RGB8 | RGB32F | plane rows  -> Y -> DCT-II along x   -> coefficient plane
//   col_fwd   : plane (in place)            ->      DCT-II along y
//   col_inv   : plane (in place)            ->      DCT-III along y
//   row_inv   : plane rows -> DCT-III along x -> (x 4/(W*H)) -> Y' + I,Q(original RGB) -> RGB8 | RGB32F | plane
//
// One CTA owns a tile of P "line pairs": two real lines are packed as one complex line so a single
// complex FFT transforms both.  The bodies are __host__ __device__ and take
// (tid, nthreads, tile) explicitly so tests/emul can run them with std::thread + std::barrier.
#pragma once
#include "color.cuh"

#if defined(__CUDA_ARCH__)
#define SSW_SYNC() __syncthreads()
#elif defined(__CUDACC__)
#define SSW_SYNC() ((void)0)  // host pass of nvcc: libssw never runs the bodies on the CPU
#else
namespace ssw { void host_barrier(); }  // tests/emul only (g++ build)
#define SSW_SYNC() ::ssw::host_barrier()
#endif

namespace ssw {

enum { PIX_RGB8 = 0, PIX_RGB32F = 1, PIX_PLANE = 2 };

struct LineArgs {
    DctPlanDev plan;   // plan for the line length of this pass
    int w, h;          // frame size
    int P;             // line pairs per tile
    const void* src;   // row_fwd: pixels / plane; row_inv DST=RGB*: original pixels (for I,Q)
    float* plane;      // coefficient plane [h][w]
    void* dst;         // row_inv destination
    float scale0;      // forward: extra factor for k == 0 (DCT2Orthogonal); inverse: output scale
    float scalen;      // forward: extra factor for k  > 0
    long long src_stride, plane_stride, dst_stride;  // per-image strides (elements) for batched launches
    int tiles_per_image;
};

// ------------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
#define SSW_STAGE_FN __host__ __device__ __noinline__
#else
#define SSW_STAGE_FN inline
#endif

// Each stage is its own (non-inlined) function: one giant inlined switch made ptxas allocate for the
// union of all radices (254 registers + spills); separately every stage needs <= 56.
template <int R>
SSW_STAGE_FN void run_stage(cplx* s, int n, int tp, int ns, unsigned ns_magic, const cplx* tw, int tpr, bool active) {
    StageRegs<R> rg;
    if (active) stage_load<R>(s, n, tpr, tp, rg);
    SSW_SYNC();
    if (active) stage_store<R>(s, n, ns, ns_magic, tw, tpr, tp, rg);
    SSW_SYNC();
}

SSW_STAGE_FN void run_generic_stage(cplx* s, int n, int tp, int radix, int ns, const cplx* wn, int tpr, bool active) {
    GenericRegs rg;
    if (active) gstage_compute(s, n, radix, ns, wn, tpr, tp, rg);
    SSW_SYNC();
    if (active) gstage_store(s, n, tpr, tp, rg);
    SSW_SYNC();
}

// complex FFT of P line pairs resident in shared memory (in place); all threads must call
SSW_HD void fft_tile(cplx* smem, const DctPlanDev& pl, int P, int tid, int nthreads) {
    const int G = nthreads / pl.tp;       // line pairs transformed concurrently
    const int g = tid / pl.tp;
    const int tpr = tid - g * pl.tp;
    const int n = pl.n, tp = pl.tp;
    for (int pg = 0; pg < P; pg += G) {
        const int p = pg + g;
        const bool active = (g < G) && (p < P);
        cplx* s = smem + (size_t)p * pl.npad;
        for (int st = 0; st < pl.nstages; ++st) {
            const int radix = pl.stages[st].radix, ns = pl.stages[st].ns;
            if (pl.stages[st].generic) { run_generic_stage(s, n, tp, radix, ns, pl.wn, tpr, active); continue; }
            const unsigned mg = pl.ns_magic[st];
            const cplx* tw = pl.stage_tw + pl.stages[st].tw_offset;
            switch (radix) {
                case 2: run_stage<2>(s, n, tp, ns, mg, tw, tpr, active); break;
                case 3: run_stage<3>(s, n, tp, ns, mg, tw, tpr, active); break;
                case 4: run_stage<4>(s, n, tp, ns, mg, tw, tpr, active); break;
                case 5: run_stage<5>(s, n, tp, ns, mg, tw, tpr, active); break;
                case 6: run_stage<6>(s, n, tp, ns, mg, tw, tpr, active); break;
                case 8: run_stage<8>(s, n, tp, ns, mg, tw, tpr, active); break;
                case 9: run_stage<9>(s, n, tp, ns, mg, tw, tpr, active); break;
                case 10: run_stage<10>(s, n, tp, ns, mg, tw, tpr, active); break;
                case 12: run_stage<12>(s, n, tp, ns, mg, tw, tpr, active); break;
                case 15: run_stage<15>(s, n, tp, ns, mg, tw, tpr, active); break;
                default: run_stage<16>(s, n, tp, ns, mg, tw, tpr, active); break;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pixel access
// ------------------------------------------------------------------------------------------------
template <int SRC>
SSW_HD void load_rgb(const void* src, long long pix, float& r, float& g, float& b) {
    if constexpr (SRC == PIX_RGB8) {
        const unsigned char* p = (const unsigned char*)src + 3 * pix;
        r = u8_to_unit(p[0]); g = u8_to_unit(p[1]); b = u8_to_unit(p[2]);
    } else {
        const float* p = (const float*)src + 3 * pix;
        r = p[0]; g = p[1]; b = p[2];
    }
}

template <int SRC>
SSW_HD float load_luma(const void* src, long long pix) {
    if constexpr (SRC == PIX_PLANE) {
        return ((const float*)src)[pix];
    } else {
        float r, g, b;
        load_rgb<SRC>(src, pix, r, g, b);
        return rgb_to_y(r, g, b);
    }
}

// ------------------------------------------------------------------------------------------------
// forward row pass: tile = rows [2P*tile, 2P*(tile+1))
// ------------------------------------------------------------------------------------------------
template <int SRC>
SSW_HD void row_fwd_body(const LineArgs& a, cplx* smem, int tile, int tid, int nthreads) {
    const DctPlanDev& pl = a.plan;
    const int n = a.w;
    const int img = tile / a.tiles_per_image;
    const int row0 = (tile - img * a.tiles_per_image) * 2 * a.P;
    const void* src = (SRC == PIX_RGB8) ? (const void*)((const unsigned char*)a.src + 3 * img * a.src_stride)
                    : (SRC == PIX_RGB32F) ? (const void*)((const float*)a.src + 3 * img * a.src_stride)
                                          : (const void*)((const float*)a.src + img * a.src_stride);
    float* plane = a.plane + img * a.plane_stride;
    for (int p = 0; p < a.P; ++p) {
        const int ra = row0 + 2 * p, rb = ra + 1;
        cplx* s = smem + (size_t)p * pl.npad;
        if (ra >= a.h) {  // whole pair beyond the frame: keep the FFT input finite
            for (int m = tid; m < n; m += nthreads) s[padi(m)] = mk(0.f, 0.f);
            continue;
        }
        for (int m = tid; m < n; m += nthreads) {
            const float ya = load_luma<SRC>(src, (long long)ra * n + m);
            const float yb = (rb < a.h) ? load_luma<SRC>(src, (long long)rb * n + m) : 0.f;
            s[padi(makhoul(m, n))] = mk(ya, yb);
        }
    }
    SSW_SYNC();
    fft_tile(smem, pl, a.P, tid, nthreads);
    const int half = n >> 1;
    for (int p = 0; p < a.P; ++p) {
        const int ra = row0 + 2 * p, rb = ra + 1;
        if (ra >= a.h) break;
        const cplx* s = smem + (size_t)p * pl.npad;
        float* oa = plane + (long long)ra * n;
        float* ob = plane + (long long)rb * n;
        const bool hb = rb < a.h;
        for (int k = tid; k <= half; k += nthreads) {
            const int kr = k ? n - k : 0;
            float xa, xb, ya, yb;
            dct2_post(s[padi(k)], s[padi(kr)], SSW_LDG(&pl.t4[k]), xa, xb, ya, yb);
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

// ------------------------------------------------------------------------------------------------
// forward column pass, in place: tile = columns [2P*tile, 2P*(tile+1)); a float2 of two adjacent
// columns in one row *is* one complex sample of the packed line pair.
// ------------------------------------------------------------------------------------------------
SSW_HD void col_fwd_body(const LineArgs& a, cplx* smem, int tile, int tid, int nthreads) {
    const DctPlanDev& pl = a.plan;
    const int n = a.h, w = a.w, P = a.P;
    const int img = tile / a.tiles_per_image;
    const int c0 = (tile - img * a.tiles_per_image) * 2 * P;
    float* plane = a.plane + img * a.plane_stride;
    const bool vec = ((w & 1) == 0) && ((((size_t)plane) & 7) == 0);
    for (int e = tid; e < n * P; e += nthreads) {
        const int r = e / P, p = e - r * P;
        const int c = c0 + 2 * p;
        cplx v = mk(0.f, 0.f);
        const float* q = plane + (long long)r * w + c;
        if (c + 1 < w) {
            if (vec) v = *(const cplx*)q; else v = mk(q[0], q[1]);
        } else if (c < w) {
            v = mk(q[0], 0.f);
        }
        smem[(size_t)p * pl.npad + padi(makhoul(r, n))] = v;
    }
    SSW_SYNC();
    fft_tile(smem, pl, P, tid, nthreads);
    const int half = n >> 1;
    for (int e = tid; e < (half + 1) * P; e += nthreads) {
        const int k = e / P, p = e - k * P;
        const int c = c0 + 2 * p;
        if (c >= w) continue;
        const cplx* s = smem + (size_t)p * pl.npad;
        const int kr = k ? n - k : 0;
        float xa, xb, ya, yb;
        dct2_post(s[padi(k)], s[padi(kr)], SSW_LDG(&pl.t4[k]), xa, xb, ya, yb);
        const float sk = k ? a.scalen : a.scale0;
        float* q = plane + (long long)k * w + c;
        if (c + 1 < w) {
            if (vec) *(cplx*)q = mk(xa * sk, xb * sk); else { q[0] = xa * sk; q[1] = xb * sk; }
        } else {
            q[0] = xa * sk;
        }
        if (k && kr != k) {
            float* qr = plane + (long long)kr * w + c;
            if (c + 1 < w) {
                if (vec) *(cplx*)qr = mk(ya * a.scalen, yb * a.scalen); else { qr[0] = ya * a.scalen; qr[1] = yb * a.scalen; }
            } else {
                qr[0] = ya * a.scalen;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// inverse column pass, in place
// ------------------------------------------------------------------------------------------------
SSW_HD void col_inv_body(const LineArgs& a, cplx* smem, int tile, int tid, int nthreads) {
    const DctPlanDev& pl = a.plan;
    const int n = a.h, w = a.w, P = a.P;
    const int img = tile / a.tiles_per_image;
    const int c0 = (tile - img * a.tiles_per_image) * 2 * P;
    float* plane = a.plane + img * a.plane_stride;
    const bool vec = ((w & 1) == 0) && ((((size_t)plane) & 7) == 0);
    const int half = n >> 1;
    for (int e = tid; e < (half + 1) * P; e += nthreads) {
        const int k = e / P, p = e - k * P;
        const int c = c0 + 2 * p;
        const int kr = k ? n - k : 0;
        cplx pv = mk(0.f, 0.f), qv = mk(0.f, 0.f);
        if (c < w) {
            const float* q = plane + (long long)k * w + c;
            if (c + 1 < w) { if (vec) pv = *(const cplx*)q; else pv = mk(q[0], q[1]); } else pv = mk(q[0], 0.f);
            if (k) {
                const float* qr = plane + (long long)kr * w + c;
                if (c + 1 < w) { if (vec) qv = *(const cplx*)qr; else qv = mk(qr[0], qr[1]); } else qv = mk(qr[0], 0.f);
            }
        }
        cplx zk, zr;
        dct3_pre(pv.x, pv.y, qv.x, qv.y, SSW_LDG(&pl.t4[k]), zk, zr);
        cplx* s = smem + (size_t)p * pl.npad;
        s[padi(k)] = zk;
        if (k && kr != k) s[padi(kr)] = zr;
    }
    SSW_SYNC();
    fft_tile(smem, pl, P, tid, nthreads);
    for (int e = tid; e < n * P; e += nthreads) {
        const int r = e / P, p = e - r * P;
        const int c = c0 + 2 * p;
        if (c >= w) continue;
        const cplx f = smem[(size_t)p * pl.npad + padi(makhoul(r, n))];
        float* q = plane + (long long)r * w + c;
        const float va = f.x * a.scale0, vb = -f.y * a.scale0;
        if (c + 1 < w) { if (vec) *(cplx*)q = mk(va, vb); else { q[0] = va; q[1] = vb; } } else q[0] = va;
    }
}

// ------------------------------------------------------------------------------------------------
// inverse row pass with the YIQ -> RGB conversion fused (I,Q recomputed from the original pixels)
// ------------------------------------------------------------------------------------------------
template <int DST, int SRC>
SSW_HD void store_pixel(const LineArgs& a, int img, long long pix, float y) {
    if constexpr (DST == PIX_PLANE) {
        ((float*)a.dst + img * a.dst_stride)[pix] = y;
    } else {
        const void* src = (SRC == PIX_RGB8) ? (const void*)((const unsigned char*)a.src + 3 * img * a.src_stride)
                                            : (const void*)((const float*)a.src + 3 * img * a.src_stride);
        float r, g, b;
        load_rgb<SRC>(src, pix, r, g, b);
        const float i = rgb_to_i(r, g, b), q = rgb_to_q(r, g, b);
        yiq_to_rgb(y, i, q, r, g, b);
        if constexpr (DST == PIX_RGB8) {
            unsigned char* o = (unsigned char*)a.dst + 3 * (img * a.dst_stride + pix);
            o[0] = (unsigned char)unit_to_u8(r); o[1] = (unsigned char)unit_to_u8(g); o[2] = (unsigned char)unit_to_u8(b);
        } else {
            float* o = (float*)a.dst + 3 * (img * a.dst_stride + pix);
            o[0] = r; o[1] = g; o[2] = b;
        }
    }
}

template <int DST, int SRC>
SSW_HD void row_inv_body(const LineArgs& a, cplx* smem, int tile, int tid, int nthreads) {
    const DctPlanDev& pl = a.plan;
    const int n = a.w;
    const int img = tile / a.tiles_per_image;
    const int row0 = (tile - img * a.tiles_per_image) * 2 * a.P;
    const float* plane = a.plane + img * a.plane_stride;
    const int half = n >> 1;
    for (int p = 0; p < a.P; ++p) {
        const int ra = row0 + 2 * p, rb = ra + 1;
        cplx* s = smem + (size_t)p * pl.npad;
        const bool ha = ra < a.h, hb = rb < a.h;
        const float* ia = plane + (long long)ra * n;
        const float* ib = plane + (long long)rb * n;
        for (int k = tid; k <= half; k += nthreads) {
            const int kr = k ? n - k : 0;
            const float pa = ha ? ia[k] : 0.f, pb = hb ? ib[k] : 0.f;
            const float qa = (k && ha) ? ia[kr] : 0.f, qb = (k && hb) ? ib[kr] : 0.f;
            cplx zk, zr;
            dct3_pre(pa, pb, qa, qb, SSW_LDG(&pl.t4[k]), zk, zr);
            s[padi(k)] = zk;
            if (k && kr != k) s[padi(kr)] = zr;
        }
    }
    SSW_SYNC();
    fft_tile(smem, pl, a.P, tid, nthreads);
    for (int p = 0; p < a.P; ++p) {
        const int ra = row0 + 2 * p, rb = ra + 1;
        if (ra >= a.h) break;
        const cplx* s = smem + (size_t)p * pl.npad;
        const bool hb = rb < a.h;
        for (int m = tid; m < n; m += nthreads) {
            const cplx f = s[padi(makhoul(m, n))];
            store_pixel<DST, SRC>(a, img, (long long)ra * n + m, f.x * a.scale0);
            if (hb) store_pixel<DST, SRC>(a, img, (long long)rb * n + m, -f.y * a.scale0);
        }
    }
}

#if defined(__CUDACC__)
extern __shared__ __align__(16) unsigned char ssw_dyn_smem[];

// 1024 threads/CTA must be launchable (16384-point lines), which caps the kernels at 64 registers;
// every stage function needs <= 56.  __grid_constant__: the plan is indexed dynamically, keep it in
// the constant bank instead of a per-thread local copy.
#define SSW_LINE_KERNEL __global__ void __launch_bounds__(1024, 1)
template <int SRC>
SSW_LINE_KERNEL row_fwd_kernel(const __grid_constant__ LineArgs a) { row_fwd_body<SRC>(a, (cplx*)ssw_dyn_smem, blockIdx.x, threadIdx.x, blockDim.x); }
SSW_LINE_KERNEL col_fwd_kernel(const __grid_constant__ LineArgs a) { col_fwd_body(a, (cplx*)ssw_dyn_smem, blockIdx.x, threadIdx.x, blockDim.x); }
SSW_LINE_KERNEL col_inv_kernel(const __grid_constant__ LineArgs a) { col_inv_body(a, (cplx*)ssw_dyn_smem, blockIdx.x, threadIdx.x, blockDim.x); }
template <int DST, int SRC>
SSW_LINE_KERNEL row_inv_kernel(const __grid_constant__ LineArgs a) { row_inv_body<DST, SRC>(a, (cplx*)ssw_dyn_smem, blockIdx.x, threadIdx.x, blockDim.x); }
#endif

}  // namespace ssw
