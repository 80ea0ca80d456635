// Embed inverse as a low-rank update (fused embed pipeline).
//
// The reference inverts the WHOLE modified coefficient plane (Writer::result, /root/reference/src/algorithm.rs:361-379:
// DCT-III per pass x0.5, then x4/(W*H), /root/reference/src/dct2d.rs:107-111,213-217).  The modified plane differs from
// the forward transform of the original frame in only k coefficients, and the transform is linear:
//     Y' = IDCT(C + D) = Y + IDCT(D),   IDCT(D)[r][c] = 1/(W H) * sum_j D_j a(u_j) a(v_j) cos(pi u_j (2r+1) / 2H) cos(pi v_j (2c+1) / 2W)
// (a(0) = 1/2, a(n) = 1; (u_j, v_j) = row / column of the j-th ordered coefficient, D_j = f(c_j, w_j) - c_j).
// Natural images keep their k largest coefficients in a few dozen low-frequency rows, so IDCT(D) = CY^T (D CX) is a
// product with inner dimension Kr = 1 + max_j u_j (~60 for the 4K / 1080p frames):
//     lowrank_rows   T[u][c]  = sum_{j: u_j = u} D_j a(v_j)/(W H) cos(pi v_j (2c+1) / 2W)          (Kr x W, written over the plane)
//     lowrank_apply  out[r][c] = RGB8( Y(orig pixel) + sum_u a(u) cos(pi u (2r+1) / 2H) T[u][c], chroma(orig pixel) )
// Y is the exact luma of the original pixels (the value whose transform the reference inverts), so the result differs
// from the reference's round trip only by FP32 rounding of either path -- RGB8 within +-1 LSB at isolated ties, like any
// two correct FP32 transforms (measured: tests/test_gpu_parity.py).  Cost: Kr FMAs per pixel and one read + one write of
// the RGB8 frame instead of two full line passes; any Kr is handled (64 rows per sweep), it is simply slower for frames
// whose energy is not concentrated.
#pragma once
#include "dct_fast.cuh"
#include "select_kernels.cuh"

namespace ssw {

constexpr int kLrRows = 64;         // coefficient rows per sweep
constexpr int kLrTileX = 128, kLrTileY = 64;
constexpr unsigned kLrBad = 0xFFFFFFFFu;   // maxrow value of a frame whose ordering failed: no update, the frame stays unmarked

// cos(pi * n / (2 N)) for an integer n reduced mod 4N (exact argument reduction; cospif on [0, 1/2])
__device__ __forceinline__ float cos_quarter(unsigned n, unsigned N) {   // 0 <= n < 4N
    if (n > 2u * N) n = 4u * N - n;               // cos(2 pi - x) = cos x          -> [0, 2N]
    const bool neg = n > N;
    if (neg) n = 2u * N - n;                      // cos(pi - x) = -cos x           -> [0, N]
    const float v = cospif((float)n / (float)(2u * N));
    return neg ? -v : v;
}

// a(v) cos(pi v (2c+1) / 2n) for v < kLrTab, c < n: the cosine factors of the first kLrTab coefficient rows / columns of
// an n-point line, built once per line length (ssw_ctx caches the table); rows / columns beyond it are evaluated in place
constexpr unsigned kLrTab = 128;

__device__ __forceinline__ float lr_factor(unsigned v, unsigned c, unsigned n) {   // v, c < n <= 65535
    return (v ? 1.0f : 0.5f) * cos_quarter((v * (2u * c + 1u)) % (4u * n), n);
}

__global__ void lowrank_table_kernel(float* __restrict__ t, unsigned n) {
    const unsigned c = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (c < n) t[(size_t)v * n + c] = lr_factor(v, c, n);
}

// T rows of one 32-column strip: grid (ceil(W/32), batch), 256 threads.  idx / delta: [batch][k] (rank order).
// The changed coefficients of a sweep of 64 rows are scattered into a dense 64 x kLrTab block D in shared memory (unique
// positions: plain stores); T = D * CX is then a small dense product -- lane = column c, warp q = rows u = q (mod 8),
// 128 independent coalesced table loads per thread, broadcast reads of D.  Coefficient columns beyond the table (frames
// whose energy is not concentrated) are added entry by entry, in rank order.  No atomics: the result is deterministic.
__global__ void __launch_bounds__(256)
lowrank_rows_kernel(float* __restrict__ planes, long long plane_stride, unsigned w, unsigned h, const unsigned* __restrict__ idx,
                    const float* __restrict__ delta, unsigned k, const unsigned* __restrict__ maxrow, const float* __restrict__ cx_tab) {
    pdl_enter();
    __shared__ float dblk[kLrRows][kLrTab];       // 32 KB
    __shared__ float wide[kLrRows][33];
    __shared__ unsigned n_wide;
    extern __shared__ unsigned long long ent[];   // k entries: (row << 48 | column << 32 | scaled delta bits)
    const unsigned img = blockIdx.y, x0 = blockIdx.x * 32u;
    const unsigned mr = maxrow[img];
    if (mr == kLrBad) return;
    const unsigned kr = mr + 1u;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float scale = 1.0f / ((float)w * (float)h);
    for (unsigned j = threadIdx.x; j < k; j += 256) {
        const unsigned p = idx[(size_t)img * k + j], u = p / w, v = p - u * w;
        ent[j] = ((unsigned long long)u << 48) | ((unsigned long long)v << 32) | (unsigned long long)__float_as_uint(delta[(size_t)img * k + j] * scale);
    }
    float* plane = planes + (long long)img * plane_stride;
    const unsigned c = min(x0 + lane, w - 1u);
    for (unsigned u0 = 0; u0 < kr; u0 += kLrRows) {
        for (int i = threadIdx.x; i < kLrRows * (int)kLrTab; i += 256) (&dblk[0][0])[i] = 0.f;
        for (int i = threadIdx.x; i < kLrRows * 33; i += 256) (&wide[0][0])[i] = 0.f;
        if (threadIdx.x == 0) n_wide = 0u;
        __syncthreads();
        for (unsigned j = threadIdx.x; j < k; j += 256) {
            const unsigned long long e = ent[j];
            const unsigned ul = (unsigned)(e >> 48) - u0, v = (unsigned)(e >> 32) & 0xFFFFu;
            if (ul < (unsigned)kLrRows) {
                if (v < kLrTab) dblk[ul][v] = __uint_as_float((unsigned)e);
                else atomicAdd(&n_wide, 1u);
            }
        }
        __syncthreads();
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        const unsigned kc = min((unsigned)kLrTab, w);
#pragma unroll 4
        for (unsigned v = 0; v < kc; ++v) {
            const float f = __ldg(cx_tab + (size_t)v * w + c);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(dblk[warp + 8 * i][v], f, acc[i]);
        }
        if (n_wide) {   // coefficient columns beyond the table: entry by entry, rank order, each row by the warp that owns it
            for (unsigned j = 0; j < k; ++j) {
                const unsigned long long e = ent[j];                       // broadcast read
                const unsigned ul = (unsigned)(e >> 48) - u0, v = (unsigned)(e >> 32) & 0xFFFFu;
                if (ul >= (unsigned)kLrRows || (ul & 7u) != warp || v < kLrTab) continue;   // (uniform over the warp)
                wide[ul][lane] = fmaf(__uint_as_float((unsigned)e), lr_factor(v, c, w), wide[ul][lane]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const unsigned r = warp + 8 * i;
            if (u0 + r < kr && x0 + lane < w) plane[(size_t)(u0 + r) * w + x0 + lane] = acc[i] + wide[r][lane];
        }
        __syncthreads();
    }
}

// one thread's share of a tile fill: the CTA's 256 threads load rows8 x 32 float4 of T and rows8 x 64 CY factors; every
// thread issues ALL its loads before the first use (a loop with a load -> store chain per iteration costs one L2 round
// trip per iteration: measured 38 % of the kernel's stall samples)
__device__ __forceinline__ void lr_fetch_tile(const float* __restrict__ plane, const float* __restrict__ cy_tab, unsigned w, unsigned h,
                                              unsigned x0, unsigned y0, unsigned u0, unsigned rows, float4 (&tv)[8], float (&cv)[16]) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const unsigned i = threadIdx.x + 256u * it, r = i >> 5, q = i & 31u;
        tv[it] = (r < rows && x0 + 4 * q < w) ? __ldg((const float4*)(plane + (size_t)(u0 + r) * w + x0 + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < 16; ++it) {
        const unsigned i = threadIdx.x + 256u * it, r = i >> 6, yy = i & 63u;
        const unsigned u = u0 + r, y = min(y0 + yy, h - 1u);
        cv[it] = (r < rows && u < kLrTab) ? __ldg(cy_tab + (size_t)u * h + y) : 0.f;
    }
}
// rows beyond the table (frames whose energy is not concentrated): evaluated in place
__device__ __forceinline__ float lr_cy_slow(unsigned it, unsigned u0, unsigned y0, unsigned rows, unsigned h, float tabulated) {
    const unsigned i = threadIdx.x + 256u * it, r = i >> 6, yy = i & 63u;
    const unsigned u = u0 + r;
    return (r < rows && u >= kLrTab) ? lr_factor(u, min(y0 + yy, h - 1u), h) : tabulated;
}
// the epilogue reads the original pixels of the thread's rows: ask for their lines while the products run
__device__ __forceinline__ void lr_prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// out = RGB8(Y(orig) + CY^T T, chroma(orig)); grid (ceil(W/128), ceil(H/64), batch), 256 threads: warp = 8 rows, lane = 4 pixels
__global__ void __launch_bounds__(256)
lowrank_apply_kernel(const float* __restrict__ planes, long long plane_stride, unsigned w, unsigned h,
                     const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, const unsigned* __restrict__ maxrow,
                     const float* __restrict__ cy_tab, float nz) {
    pdl_enter();
#if defined(__CUDA_ARCH__)
    __shared__ __align__(16) float ts[kLrRows][kLrTileX];
    __shared__ __align__(16) float cy[kLrRows][kLrTileY];
    const unsigned img = blockIdx.z, x0 = blockIdx.x * kLrTileX, y0 = blockIdx.y * kLrTileY;
    const unsigned mr = maxrow[img];
    const unsigned kr = (mr == kLrBad) ? 0u : mr + 1u;
    const float* plane = planes + (long long)img * plane_stride;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    float2 acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
    if (x0 + 4 * tx < w) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const unsigned r = min(y0 + 8 * ty + i, h - 1u);
            lr_prefetch(src + 3 * ((size_t)img * w * h + (size_t)r * w + x0 + 4 * tx));
        }
    }
    for (unsigned u0 = 0; u0 < kr; u0 += kLrRows) {
        const unsigned rows = min((unsigned)kLrRows, kr - u0);
        float4 tv[8];
        float cv[16];
        lr_fetch_tile(plane, cy_tab, w, h, x0, y0, u0, rows, tv, cv);
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const unsigned i = threadIdx.x + 256u * it;
            *(float4*)&ts[i >> 5][4 * (i & 31u)] = tv[it];
        }
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            const unsigned i = threadIdx.x + 256u * it;
            cy[i >> 6][i & 63u] = (u0 + kLrRows > kLrTab) ? lr_cy_slow(it, u0, y0, rows, h, cv[it]) : cv[it];
        }
        __syncthreads();
#pragma unroll 4
        for (unsigned r = 0; r < rows; ++r) {
            const float4 t = *(const float4*)&ts[r][4 * tx];
            const float4 c0 = *(const float4*)&cy[r][8 * ty], c1 = *(const float4*)&cy[r][8 * ty + 4];
            const float2 t01 = make_float2(t.x, t.y), t23 = make_float2(t.z, t.w);
            const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[i][0] = __ffma2_rn(make_float2(cc[i], cc[i]), t01, acc[i][0]);
                acc[i][1] = __ffma2_rn(make_float2(cc[i], cc[i]), t23, acc[i][1]);
            }
        }
    }
    const unsigned x = x0 + 4 * tx;
    if (x >= w) return;
    const unsigned char* s = src + 3 * (size_t)img * w * h;
    unsigned char* d = dst + 3 * (size_t)img * w * h;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        const unsigned ra = y0 + 8 * ty + i, rb = ra + 1;
        if (ra >= h) break;
        const bool hb = rb < h;
        const size_t pa = (size_t)ra * w + x, pb = (size_t)(hb ? rb : ra) * w + x;
        const unsigned* qa = (const unsigned*)(s + 3 * pa);
        const unsigned* qb = (const unsigned*)(s + 3 * pb);
        const unsigned wa[3] = {fast::ldw(qa), fast::ldw(qa + 1), fast::ldw(qa + 2)}, wb[3] = {fast::ldw(qb), fast::ldw(qb + 1), fast::ldw(qb + 2)};
        cplx y2[4];
        fast::luma4x2_words(wa, wb, nz, y2);
        // (row A, row B) pairs of the update: acc[i] = row A, acc[i+1] = row B
        y2[0] = __fadd2_rn(y2[0], make_float2(acc[i][0].x, acc[i + 1][0].x));
        y2[1] = __fadd2_rn(y2[1], make_float2(acc[i][0].y, acc[i + 1][0].y));
        y2[2] = __fadd2_rn(y2[2], make_float2(acc[i][1].x, acc[i + 1][1].x));
        y2[3] = __fadd2_rn(y2[3], make_float2(acc[i][1].y, acc[i + 1][1].y));
        unsigned oa[3], ob[3];
        fast::rgb8_out4x2_words(wa, wb, nz, y2, oa, ob);
        unsigned* da = (unsigned*)(d + 3 * pa);
        unsigned* db = (unsigned*)(d + 3 * pb);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            da[j] = oa[j];
            if (hb) db[j] = ob[j];
        }
    }
#endif
}

// ---- the same on the tensor cores -------------------------------------------------------------------------------------
// U[64 y][128 x] = CY^T[64 y][Kr] * T[Kr][128 x] per CTA is a GEMM with a short inner dimension; this variant issues it as
// warp-level mma.sync.m16n8k8 TF32 instructions with FP32 accumulators.  TF32 keeps 10 mantissa bits, far too few for the
// +-1 LSB / <= 64-flip bound of the golden-PNG test (a plain TF32 product flips ~0.3 % of the bytes), so both operands
// are split v = hi + lo (hi = cvt.rna.tf32(v), lo = v - hi) and three products hi*hi + hi*lo + lo*hi are accumulated:
// relative error ~2^-21 per product, the accuracy class of the FP32 FMA form above.
//   warp tile 32 y x 32 x = 2 m-tiles x 4 n-tiles; per 8-deep k-step: 16 LDS.32, 32 split instructions, 24 MMAs.
//   Operands are read straight from row-major shared tiles with XOR-swizzled columns, chosen so that the four k-rows a
//   fragment load touches fall on disjoint bank sets (no padding, fills stay 128-bit row stores).
//   Column n of n-tile j of a 16-pixel group stands for pixel 4*(n/2) + 2*(j%2) + n%2: a thread's accumulators of an
//   n-tile pair are then 4 CONSECUTIVE pixels of rows g and g+8 -- exactly what the packed two-row colour helpers take.
// hi part of the operand split: the top 19 bits (what the tensor core reads of an f32 register); v - hi is exact, so
// hi + lo == v whatever the rounding of hi -- a mask (one LOP3) instead of cvt.rna.tf32 (IADD3 + LOP3 + SEL on sm_100a)
__device__ __forceinline__ unsigned tf32_hi(float v) { return __float_as_uint(v) & 0xFFFFE000u; }
__device__ __forceinline__ void mma_tf32(float* d, const unsigned* a, const unsigned* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(256)
lowrank_apply_mma_kernel(const float* __restrict__ planes, long long plane_stride, unsigned w, unsigned h,
                         const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, const unsigned* __restrict__ maxrow,
                         const float* __restrict__ cy_tab, float nz) {
    pdl_enter();
#if defined(__CUDA_ARCH__)
    __shared__ __align__(16) float ts[kLrRows][kLrTileX];   // T rows, element (k, x) at column x ^ tswz(k)
    __shared__ __align__(16) float cy[kLrRows][kLrTileY];   // CY rows, element (k, y) at column y ^ 8*(k & 3)
    const unsigned img = blockIdx.z, x0 = blockIdx.x * kLrTileX, y0 = blockIdx.y * kLrTileY;
    const unsigned mr = maxrow[img];
    const unsigned kr = (mr == kLrBad) ? 0u : mr + 1u;
    const float* plane = planes + (long long)img * plane_stride;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned g = lane >> 2, t = lane & 3;
    const unsigned wy = warp >> 2, wx = warp & 3;            // warp tile: rows 32 wy .., columns 32 wx ..
    const unsigned tsw = ((t & 1u) ? 2u : 0u) | ((t & 2u) ? 16u : 0u);   // column swizzle of T row k, k & 3 == t
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const unsigned r = min(y0 + 32 * wy + 8 * i + g, h - 1u), x = x0 + 32 * wx + 4 * t;
        if (x < w) lr_prefetch(src + 3 * ((size_t)img * w * h + (size_t)r * w + x));
    }
    for (unsigned u0 = 0; u0 < kr; u0 += kLrRows) {
        const unsigned rows = min((unsigned)kLrRows, kr - u0), rows8 = (rows + 7u) & ~7u;
        float4 tv[8];
        float cv[16];
        lr_fetch_tile(plane, cy_tab, w, h, x0, y0, u0, rows, tv, cv);
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const unsigned i = threadIdx.x + 256u * it, r = i >> 5, q = i & 31u;
            float4 v = tv[it];
            if (r & 1u) v = make_float4(v.z, v.w, v.x, v.y);                      // x ^ 2 inside the aligned group of 4
            *(float4*)&ts[r][(4 * q) ^ ((r & 2u) ? 16u : 0u)] = v;                // x ^ 16 moves the whole group
        }
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            const unsigned i = threadIdx.x + 256u * it, r = i >> 6, yy = i & 63u;
            cy[r][yy ^ (8u * (r & 3u))] = (u0 + kLrRows > kLrTab) ? lr_cy_slow(it, u0, y0, rows, h, cv[it]) : cv[it];
        }
        __syncthreads();
        for (unsigned ks = 0; ks < rows8; ks += 8) {
            // A fragments (row-major 16 x 8): a0 (g, t), a1 (g+8, t), a2 (g, t+4), a3 (g+8, t+4); A[m][k] = CY[k][m]
            unsigned ah[2][4], al[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const unsigned m = 32 * wy + 16 * i + g;
                const float a[4] = {cy[ks + t][m ^ (8u * t)], cy[ks + t][(m + 8) ^ (8u * t)], cy[ks + t + 4][m ^ (8u * t)], cy[ks + t + 4][(m + 8) ^ (8u * t)]};
#pragma unroll
                for (int e = 0; e < 4; ++e) { ah[i][e] = tf32_hi(a[e]); al[i][e] = __float_as_uint(a[e] - __uint_as_float(ah[i][e])); }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // B fragments (8 x 8, column n = g): b0 (k = t), b1 (k = t + 4); B[k][n] = T[k][pixel(n)]
                const unsigned x = 32 * wx + 16 * (j >> 1) + 4 * (g >> 1) + 2 * (j & 1) + (g & 1);
                const float b[2] = {ts[ks + t][x ^ tsw], ts[ks + t + 4][x ^ tsw]};
                unsigned bh[2], bl[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) { bh[e] = tf32_hi(b[e]); bl[e] = __float_as_uint(b[e] - __uint_as_float(bh[e])); }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    mma_tf32(acc[i][j], al[i], bh);
                    mma_tf32(acc[i][j], ah[i], bl);
                    mma_tf32(acc[i][j], ah[i], bh);
                }
            }
        }
    }
    const unsigned char* s = src + 3 * (size_t)img * w * h;
    unsigned char* d = dst + 3 * (size_t)img * w * h;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const unsigned ra = y0 + 32 * wy + 16 * i + g, rb = ra + 8;
        if (ra >= h) continue;
        const bool hb = rb < h;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const unsigned x = x0 + 32 * wx + 16 * q + 4 * t;
            if (x >= w) continue;
            const size_t pa = (size_t)ra * w + x, pb = (size_t)(hb ? rb : ra) * w + x;
            const unsigned* qa = (const unsigned*)(s + 3 * pa);
            const unsigned* qb = (const unsigned*)(s + 3 * pb);
            const unsigned wa[3] = {__ldg(qa), __ldg(qa + 1), __ldg(qa + 2)}, wb[3] = {__ldg(qb), __ldg(qb + 1), __ldg(qb + 2)};
            cplx y2[4];
            fast::luma4x2_words(wa, wb, nz, y2);
            // accumulator layout: c0 (g, 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1); n-tiles 2q, 2q+1 = pixels +0,+1 / +2,+3
            y2[0] = __fadd2_rn(y2[0], make_float2(acc[i][2 * q][0], acc[i][2 * q][2]));
            y2[1] = __fadd2_rn(y2[1], make_float2(acc[i][2 * q][1], acc[i][2 * q][3]));
            y2[2] = __fadd2_rn(y2[2], make_float2(acc[i][2 * q + 1][0], acc[i][2 * q + 1][2]));
            y2[3] = __fadd2_rn(y2[3], make_float2(acc[i][2 * q + 1][1], acc[i][2 * q + 1][3]));
            unsigned oa[3], ob[3];
            fast::rgb8_out4x2_words(wa, wb, nz, y2, oa, ob);
            unsigned* da = (unsigned*)(d + 3 * pa);
            unsigned* db = (unsigned*)(d + 3 * pb);
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                da[e] = oa[e];
                if (hb) db[e] = ob[e];
            }
        }
    }
#endif
}

}  // namespace ssw
