// Embedding scatter, extraction gather, similarity, N(0,1) marks, planar YIQ and synthetic frames.
#pragma once
#include "pdl.cuh"
#include <cstdint>
#include <cuda_runtime.h>

#include "color.cuh"

namespace ssw {

// ------------------------------------------------------------------------------------------------
// Writer::embed_watermark -- /root/reference/src/algorithm.rs:382-410, insert functions :414-432.
// One thread per rank i: coefficient idx[i] is modulated by mark value i.  No FMA contraction so the
// result is bit-identical to the reference given the same coefficient (Option 3 up to expf rounding).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float insert_fn(int method, float alpha, float orig, float w) {
    if (method == 1) return __fadd_rn(orig, __fmul_rn(alpha, w));
    if (method == 2) return __fmul_rn(orig, __fadd_rn(1.0f, __fmul_rn(alpha, w)));
    return __fmul_rn(orig, expf(__fmul_rn(alpha, w)));
}

// extract functions, /root/reference/src/algorithm.rs:566-593 (separately rounded operations, as the reference)
__device__ __forceinline__ float extract_fn(int method, float alpha, float b, float d) {
    if (method == 1) return __fdiv_rn(__fsub_rn(d, b), alpha);
    if (method == 2) return __fdiv_rn(__fsub_rn(d, b), __fmul_rn(b, alpha));
    return __fdiv_rn(logf(__fdiv_rn(d, b)), alpha);
}

// marks: [batch][n_marks][mark_stride] f32, lens: [n_marks] (NULL = all k)
__global__ void embed_scatter_kernel(float* __restrict__ planes, long long plane_stride,
                                     const unsigned* __restrict__ idx, long long idx_stride, unsigned k,
                                     const float* __restrict__ marks, long long mark_stride, int n_marks,
                                     const unsigned* __restrict__ lens, int method, float alpha) {
    pdl_enter();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned img = blockIdx.y;
    if (i >= k) return;
    float* plane = planes + (long long)img * plane_stride;
    const unsigned p = idx[(long long)img * idx_stride + i];
    if (p == 0xFFFFFFFFu) return;   // the ordering of this frame failed (candidate overflow, reported): leave it unmarked
    const float* mk0 = marks + (long long)img * n_marks * mark_stride;
    const float orig = plane[p];
    if (n_marks == 1) {
        if (!lens || i < lens[0]) plane[p] = insert_fn(method, alpha, orig, mk0[i]);  // :394-398
        return;
    }
    float c = orig;  // :399-408: deltas against the original coefficient, summed in mark order
    for (int m = 0; m < n_marks; ++m) {
        if (lens && i >= lens[m]) continue;
        const float updated = insert_fn(method, alpha, orig, mk0[(long long)m * mark_stride + i]);
        c = __fadd_rn(c, __fsub_rn(updated, orig));
    }
    plane[p] = c;
}

// ------------------------------------------------------------------------------------------------
// Reader::extract_watermark -- src/algorithm.rs:543-562, extract functions :566-593
// ------------------------------------------------------------------------------------------------
__global__ void extract_gather_kernel(const float* __restrict__ base, const float* __restrict__ derived,
                                      long long plane_stride, const unsigned* __restrict__ idx,
                                      long long idx_stride, unsigned n, int method, float alpha,
                                      float* __restrict__ out, long long out_stride) {
    pdl_enter();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned img = blockIdx.y;
    if (i >= n) return;
    const unsigned p = idx[(long long)img * idx_stride + i];
    if (p == 0xFFFFFFFFu) { out[(long long)img * out_stride + i] = 0.f; return; }   // ordering failed (reported): defined output
    const float b = base[(long long)img * plane_stride + p];
    const float d = derived[(long long)img * plane_stride + p];
    out[(long long)img * out_stride + i] = extract_fn(method, alpha, b, d);
}

// ------------------------------------------------------------------------------------------------
// Sharded frames: each rank keeps the coefficients of its columns [col0, col0+ncols)
// transposed, local plane [ncols][height].  The ordered index list holds the reference's flat indices
// p = r*width + c; only the owner of column c touches coefficient p.
// ------------------------------------------------------------------------------------------------
struct ShardLayout {
    unsigned width, height;   // whole frame
    unsigned col0, ncols;     // columns owned by this rank
};

__device__ __forceinline__ bool shard_local(const ShardLayout& L, unsigned p, unsigned long long* q) {
    const unsigned r = p / L.width, c = p - r * L.width;
    if (c < L.col0 || c >= L.col0 + L.ncols) return false;
    *q = (unsigned long long)(c - L.col0) * L.height + r;
    return true;
}

__global__ void embed_scatter_shard_kernel(float* __restrict__ plane, ShardLayout L, const unsigned* __restrict__ idx,
                                           unsigned k, const float* __restrict__ marks, long long mark_stride,
                                           int n_marks, const unsigned* __restrict__ lens, int method, float alpha) {
    pdl_enter();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    unsigned long long q;
    if (!shard_local(L, idx[i], &q)) return;
    const float orig = plane[q];
    if (n_marks == 1) {
        if (!lens || i < lens[0]) plane[q] = insert_fn(method, alpha, orig, marks[i]);
        return;
    }
    float c = orig;
    for (int m = 0; m < n_marks; ++m) {
        if (lens && i >= lens[m]) continue;
        const float updated = insert_fn(method, alpha, orig, marks[(long long)m * mark_stride + i]);
        c = __fadd_rn(c, __fsub_rn(updated, orig));
    }
    plane[q] = c;
}

// out[i] = extracted value if this rank owns coefficient idx[i], else 0 (the ranks' vectors are summed)
__global__ void extract_gather_shard_kernel(const float* __restrict__ base, const float* __restrict__ derived,
                                            ShardLayout L, const unsigned* __restrict__ idx, unsigned n, int method,
                                            float alpha, float* __restrict__ out) {
    pdl_enter();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long q;
    float r = 0.f;
    if (shard_local(L, idx[i], &q)) {
        r = extract_fn(method, alpha, base[q], derived[q]);
    }
    out[i] = r;
}

// batched 2-D transpose of f32 tiles through shared memory (both sides coalesced):
// dst[b][c][r] = src[b][r][c], r < rows, c < cols, leading dimensions src_ld / dst_ld
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ src, unsigned rows, unsigned cols, long long src_ld, long long src_bstride,
                 float* __restrict__ dst, long long dst_ld, long long dst_bstride) {
    pdl_enter();
    __shared__ float tile[32][33];
    const float* s = src + (long long)blockIdx.z * src_bstride;
    float* d = dst + (long long)blockIdx.z * dst_bstride;
    const unsigned c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const unsigned r = r0 + ty + i, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + i][tx] = s[(long long)r * src_ld + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const unsigned c = c0 + ty + i, r = r0 + tx;
        if (r < rows && c < cols) d[(long long)c * dst_ld + r] = tile[tx][ty + i];
    }
}

// ------------------------------------------------------------------------------------------------
// Tester::similarity -- src/algorithm.rs:696-714, against a bank of marks [n_marks][n].
// One thread per stored mark walks the vectors in the reference's sequential order with separate
// f32 multiply and add, so every score is bit-identical to the reference loop.  The bank tile is
// staged through shared memory so global reads stay coalesced (each mark row is contiguous).
// The mark-independent denominator is computed once per extracted vector (similarity_den_kernel).
// ------------------------------------------------------------------------------------------------
constexpr int kSimMarks = 128;  // marks (threads) per CTA
constexpr int kSimChunk = 32;   // elements staged per step

// sum of squares of each extracted vector in the reference's sequential f32 order (the denominator of
// every score of that vector; src/algorithm.rs:709 accumulates it inside the same loop)
// Sequential f32 sum  acc = (((0 + v_0) + v_1) + ...)  of per-element products, by one warp: the products
// (round-to-nearest multiplies, order independent) are formed lane-parallel -- all global loads of a 2048-
// element block in flight at once -- and parked in shared memory; only the adds run in the reference's order,
// one broadcast LDS (off the dependent path) + one FADD per element.  kind 0: x*m, kind 1: x*x.
constexpr int kSeqBlock = 2048;

__device__ __forceinline__ float seq_sum_products(const float* __restrict__ x, const float* __restrict__ m, unsigned n,
                                                  int lane, int kind, float* __restrict__ prod /* [kSeqBlock] of this warp */) {
    float acc = 0.f;
    for (unsigned b0 = 0; b0 < n; b0 += kSeqBlock) {
        const unsigned len = min((unsigned)kSeqBlock, n - b0);
#pragma unroll 8
        for (unsigned j = lane; j < len; j += 32) {
            const float xv = __ldg(x + b0 + j);
            prod[j] = kind == 0 ? __fmul_rn(xv, __ldg(m + b0 + j)) : __fmul_rn(xv, xv);
        }
        __syncwarp();
        unsigned j = 0;
        for (; j + 16 <= len; j += 16) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = prod[j + i];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc = __fadd_rn(acc, v[i]);
        }
        for (; j < len; ++j) acc = __fadd_rn(acc, prod[j]);
        __syncwarp();
    }
    return acc;
}

__global__ void __launch_bounds__(128)
similarity_den_kernel(const float* __restrict__ extracted, unsigned n, long long ext_stride,
                      unsigned n_ext, float* __restrict__ den) {
    pdl_enter();
    __shared__ float prod[4][kSeqBlock];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned e = blockIdx.x * 4 + warp;
    if (e >= n_ext) return;
    const float d = seq_sum_products(extracted + (long long)e * ext_stride, nullptr, n, lane, 1, prod[warp]);
    if (lane == 0) den[e] = d;
}

__global__ void __launch_bounds__(kSimMarks)
similarity_bank_kernel(const float* __restrict__ bank, size_t n_marks, unsigned n,
                       const float* __restrict__ extracted, long long ext_stride, const float* __restrict__ den,
                       float* __restrict__ out, long long out_stride) {
    pdl_enter();
    __shared__ float tile[kSimMarks][kSimChunk + 1];
    __shared__ float ex[kSimChunk];
    const size_t m0 = (size_t)blockIdx.x * kSimMarks;
    const unsigned e = blockIdx.y;  // extracted vector
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    constexpr int kRowsPerWarp = kSimMarks / (kSimMarks / 32);  // rows staged by each warp per chunk (32)
    const float* ext = extracted + (long long)e * ext_stride;
    const int rows = (int)min((size_t)kSimMarks, n_marks - m0);
    float nom = 0.f;
    float pre[kRowsPerWarp];  // next chunk, prefetched while the current one is consumed
    float pre_ex = 0.f;
    auto fetch = [&](unsigned j0) {
        const unsigned len = min((unsigned)kSimChunk, n - j0);
#pragma unroll
        for (int i = 0; i < kRowsPerWarp; ++i) {
            const int r = warp + i * (kSimMarks / 32);
            pre[i] = (r < rows && (unsigned)lane < len) ? __ldg(bank + (m0 + r) * n + j0 + lane) : 0.f;
        }
        pre_ex = (t < (int)len) ? __ldg(ext + j0 + t) : 0.f;
    };
    fetch(0);
    for (unsigned j0 = 0; j0 < n; j0 += kSimChunk) {
        const unsigned len = min((unsigned)kSimChunk, n - j0);
#pragma unroll
        for (int i = 0; i < kRowsPerWarp; ++i) tile[warp + i * (kSimMarks / 32)][lane] = pre[i];
        if (t < kSimChunk) ex[t] = pre_ex;
        __syncthreads();
        if (j0 + kSimChunk < n) fetch(j0 + kSimChunk);
        if (len == (unsigned)kSimChunk) {
#pragma unroll
            for (int j = 0; j < kSimChunk; ++j) nom = __fadd_rn(nom, __fmul_rn(ex[j], tile[t][j]));
        } else {
            for (unsigned j = 0; j < len; ++j) nom = __fadd_rn(nom, __fmul_rn(ex[j], tile[t][j]));
        }
        __syncthreads();
    }
    const size_t m = m0 + t;
    if (m < n_marks) out[(long long)e * out_stride + m] = __fdiv_rn(nom, __fsqrt_rn(__ldg(den + e)));
}

// ------------------------------------------------------------------------------------------------
// Bank search at HBM speed (default): ONE WARP PER STORED MARK.  The extracted vector sits in shared memory, every
// lane streams float4s of the mark row (fully coalesced 512-byte warp reads, all loads of a row in flight at once),
// four partial sums per lane, then a fixed-shape reduction ((a0+a1)+(a2+a3), xor-shuffle tree) -- deterministic, and
// within a few ulp of the sequential loop of src/algorithm.rs:696-714 (the contract asks for 1e-3 relative; the
// bit-identical one-thread-per-mark kernel above stays available: SSW_SIM_EXACT=1).
// The denominator is the sequentially summed one of similarity_den_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kWarpSimThreads = 256;       // 8 warps per CTA
constexpr int kWarpSimMaxN = 8192;         // extracted vector staged in shared memory (32 KB)

__device__ __forceinline__ float4 ld_stream4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(kWarpSimThreads)
similarity_bank_warp_kernel(const float* __restrict__ bank, size_t n_marks, unsigned n,
                            const float* __restrict__ extracted, long long ext_stride, const float* __restrict__ den,
                            float* __restrict__ out, long long out_stride) {
    pdl_enter();
    __shared__ __align__(16) float ex[kWarpSimMaxN];
    const unsigned e = blockIdx.y;
    const float* ext = extracted + (long long)e * ext_stride;
    for (unsigned j = threadIdx.x; j < n; j += kWarpSimThreads) ex[j] = __ldg(ext + j);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const size_t warp0 = (size_t)blockIdx.x * (kWarpSimThreads / 32) + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * (kWarpSimThreads / 32);
    const float rden = __fsqrt_rn(__ldg(den + e));
    const unsigned n4 = n >> 2;
    const float4* ex4 = (const float4*)ex;
    if (n4 <= 8u * 32u && (n & 3u) == 0u && ((((size_t)bank) & 15) == 0)) {
        // rows of up to 1024 values (every row 16-byte aligned): the loads of the warp's NEXT row are issued before the
        // current row is reduced, so the memory pipe never drains between rows (two register sets, loop unrolled by two)
        auto load_row = [&](size_t m, float4* v) {
            const float4* r4 = (const float4*)(bank + m * n);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned j = 32 * u + lane;
                v[u] = j < n4 ? ld_stream4(r4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto reduce_row = [&](size_t m, const float4* v) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned j = 32 * u + lane;
                const float4 x = j < n4 ? ex4[j] : make_float4(0.f, 0.f, 0.f, 0.f);
                a0 = fmaf(x.x, v[u].x, a0); a1 = fmaf(x.y, v[u].y, a1); a2 = fmaf(x.z, v[u].z, a2); a3 = fmaf(x.w, v[u].w, a3);
            }
            float sum = (a0 + a1) + (a2 + a3);
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
            if (lane == 0) out[(long long)e * out_stride + m] = __fdiv_rn(sum, rden);
        };
        float4 va[8], vb[8];
        size_t m = warp0;
        if (m < n_marks) load_row(m, va);
        while (m < n_marks) {
            size_t mn = m + nwarps;
            if (mn < n_marks) load_row(mn, vb);
            reduce_row(m, va);
            m = mn;
            if (m >= n_marks) break;
            mn = m + nwarps;
            if (mn < n_marks) load_row(mn, va);
            reduce_row(m, vb);
            m = mn;
        }
        return;
    }
    for (size_t m = warp0; m < n_marks; m += nwarps) {
        const float* row = bank + m * n;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if ((((size_t)row) & 15) == 0) {
            const float4* r4 = (const float4*)row;
            // blocks of 8 x 32 float4: every lane issues its 8 (predicated) 16-byte loads before the first use, so a
            // row costs the warp ONE memory latency whatever n is (n = 1000: 250 float4, lanes 26..31 idle in the last slot)
            for (unsigned base = 0; base < n4; base += 8 * 32) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const unsigned j = base + 32 * u + lane;
                    v[u] = j < n4 ? ld_stream4(r4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const unsigned j = base + 32 * u + lane;
                    const float4 x = j < n4 ? ex4[j] : make_float4(0.f, 0.f, 0.f, 0.f);
                    a0 = fmaf(x.x, v[u].x, a0); a1 = fmaf(x.y, v[u].y, a1); a2 = fmaf(x.z, v[u].z, a2); a3 = fmaf(x.w, v[u].w, a3);
                }
            }
            for (unsigned t = (n4 << 2) + lane; t < n; t += 32) a0 = fmaf(ex[t], __ldg(row + t), a0);
        } else {
            for (unsigned t = lane; t < n; t += 32) a0 = fmaf(ex[t], __ldg(row + t), a0);
        }
        float sum = (a0 + a1) + (a2 + a3);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
        if (lane == 0) out[(long long)e * out_stride + m] = __fdiv_rn(sum, rden);
    }
}

constexpr unsigned kBadIndex = 0xFFFFFFFFu;   // index list entry of a frame whose ordering failed (candidate overflow)

// 1:1 form: extracted vector i against mark i, sequential order (the fused extract pipeline under SSW_SIM_EXACT=1;
// by default the score is reduced by the last CTA of topk_rank, select_kernels.cuh).
constexpr int kPairsPerCta = 2;   // two warps per pair: one walks the numerator chain, one the denominator chain

__global__ void __launch_bounds__(kPairsPerCta * 64)
similarity_pairs_kernel(const float* __restrict__ marks, const float* __restrict__ extracted, unsigned n,
                        long long stride, unsigned n_pairs, float* __restrict__ out) {
    pdl_enter();
    __shared__ float part[kPairsPerCta][2];
    __shared__ float prod[kPairsPerCta * 2][kSeqBlock];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp >> 1, kind = warp & 1;
    const unsigned pair = blockIdx.x * kPairsPerCta + slot;
    if (pair < n_pairs) {
        const float v = seq_sum_products(extracted + (long long)pair * stride, marks + (long long)pair * stride, n, lane, kind,
                                         prod[warp]);
        if (lane == 0) part[slot][kind] = v;
    }
    __syncthreads();
    if (pair < n_pairs && kind == 0 && lane == 0) out[pair] = __fdiv_rn(part[slot][0], __fsqrt_rn(part[slot][1]));
}

// ------------------------------------------------------------------------------------------------
// MarkBuf::generate_normal -- src/algorithm.rs:619-626 (distributional parity only: the reference
// draws from the OS-seeded thread_rng).  Philox4x32-10 counter RNG + Box-Muller.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint4& ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
}

__global__ void normal_fill_kernel(float* __restrict__ out, size_t n, unsigned long long seed,
                                   unsigned long long stream_offset) {
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // quad index
    if (q * 4 >= n) return;
    const unsigned long long c = q + stream_offset;
    uint4 ctr = make_uint4((unsigned)c, (unsigned)(c >> 32), 0u, 0u);
    philox4x32_10(ctr, make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const float k24 = 5.9604644775390625e-8f;  // 2^-24
    const float u0 = ((float)(ctr.x >> 8) + 0.5f) * k24, u1 = ((float)(ctr.y >> 8) + 0.5f) * k24;
    const float u2 = ((float)(ctr.z >> 8) + 0.5f) * k24, u3 = ((float)(ctr.w >> 8) + 0.5f) * k24;
    const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
    float s0, c0, s1, c1;
    sincospif(2.0f * u1, &s0, &c0);
    sincospif(2.0f * u3, &s1, &c1);
    const float v[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
    for (int j = 0; j < 4; ++j)
        if (q * 4 + j < n) out[q * 4 + j] = v[j];
}

// ------------------------------------------------------------------------------------------------
// planar YIQ <-> interleaved RGB32F -- src/yiq.rs:177-197 (stand-alone form of the fused passes)
// ------------------------------------------------------------------------------------------------
__global__ void rgb32f_to_yiq_kernel(const float* __restrict__ rgb, size_t npix, float* __restrict__ y,
                                     float* __restrict__ i, float* __restrict__ q) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const float r = rgb[3 * p], g = rgb[3 * p + 1], b = rgb[3 * p + 2];
    y[p] = rgb_to_y(r, g, b);
    i[p] = rgb_to_i(r, g, b);
    q[p] = rgb_to_q(r, g, b);
}

__global__ void yiq_to_rgb32f_kernel(const float* __restrict__ y, const float* __restrict__ i,
                                     const float* __restrict__ q, size_t npix, float* __restrict__ rgb) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    float r, g, b;
    yiq_to_rgb(y[p], i[p], q[p], r, g, b);
    rgb[3 * p] = r; rgb[3 * p + 1] = g; rgb[3 * p + 2] = b;
}

// ------------------------------------------------------------------------------------------------
// synthetic natural-image-like frames; bit-identical to oracle synth_frame()
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long sm64(unsigned long long x) {
    unsigned long long z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void synth_frame_kernel(unsigned char* __restrict__ out, unsigned w, unsigned h,
                                   unsigned long long seed, unsigned first_image, unsigned row0) {
    const unsigned x = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned yl = blockIdx.y;          // row inside the generated block of h rows
    const unsigned y = row0 + yl;            // row of the frame
    const unsigned img = blockIdx.z;
    if (x >= w) return;
    const unsigned long long base = seed ^ ((unsigned long long)(first_image + img) * 0x9E3779B97F4A7C15ull);
    unsigned char* o = out + ((size_t)img * h * w + (size_t)yl * w + x) * 3;
    for (unsigned c = 0; c < 3; ++c) {
        unsigned long long acc = 0;
        for (unsigned oct = 2; oct <= 8; ++oct) {
            const unsigned long long s = 1ull << oct;
            const unsigned long long X = x >> oct, Y = y >> oct, fx = x & (s - 1), fy = y & (s - 1);
            const unsigned long long tag = base ^ ((unsigned long long)oct << 58) ^ ((unsigned long long)c << 56);
            const unsigned long long l00 = sm64(tag ^ (Y << 28) ^ X) >> 56;
            const unsigned long long l10 = sm64(tag ^ (Y << 28) ^ (X + 1)) >> 56;
            const unsigned long long l01 = sm64(tag ^ ((Y + 1) << 28) ^ X) >> 56;
            const unsigned long long l11 = sm64(tag ^ ((Y + 1) << 28) ^ (X + 1)) >> 56;
            const unsigned long long top = l00 * (s - fx) + l10 * fx, bot = l01 * (s - fx) + l11 * fx;
            acc += ((top * (s - fy) + bot * fy) >> (2 * oct)) << oct;
        }
        const long long noise = (long long)(sm64(base ^ 0xABCDEFull ^ ((unsigned long long)c << 56) ^ ((unsigned long long)y << 28) ^ x) >> 61);
        long long v = (long long)(acc / 508ull) + noise - 4;
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        o[c] = (unsigned char)v;
    }
}

}  // namespace ssw
