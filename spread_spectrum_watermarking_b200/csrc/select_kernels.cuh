// Exact ordered top-k of the coefficient plane by the reference's comparator.
//
// Replaces `obtain_indices_by_function` (/root/reference/src/algorithm.rs:200-210): a *stable*
// descending sort of indices 1..w*h-1 (flat index 0, the DC term, is skipped) by
//   Energy           fl32(c*c)                         (:214-221)
//   EnergyOrthogonal fl32(v*v), v = ortho_scaling(c)   (:235-280)
//   Legacy           v                                 (:225-232 through :235-280)
// compared with f32::total_cmp.  Stability means ties keep ascending index order, so the order is
// the descending order of the 64-bit composite  (total_cmp_key(value) << 32) | (0xFFFFFFFF - index),
// which has no ties at all.  Only the first k entries are ever consumed (mark length), so instead of
// sorting w*h-1 elements:
//   1. topk_block_bin: 4096-bin histogram (top 12 key bits) of the low-frequency block only; the bin
//                     of its k-th largest key bounds the plane's from below (see the kernel);
//      topk_hist    : (repair path) the same histogram over the whole plane, one read;
//   2. topk_collect : second read, every element whose bin >= that bin is appended to a small
//                     candidate list (k + one bin's worth of elements);
//   3. topk_sort    : one CTA bitonic-sorts the candidates by the composite key and emits the first
//                     k indices.
// Candidate overflow (pathological spectra with > kTopkCap near-equal keys) is flagged and the host
// re-runs the exact general path (radix sort of all candidates, select_general.cuh).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "pdl.cuh"

namespace ssw {

constexpr int kHistBits = 12;
constexpr int kHistBins = 1 << kHistBits;
constexpr int kTopkCap = 8192;  // candidates per image held by the single-CTA sort (64 KB of smem)

struct OrderConsts {
    int mode;  // SSW_ORDER_*
    int w;     // width of the whole frame
    float s_k0_w, s_k0_h, s_w, s_h;  // src/algorithm.rs:245-250
    // sharded frames keep their coefficients transposed: local plane [columns col0..][t_ld = frame height];
    // t_ld == 0: ordinary row-major plane, local position == the reference's flat index r*w + c
    unsigned t_ld, t_col0;
};

// local linear position -> the reference's flat index r*w + c (src/algorithm.rs:204)
__device__ __forceinline__ unsigned flat_index(unsigned q, const OrderConsts& oc) {
    if (oc.t_ld == 0u) return q;
    const unsigned c_local = q / oc.t_ld, r = q - c_local * oc.t_ld;
    return r * (unsigned)oc.w + oc.t_col0 + c_local;
}

struct TopkScratch {
    unsigned* hist;        // [batch][kHistBins], zero between calls
    unsigned* ticket;      // [batch], zero between calls
    unsigned* sel_bin;     // [batch]
    unsigned* cand_count;  // [batch], zero between calls
    unsigned long long* cand;  // [batch][kTopkCap]
    unsigned* overflow;    // [1] sticky counter of images whose candidate list overflowed
};

__device__ __forceinline__ unsigned total_cmp_key(float v) {
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ unsigned order_key(float c, unsigned p, const OrderConsts& oc) {
    if (oc.mode == 0) return total_cmp_key(__fmul_rn(c, c));
    // ortho_scaling (:240-266): scaling = 1.0 * (row) * (col); scaling * value
    float sc = __fmul_rn(1.0f, (p < (unsigned)oc.w) ? oc.s_k0_w : oc.s_w);
    sc = __fmul_rn(sc, (p % (unsigned)oc.w == 0u) ? oc.s_k0_h : oc.s_h);
    const float v = __fmul_rn(sc, c);
    return total_cmp_key(oc.mode == 1 ? __fmul_rn(v, v) : v);
}

// bin b with  #(bin > b) < k <= #(bin >= b)  of a 4096-bin shared-memory histogram (0 if fewer than k
// elements were counted); all threads of the CTA must call, blockDim.x must divide 4096 and be a
// multiple of 32 (<= 1024).  Suffix sums by warp shuffles; the one thread whose run of bins crosses
// the k-th element resolves the bin inside its run.
__device__ __forceinline__ unsigned find_kth_bin(const unsigned* sh, unsigned k) {
    __shared__ unsigned wsum[32];
    __shared__ unsigned s_bin;
    const int per = kHistBins / blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    unsigned mine = 0;
    for (int j = 0; j < per; ++j) mine += sh[threadIdx.x * per + j];
    // inclusive suffix sum inside the warp: incl = sum over lanes >= lane
    unsigned incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_down_sync(0xFFFFFFFFu, incl, d);
        if (lane + d < 32) incl += o;
    }
    if (lane == 0) wsum[warp] = incl;
    if (threadIdx.x == 0) s_bin = 0u;
    __syncthreads();
    unsigned above_warps = 0;  // elements in the bins of all higher warps
    for (int u = warp + 1; u < nwarps; ++u) above_warps += wsum[u];
    const unsigned suffix = above_warps + incl;  // elements in bins >= first bin of this thread
    const unsigned above = suffix - mine;        // elements in bins above this thread's run
    if (above < k && suffix >= k) {
        unsigned a = above;
        int b = threadIdx.x * per + per - 1;
        for (; b > (int)threadIdx.x * per; --b) {
            if (a + sh[b] >= k) break;
            a += sh[b];
        }
        s_bin = (unsigned)b;
    }
    __syncthreads();
    return s_bin;
}

// ---- 1. histogram + threshold bin ---------------------------------------------------------------
__global__ void __launch_bounds__(512)
topk_hist_kernel(const float* __restrict__ planes, long long plane_stride, unsigned n, unsigned k,
                 OrderConsts oc, TopkScratch ts) {
    pdl_enter();
    __shared__ unsigned sh[kHistBins];
    __shared__ unsigned s_last;
    const unsigned img = blockIdx.y;
    const float* plane = planes + (long long)img * plane_stride;
    for (int i = threadIdx.x; i < kHistBins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const unsigned n4 = n >> 2;
    const bool vec = ((((size_t)plane) & 15) == 0);
    if (vec) {
        const float4* p4 = (const float4*)plane;
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
            const float4 v = __ldg(p4 + i);
            const unsigned p = i << 2;
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned g = flat_index(p + u, oc);
                if (g) atomicAdd(&sh[order_key(e[u], g, oc) >> (32 - kHistBits)], 1u);
            }
        }
        for (unsigned p = (n4 << 2) + blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
            const unsigned g = flat_index(p, oc);
            if (g) atomicAdd(&sh[order_key(plane[p], g, oc) >> (32 - kHistBits)], 1u);
        }
    } else {
        for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
            const unsigned g = flat_index(p, oc);
            if (g) atomicAdd(&sh[order_key(plane[p], g, oc) >> (32 - kHistBits)], 1u);
        }
    }
    __syncthreads();
    unsigned* gh = ts.hist + (size_t)img * kHistBins;
    for (int i = threadIdx.x; i < kHistBins; i += blockDim.x)
        if (sh[i]) atomicAdd(&gh[i], sh[i]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&ts.ticket[img], 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    // last CTA of this image: find the bin b with  #(bin > b) < k <= #(bin >= b)
    __threadfence();
    for (int i = threadIdx.x; i < kHistBins; i += blockDim.x) {
        sh[i] = __ldcg(&gh[i]);
        gh[i] = 0;  // leave the scratch clean for the next call
    }
    __syncthreads();
    const unsigned b = find_kth_bin(sh, k);
    if (threadIdx.x == 0) {
        ts.sel_bin[img] = b;
        ts.ticket[img] = 0;
    }
}

// ---- 1'. threshold bin from a low-frequency block only ---------------------------------------------
// The k-th largest key of ANY subset is a lower bound of the k-th largest key of the whole plane, so
// the bin holding the k-th largest key of the top-left block (where natural images keep their
// energy) is a valid -- and for natural images tight -- selection bin for topk_collect: every
// top-k element of the plane lies in a bin >= it.  One CTA per image reads <= 32k coefficients
// instead of one full pass over the plane.  A loose bound (noise-like spectra) only costs a
// candidate overflow, which is detected by topk_sort and repaired with the full histogram.
constexpr int kBlockRows = 128, kBlockCols = 256;

__global__ void __launch_bounds__(512)
topk_block_bin_kernel(const float* __restrict__ planes, long long plane_stride, unsigned w, unsigned h, unsigned k,
                      OrderConsts oc, TopkScratch ts) {
    pdl_enter();
    // grid (x = slices of the block, y = image): per-CTA shared histogram -> global histogram -> the last CTA of
    // the image finds the bin (same ticket scheme as topk_hist_kernel; scratch is left zeroed)
    __shared__ unsigned sh[kHistBins];
    __shared__ unsigned s_last;
    const unsigned img = blockIdx.y;
    const float* plane = planes + (long long)img * plane_stride;
    for (int i = threadIdx.x; i < kHistBins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const unsigned br = min(h, (unsigned)kBlockRows), bc = min(w, (unsigned)kBlockCols);
    const unsigned total = br * bc;
    for (unsigned e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < total; e0 += 4 * gridDim.x * blockDim.x) {
        float v[4];
        unsigned p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {   // independent loads in flight
            const unsigned e = e0 + u * gridDim.x * blockDim.x;
            const unsigned r = e / bc, c = e - r * bc;
            const unsigned q = r * w + c;            // local position (w = local line length)
            p[u] = e < total ? flat_index(q, oc) : 0u;
            v[u] = p[u] ? __ldg(plane + q) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (p[u]) atomicAdd(&sh[order_key(v[u], p[u], oc) >> (32 - kHistBits)], 1u);
    }
    __syncthreads();
    unsigned* gh = ts.hist + (size_t)img * kHistBins;
    for (int i = threadIdx.x; i < kHistBins; i += blockDim.x)
        if (sh[i]) atomicAdd(&gh[i], sh[i]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&ts.ticket[img], 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = threadIdx.x; i < kHistBins; i += blockDim.x) {
        sh[i] = __ldcg(&gh[i]);
        gh[i] = 0;
    }
    __syncthreads();
    const unsigned b = find_kth_bin(sh, k);
    if (threadIdx.x == 0) {
        ts.sel_bin[img] = b;
        ts.ticket[img] = 0;
    }
}

// ---- 2. collect candidates -----------------------------------------------------------------------
__device__ __forceinline__ void topk_push(unsigned key, unsigned p, unsigned bin_sel, unsigned* count,
                                          unsigned long long* cand) {
    if ((key >> (32 - kHistBits)) >= bin_sel) {
        const unsigned pos = atomicAdd(count, 1u);
        if (pos < (unsigned)kTopkCap) cand[pos] = ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - p);
    }
}

__global__ void __launch_bounds__(512)
topk_collect_kernel(const float* __restrict__ planes, long long plane_stride, unsigned n, OrderConsts oc,
                    TopkScratch ts) {
    pdl_enter();
    const unsigned img = blockIdx.y;
    const float* plane = planes + (long long)img * plane_stride;
    const unsigned bin_sel = ts.sel_bin[img];
    unsigned* count = ts.cand_count + img;
    unsigned long long* cand = ts.cand + (size_t)img * kTopkCap;
    const unsigned n4 = n >> 2;
    const bool vec = ((((size_t)plane) & 15) == 0);
    if (vec) {
        const float4* p4 = (const float4*)plane;
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
            const float4 v = __ldg(p4 + i);
            const unsigned p = i << 2;
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned g = flat_index(p + u, oc);
                if (g) topk_push(order_key(e[u], g, oc), g, bin_sel, count, cand);
            }
        }
        for (unsigned p = (n4 << 2) + blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
            const unsigned g = flat_index(p, oc);
            if (g) topk_push(order_key(plane[p], g, oc), g, bin_sel, count, cand);
        }
    } else {
        for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
            const unsigned g = flat_index(p, oc);
            if (g) topk_push(order_key(plane[p], g, oc), g, bin_sel, count, cand);
        }
    }
}

// ---- 1'+2 fused: every collect CTA bounds the k-th key itself -----------------------------------------
// For short marks (k <= kFusedMaxK) a block of <= 8192 low-frequency coefficients (64 rows x 128 columns:
// 8x the coefficients consumed) bounds the k-th key as tightly as the larger block of topk_block_bin (measured on
// the synthetic 4K / 1080p frames: identical candidate counts), and 8192 L2-resident values cost a collect CTA
// ~2 us -- less than the separate 16-CTA kernel with its global histogram, ticket and extra launch.
// Every CTA derives the same bin from the same data, so no inter-CTA communication is needed.
constexpr int kFusedRows = 64, kFusedCols = 128, kFusedMaxK = 1024;

__global__ void __launch_bounds__(512, 4)
topk_bin_collect_kernel(const float* __restrict__ planes, long long plane_stride, unsigned w, unsigned h, unsigned k,
                        OrderConsts oc, TopkScratch ts) {
    pdl_enter();
    __shared__ unsigned sh[kHistBins];
    const unsigned img = blockIdx.y;
    const float* plane = planes + (long long)img * plane_stride;
    const unsigned n = w * h;
    for (int i = threadIdx.x; i < kHistBins; i += blockDim.x) sh[i] = 0;
    const unsigned br = min(h, (unsigned)kFusedRows), bc = min(w, (unsigned)kFusedCols);
    const unsigned total = br * bc;
    __syncthreads();
    constexpr int PER = kFusedRows * kFusedCols / 512, BATCH = 8;   // 2 x 8 independent loads in flight per thread
#pragma unroll 1
    for (int u0 = 0; u0 < PER; u0 += BATCH) {
        float v[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            const unsigned e = threadIdx.x + (u0 + u) * 512u;
            const unsigned r = e / bc, c = e - r * bc;
            v[u] = (e < total && e) ? __ldg(plane + r * w + c) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            const unsigned e = threadIdx.x + (u0 + u) * 512u;
            const unsigned r = e / bc, c = e - r * bc;
            // row-major planes only (t_ld == 0): flat index == position; e == 0 is the DC term
            if (e < total && e) atomicAdd(&sh[order_key(v[u], r * w + c, oc) >> (32 - kHistBits)], 1u);
        }
    }
    __syncthreads();
    const unsigned bin_sel = find_kth_bin(sh, k);
    if (blockIdx.x == 0 && threadIdx.x == 0) ts.sel_bin[img] = bin_sel;
    // the collect pass proper (same as topk_collect_kernel)
    unsigned* count = ts.cand_count + img;
    unsigned long long* cand = ts.cand + (size_t)img * kTopkCap;
    const unsigned n4 = n >> 2;
    if ((((size_t)plane) & 15) == 0) {
        const float4* p4 = (const float4*)plane;
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
            const float4 q4 = __ldg(p4 + i);
            const unsigned q = i << 2;
            const float e[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (q + u) topk_push(order_key(e[u], q + u, oc), q + u, bin_sel, count, cand);
        }
        for (unsigned q = (n4 << 2) + blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x)
            if (q) topk_push(order_key(plane[q], q, oc), q, bin_sel, count, cand);
    } else {
        for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x)
            if (q) topk_push(order_key(plane[q], q, oc), q, bin_sel, count, cand);
    }
}

// ---- 2'. distributed merge: concatenate the candidate lists gathered from all ranks --------------
__global__ void __launch_bounds__(256)
topk_concat_kernel(const unsigned long long* __restrict__ lists, const unsigned* __restrict__ counts, unsigned n_lists,
                   unsigned list_cap, TopkScratch ts) {
    pdl_enter();
    __shared__ unsigned off[65];
    if (threadIdx.x == 0) {
        unsigned run = 0;
        for (unsigned l = 0; l < n_lists; ++l) { off[l] = run; run += min(counts[l], list_cap); }
        off[n_lists] = run;
        ts.cand_count[0] = run;  // > kTopkCap is reported as overflow by topk_sort
    }
    __syncthreads();
    for (unsigned l = 0; l < n_lists; ++l) {
        const unsigned c = min(counts[l], list_cap);
        for (unsigned i = threadIdx.x; i < c; i += blockDim.x)
            if (off[l] + i < (unsigned)kTopkCap) ts.cand[off[l] + i] = lists[(size_t)l * list_cap + i];
    }
}

// ---- 3. sort candidates, emit the first k indices ------------------------------------------------
// Bitonic network over 256*E keys, E consecutive keys per thread held in registers: strides < E are
// compare-exchanges inside a thread, strides < 32*E go through warp shuffles, and only the few
// strides >= 32*E cross warps through shared memory (6 of the 66 steps at 2048 keys).
constexpr int kSortThreads = 256;

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int mask) {
    const unsigned lo = __shfl_xor_sync(0xFFFFFFFFu, (unsigned)v, mask);
    const unsigned hi = __shfl_xor_sync(0xFFFFFFFFu, (unsigned)(v >> 32), mask);
    return ((unsigned long long)hi << 32) | lo;
}

template <int E>
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long (&v)[E], unsigned long long* sc) {
    // The size / stride loops stay ROLLED on purpose: one CTA runs this once, so a fully unrolled network
    // (hundreds of KB of straight-line code) would execute at instruction-fetch speed.
    const unsigned tid = threadIdx.x;
#pragma unroll 1
    for (unsigned size = 2; size <= (unsigned)(kSortThreads * E); size <<= 1) {
#pragma unroll 1
        for (unsigned stride = size >> 1; stride >= (unsigned)E; stride >>= 1) {
            const unsigned tmask = stride / E;  // partner thread = tid ^ tmask, same slot
            unsigned long long o[E];
            if (tmask < 32u) {
#pragma unroll
                for (int i = 0; i < E; ++i) o[i] = shfl_xor_u64(v[i], (int)tmask);
            } else {
#pragma unroll
                for (int i = 0; i < E; ++i) sc[i * kSortThreads + tid] = v[i];
                __syncthreads();
#pragma unroll
                for (int i = 0; i < E; ++i) o[i] = sc[i * kSortThreads + (tid ^ tmask)];
                __syncthreads();
            }
            // all E slots of a thread share the direction bits (stride, size >= E)
            const bool keep_max = (((tid * E) & stride) == 0) == (((tid * E) & size) == 0);
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const bool gt = v[i] > o[i];
                v[i] = (gt == keep_max) ? v[i] : o[i];
            }
        }
        // strides < E: compare-exchanges inside the thread (at most log2(E) steps, unrolled per stride)
#pragma unroll
        for (unsigned stride = E >> 1; stride > 0; stride >>= 1) {
            if (stride < size) {
#pragma unroll
                for (int i = 0; i < E; ++i) {
                    if ((i & stride) == 0) {
                        const bool desc = ((tid * E + i) & size) == 0;
                        const unsigned long long a = v[i], b = v[i + stride];
                        if ((a < b) == desc) { v[i] = b; v[i + stride] = a; }
                    }
                }
            }
        }
    }
}

template <int E>
__device__ __forceinline__ void topk_sort_body(const unsigned long long* __restrict__ cand, unsigned cnt, unsigned k,
                                               unsigned* __restrict__ out, unsigned long long* sc) {
    unsigned long long v[E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const unsigned idx = threadIdx.x * E + i;
        v[i] = idx < cnt ? __ldcg(cand + idx) : 0ull;
    }
    bitonic_sort_desc<E>(v, sc);
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const unsigned idx = threadIdx.x * E + i;
        if (idx < k) out[idx] = idx < cnt ? (0xFFFFFFFFu - (unsigned)(v[i] & 0xFFFFFFFFull)) : 0u;
    }
}

__global__ void __launch_bounds__(kSortThreads)
topk_sort_kernel(TopkScratch ts, unsigned k, unsigned* __restrict__ idx_out, long long idx_stride) {
    pdl_enter();
    extern __shared__ unsigned long long sc[];  // kTopkCap keys (exchange buffer of the cross-warp steps)
    const unsigned img = blockIdx.x;
    const unsigned total = ts.cand_count[img];
    const unsigned cnt = total < (unsigned)kTopkCap ? total : (unsigned)kTopkCap;
    const unsigned long long* cand = ts.cand + (size_t)img * kTopkCap;
    unsigned* out = idx_out + (long long)img * idx_stride;
    if (cnt <= 8u * kSortThreads) topk_sort_body<8>(cand, cnt, k, out, sc);
    else if (cnt <= 16u * kSortThreads) topk_sort_body<16>(cand, cnt, k, out, sc);
    else topk_sort_body<32>(cand, cnt, k, out, sc);
    if (threadIdx.x == 0) {
        if (total > (unsigned)kTopkCap || cnt < k) atomicAdd(ts.overflow, 1u);
        ts.cand_count[img] = 0;
    }
}

}  // namespace ssw
