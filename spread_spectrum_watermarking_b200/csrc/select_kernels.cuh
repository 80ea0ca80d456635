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
//   1. topk_hist    : one read of the plane, 4096-bin histogram of the key's top 12 bits; the last
//                     CTA to finish finds the bin holding the k-th largest key;
//   2. topk_collect : second read, every element whose bin >= that bin is appended to a small
//                     candidate list (k + one bin's worth of elements);
//   3. topk_sort    : one CTA bitonic-sorts the candidates by the composite key and emits the first
//                     k indices.
// Candidate overflow (pathological spectra with > kTopkCap near-equal keys) is flagged and the host
// re-runs the exact general path (radix sort of all candidates, select_general.cuh).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ssw {

constexpr int kHistBits = 12;
constexpr int kHistBins = 1 << kHistBits;
constexpr int kTopkCap = 8192;  // candidates per image held by the single-CTA sort (64 KB of smem)

struct OrderConsts {
    int mode;  // SSW_ORDER_*
    int w;
    float s_k0_w, s_k0_h, s_w, s_h;  // src/algorithm.rs:245-250
};

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

// ---- 1. histogram + threshold bin ---------------------------------------------------------------
__global__ void __launch_bounds__(512)
topk_hist_kernel(const float* __restrict__ planes, long long plane_stride, unsigned n, unsigned k,
                 OrderConsts oc, TopkScratch ts) {
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
            if (p) atomicAdd(&sh[order_key(v.x, p, oc) >> (32 - kHistBits)], 1u);
            atomicAdd(&sh[order_key(v.y, p + 1, oc) >> (32 - kHistBits)], 1u);
            atomicAdd(&sh[order_key(v.z, p + 2, oc) >> (32 - kHistBits)], 1u);
            atomicAdd(&sh[order_key(v.w, p + 3, oc) >> (32 - kHistBits)], 1u);
        }
        for (unsigned p = (n4 << 2) + blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x)
            if (p) atomicAdd(&sh[order_key(plane[p], p, oc) >> (32 - kHistBits)], 1u);
    } else {
        for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x)
            if (p) atomicAdd(&sh[order_key(plane[p], p, oc) >> (32 - kHistBits)], 1u);
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
    // suffix sums over 4096 bins: each thread owns a contiguous run, then a serial pass over partials
    __shared__ unsigned part[512];
    const int per = kHistBins / blockDim.x;  // blockDim.x divides 4096
    unsigned acc = 0;
    for (int j = 0; j < per; ++j) acc += sh[threadIdx.x * per + j];
    part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned above = 0;  // elements in bins above the current run
        int t = blockDim.x - 1;
        for (; t > 0; --t) {
            if (above + part[t] >= k) break;
            above += part[t];
        }
        int b = t * per + per - 1;
        for (; b > t * per; --b) {
            if (above + sh[b] >= k) break;
            above += sh[b];
        }
        ts.sel_bin[img] = (unsigned)b;
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
            if (p) topk_push(order_key(v.x, p, oc), p, bin_sel, count, cand);
            topk_push(order_key(v.y, p + 1, oc), p + 1, bin_sel, count, cand);
            topk_push(order_key(v.z, p + 2, oc), p + 2, bin_sel, count, cand);
            topk_push(order_key(v.w, p + 3, oc), p + 3, bin_sel, count, cand);
        }
        for (unsigned p = (n4 << 2) + blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x)
            if (p) topk_push(order_key(plane[p], p, oc), p, bin_sel, count, cand);
    } else {
        for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x)
            if (p) topk_push(order_key(plane[p], p, oc), p, bin_sel, count, cand);
    }
}

// ---- 3. sort candidates, emit the first k indices ------------------------------------------------
__global__ void __launch_bounds__(1024)
topk_sort_kernel(TopkScratch ts, unsigned k, unsigned* __restrict__ idx_out, long long idx_stride) {
    extern __shared__ unsigned long long sc[];
    const unsigned img = blockIdx.x;
    const unsigned total = ts.cand_count[img];
    const unsigned cnt = total < (unsigned)kTopkCap ? total : (unsigned)kTopkCap;
    unsigned m = 1;
    while (m < cnt) m <<= 1;
    const unsigned long long* cand = ts.cand + (size_t)img * kTopkCap;
    for (unsigned i = threadIdx.x; i < m; i += blockDim.x) sc[i] = i < cnt ? cand[i] : 0ull;
    __syncthreads();
    // bitonic sort, descending
    for (unsigned size = 2; size <= m; size <<= 1) {
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
                const unsigned lo = 2 * t - (t & (stride - 1));
                const unsigned hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = sc[lo], b = sc[hi];
                if ((a < b) == desc) { sc[lo] = b; sc[hi] = a; }
            }
            __syncthreads();
        }
    }
    unsigned* out = idx_out + (long long)img * idx_stride;
    for (unsigned i = threadIdx.x; i < k; i += blockDim.x)
        out[i] = i < cnt ? (0xFFFFFFFFu - (unsigned)(sc[i] & 0xFFFFFFFFull)) : 0u;
    if (threadIdx.x == 0) {
        if (total > (unsigned)kTopkCap || cnt < k) atomicAdd(ts.overflow, 1u);
        ts.cand_count[img] = 0;
    }
}

}  // namespace ssw
