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
//   3. topk_rank    : the rank of a candidate is the number of candidates with a larger composite key
//                     (unique keys); a few CTAs count, ranks < k are written to their place.
// Candidate overflow (pathological spectra with > kTopkCap near-equal keys) is flagged and the host
// re-runs the exact general path (radix sort of all candidates, select_general.cuh).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "pdl.cuh"
#include "mark_kernels.cuh"

namespace ssw {

constexpr int kHistBits = 12;
constexpr int kHistBins = 1 << kHistBits;
constexpr int kTopkCap = 8192;  // candidates per image (64 KB of shared memory in topk_rank)

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
    unsigned* maxcol;      // [1] largest column index of a coefficient modified by the embedding of this launch (TopkApply mode 1 with
                           //     width set): cleared by topk_collect, read by the partial inverse column pass (dct_pipe.cuh, PipeArgs::col_limit)
    unsigned* maxcol_img;  // [batch] the same per image: what an image gets back from the partial inverse depends on the image alone
    unsigned* tile_max;    // [batch][tile_count] per-tile coefficient maxima consumed by topk_collect_tiles; topk_rank clears them
    unsigned tile_count;
    unsigned* maxrow;      // [batch] largest coefficient row among the first k ordered indices (low-rank embed inverse);
                           // zeroed by topk_collect, 0xFFFFFFFF = the ordering of this frame failed
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

// The same for a subset of a CTA (threads 0..nthreads-1, whole warps, synchronising on named barrier `bar_id`) and
// any thread count: thread t owns the run of bins [t*per, t*per + per), per = ceil(4096 / nthreads).  `scratch`: 33 words
// of shared memory.  Used by the forward column pipeline, whose compute warps share the CTA with a producer warp.
__device__ __forceinline__ unsigned find_kth_bin_team(const unsigned* sh, unsigned k, int tid, int nthreads, int bar_id, unsigned* scratch) {
    const int per = (kHistBins + nthreads - 1) / nthreads;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const int lo = min(tid * per, kHistBins), hi = min(lo + per, kHistBins);
    unsigned mine = 0;
    for (int j = lo; j < hi; ++j) mine += sh[j];
    unsigned incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_down_sync(0xFFFFFFFFu, incl, d);
        if (lane + d < 32) incl += o;
    }
    if (lane == 0) scratch[warp] = incl;
    if (tid == 0) scratch[32] = 0u;
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
    unsigned above_warps = 0;
    for (int u = warp + 1; u < nwarps; ++u) above_warps += scratch[u];
    const unsigned suffix = above_warps + incl, above = suffix - mine;
    if (above < k && suffix >= k) {
        unsigned a = above;
        int b = hi - 1;
        for (; b > lo; --b) {
            if (a + sh[b] >= k) break;
            a += sh[b];
        }
        scratch[32] = (unsigned)b;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
    return scratch[32];
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
// instead of one full pass over the plane (8k for marks of <= 1024 values: on the synthetic 4K / 1080p frames
// and k = 1000 the 64 x 128 block gives the same candidate count as the 128 x 256 one).  A loose bound (noise-like spectra) only costs a
// candidate overflow, which is detected by topk_rank and repaired with the full histogram.
constexpr int kBlockRows = 128, kBlockCols = 256;        // block for long marks
constexpr int kSmallRows = 64, kSmallCols = 128, kSmallMaxK = 1024;   // short marks: 8x the coefficients consumed
constexpr int kBinThreads = 1024;

__global__ void __launch_bounds__(kBinThreads)
topk_block_bin_kernel(const float* __restrict__ planes, long long plane_stride, unsigned w, unsigned h, unsigned k,
                      OrderConsts oc, TopkScratch ts) {
    // ONE CTA per image: shared histogram -> bin, no global histogram / fence / ticket round trips (the block is
    // at most 32k L2-resident coefficients: 32 per thread, 16 independent loads in flight).
    pdl_enter();
    __shared__ unsigned sh[kHistBins];
    const unsigned img = blockIdx.x;
    const float* plane = planes + (long long)img * plane_stride;
    for (int i = threadIdx.x; i < kHistBins; i += blockDim.x) sh[i] = 0;
    const bool small = k <= (unsigned)kSmallMaxK;
    const unsigned br = min(h, (unsigned)(small ? kSmallRows : kBlockRows)), bc = min(w, (unsigned)(small ? kSmallCols : kBlockCols));
    const unsigned total = br * bc;
    __syncthreads();
    constexpr unsigned BATCH = 16;
    for (unsigned e0 = threadIdx.x; e0 < total; e0 += BATCH * kBinThreads) {
        float v[BATCH];
        unsigned p[BATCH];
#pragma unroll
        for (unsigned u = 0; u < BATCH; ++u) {   // independent loads in flight
            const unsigned e = e0 + u * kBinThreads;
            const unsigned r = e / bc, c = e - r * bc;
            const unsigned q = r * w + c;            // local position (w = local line length)
            p[u] = e < total ? flat_index(q, oc) : 0u;
            v[u] = p[u] ? __ldg(plane + q) : 0.f;
        }
#pragma unroll
        for (unsigned u = 0; u < BATCH; ++u)
            if (p[u]) atomicAdd(&sh[order_key(v[u], p[u], oc) >> (32 - kHistBits)], 1u);
    }
    __syncthreads();
    const unsigned b = find_kth_bin(sh, k);
    if (threadIdx.x == 0) ts.sel_bin[img] = b;
}

// ---- 2. collect candidates -----------------------------------------------------------------------
__device__ __forceinline__ void topk_push(unsigned key, unsigned p, unsigned bin_sel, unsigned* count,
                                          unsigned long long* cand) {
    if ((key >> (32 - kHistBits)) >= bin_sel) {
        const unsigned pos = atomicAdd(count, 1u);
        if (pos < (unsigned)kTopkCap) cand[pos] = ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - p);
    }
}

// the exact test for the four coefficients of one 16-byte load (Energy ordering, flat index == position); out of line:
// one call per ~8000 loads, and the scan loop of topk_collect keeps its registers for the loads in flight
__device__ __noinline__ void topk_push4_energy(float4 v, unsigned p, unsigned bin_sel, unsigned* count, unsigned long long* cand) {
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int x = 0; x < 4; ++x)
        if (p + x) topk_push(total_cmp_key(__fmul_rn(e[x], e[x])), p + x, bin_sel, count, cand);
}

__global__ void __launch_bounds__(512)
topk_collect_kernel(const float* __restrict__ planes, long long plane_stride, unsigned n, OrderConsts oc,
                    TopkScratch ts) {
    pdl_enter();
    const unsigned img = blockIdx.y;
    const float* plane = planes + (long long)img * plane_stride;
    const unsigned bin_sel = ts.sel_bin[img];
    if (ts.maxrow && blockIdx.x == 0 && threadIdx.x == 0) ts.maxrow[img] = 0u;
    if (ts.maxcol && blockIdx.x == 0 && threadIdx.x == 0) { ts.maxcol_img[img] = 0u; if (blockIdx.y == 0) *ts.maxcol = 0u; }
    unsigned* count = ts.cand_count + img;
    unsigned long long* cand = ts.cand + (size_t)img * kTopkCap;
    const unsigned n4 = n >> 2;
    const bool vec = ((((size_t)plane) & 15) == 0);
    if (vec) {
        const float4* p4 = (const float4*)plane;
        const unsigned stride = gridDim.x * blockDim.x;
        unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
        if (oc.mode == 0 && oc.t_ld == 0u) {
            // Energy ordering of an unsharded plane (the hot case): key = bits(c*c) | 2^31, so "bin(key) >= bin_sel" is
            // "c*c >= thr" for the float whose bits are (bin_sel - 2^11) << 20 -- two instructions per coefficient instead of
            // ~15, and four independent 16-byte loads in flight per thread.  The few elements that pass (and NaNs, and every
            // element when the bound is degenerate: thr = 0 or a NaN pattern) take the exact path below, so the candidate
            // set is the one the exact test selects.
            const float thr = bin_sel > (1u << (kHistBits - 1)) ? __uint_as_float((bin_sel - (1u << (kHistBits - 1))) << (32 - kHistBits)) : 0.f;
            constexpr int U = 4;
            for (; i < n4; i += U * stride) {
                float4 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const unsigned q = i + u * stride;
                    v[u] = q < n4 ? __ldg(p4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const unsigned q = i + u * stride;
                    const bool lo = (__fmul_rn(v[u].x, v[u].x) < thr) & (__fmul_rn(v[u].y, v[u].y) < thr) &
                                    (__fmul_rn(v[u].z, v[u].z) < thr) & (__fmul_rn(v[u].w, v[u].w) < thr);
                    if (!lo && q < n4) topk_push4_energy(v[u], q << 2, bin_sel, count, cand);
                }
            }
        } else {
            for (; i < n4; i += stride) {
                const float4 v = __ldg(p4 + i);
                const unsigned p = i << 2;
                const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned g = flat_index(p + u, oc);
                    if (g) topk_push(order_key(e[u], g, oc), g, bin_sel, count, cand);
                }
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

// ---- 2''. collect candidates, tile by tile -----------------------------------------------------------
// The forward column pipeline leaves the largest |coefficient| of every column tile (dct_pipe.cuh, PipeArgs::tile_max;
// tile = tile_cols adjacent columns x all rows).  fl(c*c) is monotone in |c|, so a tile whose maximum fails the test
// "c*c >= thr" of topk_collect holds no candidate and is not read: for natural frames the scan shrinks from the whole plane to
// the tiles of the low-frequency block.  One CTA per tile; Energy ordering of an unsharded plane only (the caller checks).
// The CTA clears the word it consumed, which keeps the array zero between calls.
constexpr int kTileSplit = 8;     // CTAs per image (row ranges)
constexpr int kTileMaxTiles = 2048;   // column tiles per image the kernel can list (shared memory); wider frames take topk_collect
__global__ void __launch_bounds__(256)
topk_collect_tiles_kernel(const float* __restrict__ planes, long long plane_stride, unsigned w, unsigned h, unsigned tile_cols, unsigned tiles,
                          TopkScratch ts, const unsigned* __restrict__ tile_max) {
    pdl_enter();
    __shared__ unsigned short live[kTileMaxTiles];
    __shared__ unsigned n_live;
    const unsigned img = blockIdx.y, tid = threadIdx.x;
    const float* plane = planes + (long long)img * plane_stride;
    const unsigned bin_sel = ts.sel_bin[img];
    if (blockIdx.x == 0 && tid == 0) {
        if (ts.maxrow) ts.maxrow[img] = 0u;
        if (ts.maxcol) { ts.maxcol_img[img] = 0u; if (blockIdx.y == 0) *ts.maxcol = 0u; }
    }
    if (tid == 0) n_live = 0u;
    __syncthreads();
    const float thr = bin_sel > (1u << (kHistBits - 1)) ? __uint_as_float((bin_sel - (1u << (kHistBits - 1))) << (32 - kHistBits)) : 0.f;
    // the tiles that can hold a candidate (a NaN threshold or maximum fails the comparison: the tile is read); any order will do
    for (unsigned t = tid; t < tiles; t += blockDim.x) {
        const float m = __uint_as_float(__ldcg(tile_max + (size_t)img * tiles + t));
        if (!(__fmul_rn(m, m) < thr)) live[atomicAdd(&n_live, 1u)] = (unsigned short)t;
    }
    __syncthreads();
    const unsigned nl = n_live;
    unsigned* count = ts.cand_count + img;
    unsigned long long* cand = ts.cand + (size_t)img * kTopkCap;
    const unsigned q4 = tile_cols >> 2;   // 16-byte groups per row of a tile (w % 4 == 0, plane 16-byte aligned)
    const unsigned rows = (h + gridDim.x - 1) / gridDim.x, r0 = blockIdx.x * rows, r1 = min(h, r0 + rows);
    const unsigned per_tile = (r1 > r0 ? r1 - r0 : 0u) * q4, total = per_tile * nl;
    constexpr int U = 4;   // independent loads in flight per thread
    for (unsigned i0 = tid; i0 < total; i0 += U * blockDim.x) {
        float4 v[U];
        unsigned p[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned i = i0 + u * blockDim.x;
            bool ok = i < total;
            unsigned pos = 0xFFFFFFFFu;
            if (ok) {
                const unsigned li = i / per_tile, e = i - li * per_tile;
                const unsigned r = r0 + e / q4, c = (unsigned)live[li] * tile_cols + 4u * (e % q4);
                ok = c < w;
                if (ok) pos = r * w + c;
            }
            p[u] = pos;
            v[u] = ok ? __ldg((const float4*)(plane + pos)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool lo = (__fmul_rn(v[u].x, v[u].x) < thr) & (__fmul_rn(v[u].y, v[u].y) < thr) & (__fmul_rn(v[u].z, v[u].z) < thr) &
                            (__fmul_rn(v[u].w, v[u].w) < thr);
            if (!lo && p[u] != 0xFFFFFFFFu) topk_push4_energy(v[u], p[u], bin_sel, count, cand);
        }
    }
}

// ---- 2'. distributed merge: concatenate the candidate lists gathered from all ranks --------------
__global__ void __launch_bounds__(256)
topk_concat_kernel(const unsigned long long* __restrict__ lists, const unsigned* __restrict__ counts, unsigned n_lists,
                   unsigned list_cap, unsigned list_pitch, TopkScratch ts) {
    pdl_enter();
    __shared__ unsigned off[65];
    if (threadIdx.x == 0) {
        unsigned run = 0, claimed = 0;
        bool lost = false;
        for (unsigned l = 0; l < n_lists; ++l) {
            off[l] = run; run += min(counts[l], list_cap);
            claimed += counts[l];                       // what the ranks found, not what their lists could hold
            lost = lost || counts[l] > list_cap;
        }
        off[n_lists] = run;
        // a list that overflowed on its rank dropped candidates in arbitrary order: the merged order would be wrong even
        // if the surviving entries fit -> report a count above the capacity, which topk_rank flags as overflow
        ts.cand_count[0] = (lost && claimed <= (unsigned)kTopkCap) ? (unsigned)kTopkCap + 1u : max(claimed, run);
    }
    __syncthreads();
    for (unsigned l = 0; l < n_lists; ++l) {
        const unsigned c = min(counts[l], list_cap);
        for (unsigned i = threadIdx.x; i < c; i += blockDim.x)
            if (off[l] + i < (unsigned)kTopkCap) ts.cand[off[l] + i] = lists[(size_t)l * list_pitch + i];
    }
}

// ---- 3. order the candidates, emit the first k indices -- and consume them -----------------------
// The composite keys are unique, so the position of a candidate in the descending order is simply the number
// of candidates with a larger key.  With ~k + one bin of candidates (1049 for k = 1000 on the 4K frame) the
// n^2 comparisons (1.1 M) spread over kRankCtas CTAs take ~1 us -- far less than a single-CTA bitonic network,
// whose 66 dependent exchange steps (shuffles, shared-memory round trips, barriers) cost 12-14 us.
// CTA b ranks candidates [32 b, 32 b + 32) (+ strides of 32 * gridDim.x); 8 thread groups split the comparison
// range.  The thread that learns "candidate p has rank r < k" is also the one that knows everything the next step
// of the fused pipelines needs, so it performs it in place (TopkApply): the embedding of mark value r into
// coefficient p (Writer::embed_watermark, /root/reference/src/algorithm.rs:394-398) or the extraction of value r
// from the base / derived coefficient pair (:556-561).  The last CTA to finish (ticket) reports overflow, clears
// the per-image scratch for the next call and, for the extraction, reduces the 1:1 similarity score (:696-714) in a
// fixed-shape tree (deterministic; within a few ulp of the sequential loop, the contract is 1e-3 relative).
// A frame whose candidate list overflowed (or came up short) is NOT consumed: its index list is filled with
// kBadIndex, its coefficients stay untouched (the inverse transform then returns the unmarked frame), its extracted
// vector is zero and its score NaN -- and the sticky overflow counter tells the host (ssw_ctx_last_topk_fallbacks).
constexpr int kRankThreads = 256, kRankCtas = 64;
constexpr int kRankSlots = 32, kRankParts = kRankThreads / kRankSlots;   // candidates ranked per CTA and pass x thread groups splitting the comparisons
constexpr int kRankSpec = 8;   // keys per thread loaded before the count is known (2048 candidates: the usual list in one round trip)

struct TopkApply {
    int mode;                  // 0: indices only; 1: embed (scatter); 2: extract (gather [+ similarity]);
                               // 3: embed as deltas: out[r] = f(c, w_r) - c, plane untouched, ts.maxrow = max coefficient row (lowrank.cuh)
    unsigned width;            // mode 3: frame width (row of a flat index); mode 1: non-zero = record the largest modified column in ts.maxcol
    int method; float alpha;   // insertion / extraction option 1..3
    float* planes;             // mode 1: coefficient planes, modified in place; mode 2: base planes (read)
    const float* derived;      // mode 2
    long long plane_stride;
    const float* marks;        // mode 1: [img][mark_stride]; mode 2: optional, for the score
    long long mark_stride;
    float* out; long long out_stride;   // mode 2: extracted vectors
    float* sim;                // mode 2: optional scores [img]
};

__global__ void __launch_bounds__(kRankThreads)
topk_rank_kernel(TopkScratch ts, unsigned k, unsigned* __restrict__ idx_out, long long idx_stride, TopkApply ap) {
    pdl_enter();
    extern __shared__ unsigned long long keys[];  // up to kTopkCap candidates
    __shared__ unsigned rank[kRankSlots];
    __shared__ unsigned s_last;
    __shared__ float red[2][kRankThreads / 32];
    const unsigned img = blockIdx.y, tid = threadIdx.x;
    const unsigned long long* cand = ts.cand + (size_t)img * kTopkCap;
    // the head of the list is fetched together with its length (one L2 round trip instead of two on the latency chain
    // of the step); entries at or beyond the length are stale words of the scratch buffer and are never looked at
    static_assert(kRankSpec * kRankThreads <= kTopkCap, "speculative loads stay inside the candidate buffer");
    // (as many as a list of k + one bin usually holds; CTAs whose slots lie beyond that will most likely find no work and skip it)
    const unsigned spec_want = min((unsigned)(kRankSpec * kRankThreads), k + (k >> 2) + 64u);
    const unsigned spec_n = blockIdx.x * (unsigned)kRankSlots < spec_want ? spec_want : 0u;
    unsigned long long spec[kRankSpec];
#pragma unroll
    for (int u = 0; u < kRankSpec; ++u) spec[u] = tid + u * kRankThreads < spec_n ? __ldcg(cand + tid + u * kRankThreads) : 0ull;
    const unsigned total = __ldcg(ts.cand_count + img);
    const unsigned cnt = total < (unsigned)kTopkCap ? total : (unsigned)kTopkCap;
    const bool bad = total > (unsigned)kTopkCap || cnt < k;
    unsigned* out = idx_out + (long long)img * idx_stride;
    float* plane = ap.planes ? ap.planes + (long long)img * ap.plane_stride : nullptr;
    const float* dplane = ap.derived ? ap.derived + (long long)img * ap.plane_stride : nullptr;
    const float* mk = ap.marks ? ap.marks + (long long)img * ap.mark_stride : nullptr;
    float* ext = ap.out ? ap.out + (long long)img * ap.out_stride : nullptr;
    if (ts.tile_max && blockIdx.x == gridDim.x - 1)   // (a CTA that rarely has candidates to rank)
        for (unsigned i = tid; i < ts.tile_count; i += kRankThreads) ts.tile_max[(size_t)img * ts.tile_count + i] = 0u;
    if (bad) {
        for (unsigned r = blockIdx.x * kRankThreads + tid; r < k; r += gridDim.x * kRankThreads) {
            out[r] = kBadIndex;
            if (ap.mode >= 2) ext[r] = 0.f;
        }
        if (ap.mode == 3 && blockIdx.x == 0 && tid == 0) ts.maxrow[img] = 0xFFFFFFFFu;
    } else if (blockIdx.x * (unsigned)kRankSlots < cnt) {
#pragma unroll
        for (int u = 0; u < kRankSpec; ++u)
            if (tid + u * kRankThreads < spec_n) keys[tid + u * kRankThreads] = spec[u];
        for (unsigned j = spec_n + tid; j < cnt; j += kRankThreads) keys[j] = __ldcg(cand + j);
        const unsigned slot = tid % (unsigned)kRankSlots, part = tid / (unsigned)kRankSlots;
        const unsigned j0 = (unsigned)(((unsigned long long)cnt * part) / kRankParts), j1 = (unsigned)(((unsigned long long)cnt * (part + 1)) / kRankParts);
        for (unsigned base = blockIdx.x * (unsigned)kRankSlots; base < cnt; base += gridDim.x * (unsigned)kRankSlots) {
            if (tid < (unsigned)kRankSlots) rank[tid] = 0u;
            __syncthreads();   // keys (first round) and rank[] ready
            const unsigned i = base + slot;
            if (i < cnt) {
                const unsigned long long mine = keys[i];
                unsigned above = 0;
#pragma unroll 8
                for (unsigned j = j0; j < j1; ++j) above += (keys[j] > mine) ? 1u : 0u;   // broadcast reads
                atomicAdd(&rank[slot], above);
            }
            __syncthreads();
            if (tid < (unsigned)kRankSlots && i < cnt && rank[tid] < k) {
                const unsigned r = rank[tid], p = 0xFFFFFFFFu - (unsigned)(keys[i] & 0xFFFFFFFFull);
                out[r] = p;
                if (ap.mode == 1) {
                    plane[p] = insert_fn(ap.method, ap.alpha, plane[p], __ldg(mk + r));
                    if (ap.width) { atomicMax(ts.maxcol, p % ap.width); atomicMax(ts.maxcol_img + img, p % ap.width); }
                }
                else if (ap.mode == 2) ext[r] = extract_fn(ap.method, ap.alpha, plane[p], dplane[p]);
                else if (ap.mode == 3) {
                    const float c0 = plane[p];
                    ext[r] = __fsub_rn(insert_fn(ap.method, ap.alpha, c0, __ldg(mk + r)), c0);
                    atomicMax(ts.maxrow + img, p / ap.width);
                }
            }
            __syncthreads();
        }
    }
    __threadfence();   // this thread's results are visible device-wide before the CTA takes its ticket
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ts.ticket + img, 1u) == gridDim.x - 1) ? 1u : 0u;   // every CTA of the image has read the counters
    __syncthreads();
    if (!s_last) return;
    if (tid == 0) {
        if (bad) atomicAdd(ts.overflow, 1u);
        ts.cand_count[img] = 0;
        ts.ticket[img] = 0;
        ts.sel_bin[img] = 0;   // clears the "bin published" flag of the collecting column pipeline (dct_pipe.cuh, PipeArgs::collect)
    }
    if (ap.mode == 2 && ap.sim) {
        __threadfence();   // the other CTAs' extracted values (their fence + ticket precede ours)
        float num = 0.f, den = 0.f;
        if (!bad)
            for (unsigned i = tid; i < k; i += kRankThreads) {
                const float e = __ldcg(ext + i);
                num = fmaf(e, __ldg(mk + i), num);
                den = fmaf(e, e, den);
            }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            num += __shfl_xor_sync(0xFFFFFFFFu, num, d);
            den += __shfl_xor_sync(0xFFFFFFFFu, den, d);
        }
        if ((tid & 31u) == 0u) { red[0][tid >> 5] = num; red[1][tid >> 5] = den; }
        __syncthreads();
        if (tid == 0) {
            float a = 0.f, q = 0.f;
            for (int wv = 0; wv < kRankThreads / 32; ++wv) { a += red[0][wv]; q += red[1][wv]; }
            ap.sim[img] = bad ? __int_as_float(0x7FC00000) : __fdiv_rn(a, __fsqrt_rn(q));
        }
    }
}

}  // namespace ssw
