// General (any k, any spectrum) coefficient ordering: a stable LSD radix sort of ALL w*h-1
// (key, index) pairs -- literally the reference's `sort_by` over every AC coefficient
// (/root/reference/src/algorithm.rs:204-205), used when the requested length exceeds what the
// single-CTA candidate sort holds or when the fast path reports a candidate overflow (frames with
// thousands of exactly tied energies, e.g. constant or synthetic-pattern images).
//
// Sorting ~key ascending with a *stable* sort from an index-ascending start order reproduces
// "descending by total_cmp, ties keep ascending index".  4 passes of 8 bits; every warp owns a
// contiguous segment so that per-digit ranks preserve input order.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "select_kernels.cuh"

namespace ssw {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortWarps = 8;  // warps per CTA

__global__ void gs_init_kernel(const float* __restrict__ plane, unsigned n, OrderConsts oc,
                               unsigned* __restrict__ keys, unsigned* __restrict__ vals) {
    // element e (0-based) <-> coefficient p = e + 1
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e + 1 < n; e += gridDim.x * blockDim.x) {
        const unsigned p = e + 1;
        keys[e] = ~order_key(plane[p], p, oc);
        vals[e] = p;
    }
}

// counts[digit][segment]
__global__ void __launch_bounds__(kSortWarps * 32)
gs_hist_kernel(const unsigned* __restrict__ keys, unsigned m, unsigned seg_len, unsigned nseg, int shift,
               unsigned* __restrict__ counts) {
    __shared__ unsigned sh[kSortWarps][kRadix];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned seg = blockIdx.x * kSortWarps + warp;
    for (int i = lane; i < kRadix; i += 32) sh[warp][i] = 0;
    __syncwarp();
    if (seg < nseg) {
        const unsigned lo = seg * seg_len, hi = min(m, lo + seg_len);
        for (unsigned i = lo + lane; i < hi; i += 32) atomicAdd(&sh[warp][(keys[i] >> shift) & (kRadix - 1)], 1u);
    }
    __syncwarp();
    if (seg < nseg)
        for (int d = lane; d < kRadix; d += 32) counts[(size_t)d * nseg + seg] = sh[warp][d];
}

// exclusive scan of `len` counters by a single CTA (len <= a few million)
__global__ void __launch_bounds__(1024) gs_scan_kernel(unsigned* __restrict__ counts, size_t len) {
    __shared__ unsigned part[1024];
    const size_t per = (len + blockDim.x - 1) / blockDim.x;
    const size_t lo = min(len, (size_t)threadIdx.x * per), hi = min(len, lo + per);
    unsigned acc = 0;
    for (size_t i = lo; i < hi; ++i) acc += counts[i];
    part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned run = 0;
        for (unsigned t = 0; t < blockDim.x; ++t) { const unsigned v = part[t]; part[t] = run; run += v; }
    }
    __syncthreads();
    unsigned run = part[threadIdx.x];
    for (size_t i = lo; i < hi; ++i) { const unsigned v = counts[i]; counts[i] = run; run += v; }
}

__global__ void __launch_bounds__(kSortWarps * 32)
gs_scatter_kernel(const unsigned* __restrict__ keys, const unsigned* __restrict__ vals, unsigned m,
                  unsigned seg_len, unsigned nseg, int shift, const unsigned* __restrict__ offsets,
                  unsigned* __restrict__ keys_out, unsigned* __restrict__ vals_out) {
    __shared__ unsigned off[kSortWarps][kRadix];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned seg = blockIdx.x * kSortWarps + warp;
    if (seg >= nseg) return;
    for (int d = lane; d < kRadix; d += 32) off[warp][d] = offsets[(size_t)d * nseg + seg];
    __syncwarp();
    const unsigned lo = seg * seg_len, hi = min(m, lo + seg_len);
    for (unsigned base = lo; base < hi; base += 32) {
        const unsigned i = base + lane;
        const bool valid = i < hi;
        unsigned key = 0, val = 0;
        if (valid) { key = keys[i]; val = vals[i]; }
        const unsigned d = (key >> shift) & (kRadix - 1);
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, valid ? d : (0x100u + lane));
        const unsigned rank = __popc(peers & ((1u << lane) - 1u));
        unsigned pos = 0;
        if (valid) pos = off[warp][d] + rank;
        __syncwarp();
        if (valid && rank == 0) off[warp][d] += __popc(peers);
        __syncwarp();
        if (valid) { keys_out[pos] = key; vals_out[pos] = val; }
    }
}

struct GeneralSelect {
    unsigned* buf = nullptr;  // keys0 | vals0 | keys1 | vals1 | counts
    size_t cap_m = 0, cap_counts = 0;
    std::string error;

    void release() {
        if (buf) cudaFree(buf);
        buf = nullptr;
        cap_m = cap_counts = 0;
    }

    // returns ssw_status-compatible code (0 ok, -2 CUDA error)
    int run(cudaStream_t stream, const float* d_plane, unsigned n, const OrderConsts& oc, size_t k,
            unsigned* d_idx, uint64_t* launches) {
        const unsigned m = n - 1;
        if (m == 0 || k == 0) return 0;
        unsigned seg_len = 2048;
        while ((size_t)(m + seg_len - 1) / seg_len > 16384) seg_len *= 2;
        const unsigned nseg = (m + seg_len - 1) / seg_len;
        const size_t ncount = (size_t)kRadix * nseg;
        if (m > cap_m || ncount > cap_counts) {
            cudaStreamSynchronize(stream);
            release();
            cudaError_t e = cudaMalloc(&buf, ((size_t)m * 4 + ncount) * sizeof(unsigned));
            if (e != cudaSuccess) { error = std::string("general select scratch: ") + cudaGetErrorString(e); return -2; }
            cap_m = m; cap_counts = ncount;
        }
        unsigned* keys[2] = {buf, buf + 2 * (size_t)cap_m};
        unsigned* vals[2] = {buf + (size_t)cap_m, buf + 3 * (size_t)cap_m};
        unsigned* counts = buf + 4 * (size_t)cap_m;
        gs_init_kernel<<<1184, 256, 0, stream>>>(d_plane, n, oc, keys[0], vals[0]);
        const unsigned blocks = (nseg + kSortWarps - 1) / kSortWarps;
        int cur = 0;
        for (int pass = 0; pass < 32 / kRadixBits; ++pass) {
            const int shift = pass * kRadixBits;
            gs_hist_kernel<<<blocks, kSortWarps * 32, 0, stream>>>(keys[cur], m, seg_len, nseg, shift, counts);
            gs_scan_kernel<<<1, 1024, 0, stream>>>(counts, ncount);
            gs_scatter_kernel<<<blocks, kSortWarps * 32, 0, stream>>>(keys[cur], vals[cur], m, seg_len, nseg, shift,
                                                                    counts, keys[cur ^ 1], vals[cur ^ 1]);
            cur ^= 1;
        }
        if (launches) *launches += 1 + 3 * (32 / kRadixBits);
        cudaError_t e = cudaMemcpyAsync(d_idx, vals[cur], k * sizeof(unsigned), cudaMemcpyDeviceToDevice, stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { error = std::string("general select: ") + cudaGetErrorString(e); return -2; }
        return 0;
    }
};

}  // namespace ssw
