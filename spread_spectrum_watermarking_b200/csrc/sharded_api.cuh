// Row-sharded single frames behind the C ABI (BASELINE.json configs[3]); included by ssw_api.cu.
//
// The reference holds one whole frame on one core (`Writer::new`, /root/reference/src/algorithm.rs:295-316).  Here one
// process per GPU owns the pixel rows [g*H/G, (g+1)*H/G) of the frame and -- after the exchange between the two passes
// of the separable transform (/root/reference/src/dct2d.rs:93-98) -- the coefficient columns [g*W/G, (g+1)*W/G), kept
// TRANSPOSED (local plane P = [W/G][H]) so that both passes run over contiguous lines.
//
// The exchange is not a collective: every rank maps its peers' planes (cudaIpc*, NVLink / NVSwitch peer access) and the
// block transposes STORE their tiles straight into the owner's plane, already in the layout its next pass reads
// (SURVEY 8(e) "preferred end state").  The local pass runs in slices; the push of slice i (side stream) overlaps the
// line kernels of slice i+1.  Per 2-D transform a rank sends (G-1)/G of its W*H*4/G plane bytes over NVLink.
// NCCL (loaded with dlopen: libssw has no link-time dependency on it) carries only the small pieces: the exchange of
// the IPC handles, the barrier that closes a push phase, the max / all-gather of the distributed top-k and the sum of
// the extracted vector.
#pragma once
#include <dlfcn.h>

// ---- the handful of NCCL entry points, resolved at run time (ABI of NCCL 2.x) ---------------------------------------
namespace ssw_nccl {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5, ncclFloat32 = 7 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };
struct Api {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
};
static Api* api() {
    static Api a;
    if (a.lib || !a.error.empty()) return &a;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) { a.error = std::string("NCCL is not available (dlopen libnccl.so.2): ") + dlerror(); return &a; }
    auto sym = [&](const char* n) { void* p = dlsym(a.lib, n); if (!p) a.error = std::string("missing NCCL symbol ") + n; return p; };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    return &a;
}
}  // namespace ssw_nccl

#define CKN(call)                                                                                        \
    do {                                                                                                 \
        int r_ = (call);                                                                                 \
        if (r_ != ssw_nccl::ncclSuccess)                                                                 \
            return fail(SSW_ERR_CUDA, std::string(#call) + ": " + ssw_nccl::api()->GetErrorString(r_));  \
    } while (0)

// ---- block transpose that stores into the owners' planes --------------------------------------------------------------
// src: nr local lines of length G*cb (leading dimension ld).  Block g = columns [g*cb, (g+1)*cb) goes, transposed, to
// rank g:  dst[g][c * dst_ld + dst_off + r] = src[r * ld + g*cb + c]   (c < cb, r < nr).
// 64 x 32 tiles: a warp stores 64 consecutive floats (256 bytes) of one destination line.
struct PeerPtrs { float* p[16]; };

// Destination order: rank r sends to r+1, r+2, ... (mod G) and copies its own block last, so at any moment every rank's
// stores go to a different owner (no incast on one NVLink ingress while the others idle).
__global__ void __launch_bounds__(256)
transpose_push_kernel(const float* __restrict__ src, unsigned nr, unsigned cb, long long ld, PeerPtrs dst, long long dst_ld, long long dst_off,
                      unsigned rank) {
    pdl_enter();
    __shared__ float tile[64][33];
    const unsigned g = (blockIdx.z + rank + 1u) % gridDim.z;
    const float* s = src + (long long)g * cb;
    float* d = dst.p[g];
    const unsigned c0 = blockIdx.x * 32, r0 = blockIdx.y * 64;
    const unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 warps
#pragma unroll
    for (int i = 0; i < 64; i += 8) {
        const unsigned r = r0 + ty + i, c = c0 + tx;
        if (r < nr && c < cb) tile[ty + i][tx] = __ldg(s + (long long)r * ld + c);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const unsigned c = c0 + ty + i;
        if (c < cb) {
            float* line = d + (long long)c * dst_ld + dst_off + r0;
            if (r0 + tx < nr) line[tx] = tile[tx][ty + i];
            if (r0 + 32 + tx < nr) line[32 + tx] = tile[32 + tx][ty + i];
        }
    }
}

// ---- the per-rank object ------------------------------------------------------------------------------------------------
struct ssw_sharded {
    CtxRef ctx;
    int rank = 0, world = 1;
    uint32_t w = 0, h = 0, hb = 0, wb = 0;   // frame, rows / columns of this rank
    ssw_nccl::ncclComm_t comm = nullptr;
    // shared (peer-mapped) buffers: R = [hb][W] coefficient rows (target of the inverse exchange), P[0..1] = [wb][H]
    // transposed coefficient columns (targets of the forward exchange; two: base and derived frame of an extraction)
    float* R = nullptr;
    float* P[2] = {nullptr, nullptr};
    PeerPtrs peerR{}, peerP[2]{};
    std::vector<void*> opened;            // cudaIpcOpenMemHandle results (closed in destroy)
    cudaStream_t push = nullptr;          // side stream of the pushes
    std::vector<cudaEvent_t> ev;
    cudaEvent_t ev_done = nullptr;
    unsigned* d_bin = nullptr;            // [1]
    unsigned long long* d_cand = nullptr; // [SSW_TOPK_CAP + 1]: candidate list, then the count (one all-gather moves both)
    unsigned long long* d_lists = nullptr;// [world][SSW_TOPK_CAP + 1]
    unsigned* d_counts = nullptr;         // [world]
    unsigned* d_idx = nullptr;            // [kTopkCap / 2]
    unsigned* d_overflow = nullptr;       // [1] sticky
    float* d_barrier = nullptr;           // [1]
    float* d_part = nullptr;              // [kTopkCap / 2]
    size_t last_k = 0;
    int chunks = 1;
};

extern "C" int ssw_sharded_unique_id(void* id_out) {
    if (!id_out) return fail(SSW_ERR_INVALID, "NULL argument");
    ssw_nccl::Api* n = ssw_nccl::api();
    if (!n->error.empty()) return fail(SSW_ERR_UNSUPPORTED, n->error);
    ssw_nccl::ncclUniqueId id;
    CKN(n->GetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof(id));
    return SSW_OK;
}

static int sharded_barrier(ssw_sharded* s) {   // closes a push phase: every rank's stores have been performed
    if (s->world == 1) return SSW_OK;
    CKN(ssw_nccl::api()->AllReduce(s->d_barrier, s->d_barrier, 1, ssw_nccl::ncclFloat32, ssw_nccl::ncclSum, s->comm, s->ctx->stream));
    return SSW_OK;
}

extern "C" int ssw_sharded_destroy(ssw_sharded* s) {
    if (!s) return SSW_OK;
    ssw_ctx* c = s->ctx;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (s->push) { cudaStreamSynchronize(s->push); }
    if (s->comm && s->world > 1) {
        // nobody unmaps or frees while a peer may still be storing into these buffers
        ssw_nccl::api()->AllReduce(s->d_barrier, s->d_barrier, 1, ssw_nccl::ncclFloat32, ssw_nccl::ncclSum, s->comm, c->stream);
        cudaStreamSynchronize(c->stream);
    }
    for (void* p : s->opened) cudaIpcCloseMemHandle(p);
    if (s->comm) ssw_nccl::api()->CommDestroy(s->comm);
    cudaFree(s->R); cudaFree(s->P[0]); cudaFree(s->P[1]);
    cudaFree(s->d_bin); cudaFree(s->d_cand); cudaFree(s->d_lists); cudaFree(s->d_counts); cudaFree(s->d_idx);
    cudaFree(s->d_overflow); cudaFree(s->d_barrier); cudaFree(s->d_part);
    for (cudaEvent_t e : s->ev) cudaEventDestroy(e);
    if (s->ev_done) cudaEventDestroy(s->ev_done);
    if (s->push) cudaStreamDestroy(s->push);
    delete s;
    return SSW_OK;
}

extern "C" int ssw_sharded_create(ssw_ctx* c, const void* id, int rank, int world, uint32_t width, uint32_t height, ssw_sharded** out) {
    if (!c || !out || (world > 1 && !id)) return fail(SSW_ERR_INVALID, "NULL argument");
    if (world < 1 || world > 16 || rank < 0 || rank >= world) return fail(SSW_ERR_INVALID, "rank / world out of range (1..16 ranks)");
    CKS(check_dims(width, height));
    if (width % (uint32_t)world || height % (uint32_t)world)
        return fail(SSW_ERR_UNSUPPORTED, "sharded frames need width and height divisible by the number of ranks");
    CKS(ctx_bind(c));
    auto s = std::make_unique<ssw_sharded>();
    s->ctx = c; s->rank = rank; s->world = world; s->w = width; s->h = height;
    s->hb = height / (uint32_t)world; s->wb = width / (uint32_t)world;
    const size_t plane_bytes = (size_t)s->hb * width * sizeof(float);   // == wb * height * 4
    CK(cudaMalloc(&s->R, plane_bytes));
    CK(cudaMalloc(&s->P[0], plane_bytes));
    CK(cudaMalloc(&s->P[1], plane_bytes));
    CK(cudaMalloc(&s->d_bin, sizeof(unsigned)));
    CK(cudaMalloc(&s->d_cand, (kTopkCap + 1) * sizeof(unsigned long long)));
    CK(cudaMalloc(&s->d_lists, (size_t)world * (kTopkCap + 1) * sizeof(unsigned long long)));
    CK(cudaMalloc(&s->d_counts, world * sizeof(unsigned)));
    CK(cudaMalloc(&s->d_idx, (kTopkCap / 2) * sizeof(unsigned)));
    CK(cudaMalloc(&s->d_overflow, sizeof(unsigned)));
    CK(cudaMalloc(&s->d_barrier, sizeof(float)));
    CK(cudaMalloc(&s->d_part, (kTopkCap / 2) * sizeof(float)));
    CK(cudaMemset(s->d_overflow, 0, sizeof(unsigned)));
    CK(cudaMemset(s->d_barrier, 0, sizeof(float)));
    {   // the pushes run beside the line kernels of the next slice: highest priority, so that their CTAs take the SM slots
        // the line kernel frees first (the exchange, not the arithmetic, is the longer of the two)
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&s->push, cudaStreamNonBlocking, hi));
    }
    CK(cudaEventCreateWithFlags(&s->ev_done, cudaEventDisableTiming));
    // slices of a local pass whose push overlaps the next slice's line kernels
    s->chunks = 1;
    if (world > 1)
        for (int cnum : {4, 2})
            if (s->hb % (2 * cnum) == 0 && s->wb % (2 * cnum) == 0 && s->hb / cnum >= 64 && s->wb / cnum >= 64) { s->chunks = cnum; break; }
    if (const char* e = getenv("SSW_SHARD_CHUNKS")) {
        const int v = atoi(e);
        if (v >= 1 && s->hb % (2 * v) == 0 && s->wb % (2 * v) == 0) s->chunks = v;
    }
    for (int i = 0; i < s->chunks; ++i) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s->ev.push_back(e);
    }
    s->peerR.p[rank] = s->R; s->peerP[0].p[rank] = s->P[0]; s->peerP[1].p[rank] = s->P[1];
    if (world > 1) {
        ssw_nccl::Api* n = ssw_nccl::api();
        if (!n->error.empty()) return fail(SSW_ERR_UNSUPPORTED, n->error);
        ssw_nccl::ncclUniqueId uid;
        std::memcpy(&uid, id, sizeof(uid));
        CKN(n->CommInitRank(&s->comm, world, uid, rank));
        // exchange the IPC handles of the three shared buffers (all-gather of 3 x 64 bytes per rank)
        struct Handles { cudaIpcMemHandle_t h[3]; };
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        Handles mine;
        CK(cudaIpcGetMemHandle(&mine.h[0], s->R));
        CK(cudaIpcGetMemHandle(&mine.h[1], s->P[0]));
        CK(cudaIpcGetMemHandle(&mine.h[2], s->P[1]));
        Handles *d_mine = nullptr, *d_all = nullptr;
        CK(cudaMalloc(&d_mine, sizeof(Handles)));
        CK(cudaMalloc(&d_all, sizeof(Handles) * world));
        CK(cudaMemcpyAsync(d_mine, &mine, sizeof(Handles), cudaMemcpyHostToDevice, c->stream));
        CKN(n->AllGather(d_mine, d_all, sizeof(Handles), ssw_nccl::ncclChar, s->comm, c->stream));
        std::vector<Handles> all(world);
        CK(cudaMemcpyAsync(all.data(), d_all, sizeof(Handles) * world, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        cudaFree(d_mine); cudaFree(d_all);
        for (int g = 0; g < world; ++g) {
            if (g == rank) continue;
            void* ptr[3];
            for (int b = 0; b < 3; ++b) {
                cudaError_t e = cudaIpcOpenMemHandle(&ptr[b], all[g].h[b], cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess)
                    return fail(SSW_ERR_CUDA, std::string("cudaIpcOpenMemHandle (peer plane of rank ") + std::to_string(g) + "): " + cudaGetErrorString(e));
                s->opened.push_back(ptr[b]);
            }
            s->peerR.p[g] = (float*)ptr[0]; s->peerP[0].p[g] = (float*)ptr[1]; s->peerP[1].p[g] = (float*)ptr[2];
        }
        CKS(sharded_barrier(s.get()));
        CK(cudaStreamSynchronize(c->stream));
    }
    *out = s.release();
    return SSW_OK;
}

// one local pass in slices + the push of every slice into the owners' planes; ends with the barrier
//   forward (to_columns): lines = this rank's hb pixel rows (length W) -> R slices -> peers' P[which] ([wb][H], offset rank*hb)
//   inverse             : lines = this rank's wb coefficient columns P[which] (length H), DCT-III in place -> peers' R ([hb][W], offset rank*wb)
static int sharded_exchange(ssw_sharded* s, bool forward, const uint8_t* rows_rgb8, int which) {
    ssw_ctx* c = s->ctx;
    const int G = s->world;
    const uint32_t n_lines = forward ? s->hb : s->wb, n = forward ? s->w : s->h;
    const uint32_t cb = forward ? s->wb : s->hb;                 // block width = lines the destination owns
    const uint32_t lc = n_lines / (uint32_t)s->chunks;
    const PeerPtrs& dst = forward ? s->peerP[which] : s->peerR;
    const long long dst_ld = forward ? (long long)s->h : (long long)s->w;
    for (int ci = 0; ci < s->chunks; ++ci) {
        const size_t l0 = (size_t)ci * lc;
        float* lines;
        if (forward) {
            lines = s->R + l0 * n;
            CKS(run_rows_forward(c, PIX_RGB8, rows_rgb8 + l0 * n * 3, (int)n, (int)lc, 1, lines, 1.f, 1.f));
        } else {
            lines = s->P[which] + l0 * n;
            CKS(run_rows_inverse(c, lines, PIX_PLANE, nullptr, (int)n, (int)lc, 1, PIX_PLANE, lines, 1.0f));
        }
        // per-kernel profiling times every kernel alone on the context stream; otherwise the push runs beside the next slice
        cudaStream_t ps = c->profiling ? c->stream : s->push;
        if (!c->profiling) {
            CK(cudaEventRecord(s->ev[ci], c->stream));
            CK(cudaStreamWaitEvent(s->push, s->ev[ci], 0));
        }
        {
            KScope ks(c, "transpose_push");
            const dim3 grid((cb + 31) / 32, (lc + 63) / 64, (unsigned)G);
            transpose_push_kernel<<<grid, 256, 0, ps>>>(lines, lc, cb, (long long)n, dst, dst_ld,
                                                        (long long)s->rank * n_lines + (long long)l0, (unsigned)s->rank);
        }
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(s->ev_done, s->push));
    CK(cudaStreamWaitEvent(c->stream, s->ev_done, 0));
    return sharded_barrier(s);
}

// forward transform of this rank's rows into P[which] (its coefficient columns, transposed)
static int sharded_forward(ssw_sharded* s, const uint8_t* rows, int which) {
    CKS(sharded_exchange(s, true, rows, which));
    // column pass: contiguous lines of length H, in place
    return run_rows_forward(s->ctx, PIX_PLANE, s->P[which], (int)s->h, (int)s->wb, 1, s->P[which], 1.f, 1.f);
}

// distributed ordered top-k of P[0] (obtain_indices_by_function, src/algorithm.rs:200-221) -> s->d_idx on every rank
static int sharded_topk(ssw_sharded* s, int ordering, size_t k) {
    ssw_ctx* c = s->ctx;
    ssw_nccl::Api* n = ssw_nccl::api();
    const ssw_shard sh{s->w, s->h, (uint32_t)s->rank * s->wb, s->wb};
    CKS(ssw_shard_topk_bin_dev(c, s->P[0], &sh, ordering, k, s->d_bin));
    if (s->world > 1) CKN(n->AllReduce(s->d_bin, s->d_bin, 1, ssw_nccl::ncclUint32, ssw_nccl::ncclMax, s->comm, c->stream));
    unsigned* count = (unsigned*)(s->d_cand + kTopkCap);   // the count rides behind the list: one all-gather
    CKS(ssw_shard_topk_collect_dev(c, s->P[0], &sh, ordering, s->d_bin, (uint64_t*)s->d_cand, count));
    if (s->world > 1) {
        CKN(n->AllGather(s->d_cand, s->d_lists, (kTopkCap + 1) * sizeof(unsigned long long), ssw_nccl::ncclChar, s->comm, c->stream));
    } else {
        CK(cudaMemcpyAsync(s->d_lists, s->d_cand, (kTopkCap + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
    }
    // counts[g] = low word of lists[g][kTopkCap]
    CK(cudaMemcpy2DAsync(s->d_counts, sizeof(unsigned), (const unsigned*)(s->d_lists + kTopkCap), (kTopkCap + 1) * sizeof(unsigned long long),
                         sizeof(unsigned), (size_t)s->world, cudaMemcpyDeviceToDevice, c->stream));
    // merge: lists are [world][kTopkCap + 1] -- the concat kernel takes the list pitch
    CKS(ensure_topk_scratch(c, 1));
    TopkScratch ts = c->ts;
    ts.overflow = s->d_overflow;
    {
        KScope ks(c, "topk_concat");
        launch_pdl(c, topk_concat_kernel, 1, 256, 0, c->stream, (const unsigned long long*)s->d_lists, (const unsigned*)s->d_counts,
                   (unsigned)s->world, (unsigned)kTopkCap, (unsigned)(kTopkCap + 1), ts);
    }
    const void* key = (const void*)topk_rank_kernel;
    const int smem = kTopkCap * (int)sizeof(unsigned long long);
    if (c->smem_attr.find(key) == c->smem_attr.end()) {
        CK(cudaFuncSetAttribute(topk_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        c->smem_attr[key] = smem;
    }
    TopkApply none;
    std::memset(&none, 0, sizeof(none));
    { KScope ks(c, "topk_rank"); launch_pdl(c, topk_rank_kernel, dim3(kRankCtas, 1), kRankThreads, smem, c->stream, ts, (unsigned)k, s->d_idx, 0, none); }
    CK(cudaGetLastError());
    s->last_k = k;
    return SSW_OK;
}

// Writer::new(img, cfg).mark(&[mark]).into_rgb8() (src/algorithm.rs:295-379) on this rank's rows
extern "C" int ssw_sharded_embed_rgb8_dev(ssw_sharded* s, const uint8_t* rows_dev, const ssw_config* cfg, const float* mark_dev,
                                          size_t n, uint8_t* out_rows_dev) {
    if (!s || !rows_dev || !out_rows_dev || (n && !mark_dev)) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_cfg(cfg));
    ssw_ctx* c = s->ctx;
    CKS(ctx_bind(c));
    const size_t k = std::min(n, (size_t)s->w * s->h - 1);   // zip truncation, src/algorithm.rs:396
    if (k > (size_t)kTopkCap / 2) return fail(SSW_ERR_UNSUPPORTED, "sharded frames support mark lengths up to 4096");
    CKS(sharded_forward(s, rows_dev, 0));
    if (k) {
        CKS(sharded_topk(s, cfg->ordering, k));
        const ssw_shard sh{s->w, s->h, (uint32_t)s->rank * s->wb, s->wb};
        CKS(ssw_shard_embed_dev(c, s->P[0], &sh, s->d_idx, k, mark_dev, k, 1, nullptr, cfg));
    }
    // inverse: columns (DCT-III along y) -> push back into the row owners' R -> rows (DCT-III along x, x4/(WH), colour)
    CKS(sharded_exchange(s, false, nullptr, 0));
    return run_rows_inverse(c, s->R, PIX_RGB8, rows_dev, (int)s->w, (int)s->hb, 1, PIX_RGB8, out_rows_dev,
                            4.0f / (float)((size_t)s->w * (size_t)s->h));
}

// Reader::base + Reader::derived + extract (src/algorithm.rs:462-562); every rank receives the whole vector
extern "C" int ssw_sharded_extract_rgb8_dev(ssw_sharded* s, const uint8_t* base_rows_dev, const uint8_t* derived_rows_dev,
                                            const ssw_config* cfg, size_t n, float* extracted_dev) {
    if (!s || !base_rows_dev || !derived_rows_dev || !extracted_dev) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_cfg(cfg));
    if (n >= (size_t)s->w * s->h) return fail(SSW_ERR_INVALID, "Desired extraction length exceeds available coefficients.");
    if (n > (size_t)kTopkCap / 2) return fail(SSW_ERR_UNSUPPORTED, "sharded frames support mark lengths up to 4096");
    if (n == 0) return SSW_OK;
    ssw_ctx* c = s->ctx;
    CKS(ctx_bind(c));
    CKS(sharded_forward(s, base_rows_dev, 0));
    CKS(sharded_forward(s, derived_rows_dev, 1));
    CKS(sharded_topk(s, cfg->ordering, n));
    const ssw_shard sh{s->w, s->h, (uint32_t)s->rank * s->wb, s->wb};
    float* part = s->world > 1 ? s->d_part : extracted_dev;
    CKS(ssw_shard_extract_dev(c, s->P[0], s->P[1], &sh, s->d_idx, n, cfg, part));
    if (s->world > 1)   // every index has exactly one owner: the sum assembles the vector
        CKN(ssw_nccl::api()->AllReduce(part, extracted_dev, n, ssw_nccl::ncclFloat32, ssw_nccl::ncclSum, s->comm, c->stream));
    return SSW_OK;
}

// the first n ordered indices of the last embed / extract (flat r*W + c, identical on every rank); synchronises
extern "C" int ssw_sharded_indices(ssw_sharded* s, uint32_t* out_host, size_t n) {
    if (!s || (!out_host && n)) return fail(SSW_ERR_INVALID, "NULL argument");
    if (n > s->last_k) return fail(SSW_ERR_STATE, "more indices requested than the last embed / extract ordered");
    CKS(ctx_bind(s->ctx));
    CK(cudaMemcpyAsync(out_host, s->d_idx, n * sizeof(unsigned), cudaMemcpyDeviceToHost, s->ctx->stream));
    CK(cudaStreamSynchronize(s->ctx->stream));
    return SSW_OK;
}

// this rank's coefficient columns of the base (which = 0) / derived (1) frame, transposed [W/G][H]; synchronises
extern "C" int ssw_sharded_coefficients(ssw_sharded* s, int which, float* out_host) {
    if (!s || !out_host || which < 0 || which > 1) return fail(SSW_ERR_INVALID, "bad argument");
    CKS(ctx_bind(s->ctx));
    CK(cudaMemcpyAsync(out_host, s->P[which], (size_t)s->wb * s->h * sizeof(float), cudaMemcpyDeviceToHost, s->ctx->stream));
    CK(cudaStreamSynchronize(s->ctx->stream));
    return SSW_OK;
}

// sticky candidate-overflow flag of the distributed top-k (noise-like spectrum); read-and-clear, synchronises
extern "C" int ssw_sharded_overflow(ssw_sharded* s, int* overflowed) {
    if (!s || !overflowed) return fail(SSW_ERR_INVALID, "NULL argument");
    ssw_ctx* c = s->ctx;
    CKS(ctx_bind(c));
    CK(cudaMemcpyAsync(c->h_flag, s->d_overflow, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemsetAsync(s->d_overflow, 0, sizeof(unsigned), c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *overflowed = *c->h_flag ? 1 : 0;
    return SSW_OK;
}
