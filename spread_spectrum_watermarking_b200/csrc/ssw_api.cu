// libssw: C ABI over the sm_100a kernels (see include/ssw.h for the reference items each entry
// point mirrors).  Host orchestration only -- no arithmetic of the hot path runs on the CPU.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/ssw.h"
#include "dct_kernels.cuh"
#include "fast_dispatch.h"
#include "dct_pipe.cuh"
#include <atomic>
#include "mark_kernels.cuh"
#include "select_kernels.cuh"
#include "select_general.cuh"
#include "lowrank.cuh"

using namespace ssw;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(SSW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)
#define CKS(call)                                                                                  \
    do {                                                                                           \
        int s_ = (call);                                                                           \
        if (s_ != SSW_OK) return s_;                                                               \
    } while (0)

extern "C" const char* ssw_last_error(void) { return g_err.c_str(); }
extern "C" const char* ssw_version(void) { return "ssw-b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct DevPlan {
    DctPlanHost host;
    DctPlanDev dev;
    void* tables = nullptr;
};

struct ssw_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // side stream: in the fused extract pipeline the derived frame's forward transform runs beside the base
    // frame's forward transform and ordering
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // copy streams + events of the pipelined host-buffer batch entry points
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    std::vector<cudaEvent_t> pipe_events;
    bool overlap_topk = true;              // SSW_OVERLAP_TOPK=0 keeps everything on one stream
    std::map<int, std::unique_ptr<DevPlan>> plans;
    std::map<const void*, int> smem_attr;  // kernel -> configured dynamic smem
    std::map<int, void*> fast_tw;          // line length -> stage twiddles of the compile-time plan
    cudaMemPool_t pool = nullptr;          // private stream-ordered memory pool (cudaMallocFromPoolAsync)
    std::atomic<int> refs{1};              // the owner's reference + one per live writer / reader / bank / sharded object
    long long* trace = nullptr;            // ssw_ctx_set_trace: device buffer of pipeline time stamps (dct_pipe.cuh), 16 launches x 1024 CTAs x 64
    unsigned trace_launch = 0;
    std::map<unsigned, float*> lr_tab;     // line length -> cosine factors of the low-rank embed inverse (lowrank.cuh)
    bool use_fast = true;                  // SSW_NO_FAST=1 forces the generic line kernels
    bool pdl = true;                       // SSW_PDL=0: no programmatic dependent launches
    int pdl_mode = 1;                      // SSW_PDL_MODE: 0 dependents released at kernel start everywhere; 1 (default) line kernels
                                           // before their output phase, short kernels at exit; 2 line kernels late, short kernels at start
                                           // (C2 step, 4 fresh processes each: 270.2 / 265.8 / 270.3 us)
    int col_variant = 0;                   // SSW_COL_VARIANT (tuning builds, -DSSW_TUNE)
    int row_variant = 0;                   // SSW_ROW_VARIANT (tuning builds)
    bool prefetch = false;                 // SSW_PREFETCH=1: cp.async-staged forward row pass (RowFwdPF)
    int pf_tiles = 0;                      // SSW_PF_TILES: tiles per CTA of the prefetching kernels (0 = automatic)
    bool topk_full_hist = false;           // fused pipelines: threshold bin from the whole plane (repair mode)
    bool force_line1 = false;              // SSW_FORCE_LINE1=1: single-line kernels wherever they have a plan
    int row_pipe = 1;                      // SSW_ROW_PIPE: 0 RowFwd / RowInv (one CTA per tile); 1 persistent bulk-copy pipelines (dct_pipe.cuh)
    int collect_occ = 0;                   // resident CTAs per SM of topk_collect (queried once)
    struct { const unsigned* img = nullptr; int shift = 0; float gain = 1.f; bool used = false; } row_cut;   // see launch_row_pipe
    // per-tile coefficient maxima of the forward column pipeline (PipeArgs::tile_max) for the tile-wise candidate scan:
    // want = the next forward column pass should produce them; tiles / cols = what the last one produced (0 = nothing)
    struct { int on = 1; bool want = false; unsigned* buf = nullptr; size_t cap = 0; unsigned tiles = 0, cols = 0; const float* plane = nullptr; } tile_max;   // SSW_TILE_MAX=0 disables
    int hist_relief = 1;                   // SSW_HIST_RELIEF=0: the CTAs of the forward column pipeline that build the histogram keep all their tiles
    int partial_inv = 1;                   // SSW_PARTIAL_INV=0: the fused embed sends every column back through the inverse column pass
    int row_inplace = 0;                   // SSW_ROW_INPLACE=1: inverse row pipeline with the in-place pre pass (RowPipeCfg::InvP: a third CTA per SM for 3840 / 1920-point rows)
    int col_pipe = 1;                      // SSW_COL_PIPE: 0 ColPass (one CTA per tile); 1..3 persistent TMA pipelines (dct_pipe.cuh):
                                           // 1 = 8 columns, 4 teams; 2 = 8 columns, 2 teams x 2 rounds; 3 = 4 columns, 2 teams (2 CTAs / SM)
    void* encode_tiled = nullptr;          // cuTensorMapEncodeTiled (driver entry point, resolved once)
    struct MapKey { const void* p; int w, h, batch, g; bool operator<(const MapKey& o) const {
        return std::tie(p, w, h, batch, g) < std::tie(o.p, o.w, o.h, o.batch, o.g); } };
    std::map<MapKey, std::pair<fast::TmaMap, fast::TmaMap>> tma_maps;   // plane -> (sample-side 4-D map, coefficient-side 3-D map)
    struct { bool active = false; int seg_shift = -1, chunk_shift = 0, ranks = 1, lines = 0; } seg;  // ssw_lines_forward_seg_dev
    // fused pipelines: ask the forward column pipeline for the low-frequency-block histogram of the ordering that
    // follows (want), learn whether a pipeline produced it (done) -- see run_topk_fast
    struct { bool want = false, done = false, collected = false; unsigned k = 0; int ordering = 0; } col_hist;
    bool col_split = true;                 // SSW_COL_SPLIT=0: no half tiles in the column pipelines (PipeArgs::half_tiles)
    bool col_collect = false;              // SSW_COL_COLLECT=1: the forward column pipeline appends the candidates of the ordering from its
                                           // tiles (PipeArgs::collect) instead of the topk_collect kernel.  Measured on B200 (C2): the kernel
                                           // it removes costs ~6 us inside the programmatic launch chain, the wait for the selection bin and
                                           // the extra pass over each tile cost the column pipeline ~7 us: 0.214 -> 0.222 ms/step.  Opt-in.
    bool lowrank = false;                  // SSW_LOWRANK=1: fused embed adds the low-rank update of the k modified coefficients to the
                                           // original frame (lowrank.cuh) instead of inverting the whole plane.  Measured slower on
                                           // B200 (C2: 123 vs 68 us, profiles/r2_lowrank_tensor_core.md), so it is opt-in.
    bool lowrank_mma = false;              // SSW_LOWRANK_MMA=1: the update product of the low-rank inverse on the tensor cores (3xTF32 mma.sync)
    bool col_hist_on = true;               // SSW_COL_HIST=0: selection bin from the topk_block_bin kernel even where a column pipeline runs
    bool sim_exact = false;                // SSW_SIM_EXACT=1: scores in the reference's sequential order (bit-identical)
    // asynchronous host-buffer entry points: two persistent sets of device staging buffers used alternately (the copies
    // of call i+1 run beside the kernels and the download of call i), events that mark a set free again, and the host
    // ranges with copies still in flight (a later copy that touches one of them is ordered behind it)
    struct Staging { uint8_t* in = nullptr; size_t in_cap = 0; uint8_t* out = nullptr; size_t out_cap = 0;
                     float* f32 = nullptr; size_t f32_cap = 0; cudaEvent_t done = nullptr; bool used = false; };
    Staging stage[4];
    unsigned stage_next = 0;
    cudaStream_t copy_in2 = nullptr;       // second upload stream: consecutive asynchronous calls alternate, so an upload
    unsigned async_calls = 0;              // that waits for a download (derived frames) does not hold up the next call's
    std::vector<cudaEvent_t> markers;      // ssw_ctx_marker / ssw_ctx_wait_marker: ring of events on the copy-out stream
    uint64_t marker_next = 0;
    struct HostRange { const char* p; size_t n; cudaEvent_t ev; };
    std::vector<HostRange> pending_d2h, pending_h2d;
    std::vector<cudaEvent_t> range_events;
    size_t range_next = 0;
    TopkScratch ts{};
    unsigned ts_batch = 0;
    GeneralSelect general;
    uint64_t launches = 0;
    int row_pairs = 0, col_pairs = 0;
    int sm_count = 148;
    int max_smem = 227 * 1024;
    unsigned* h_flag = nullptr;  // pinned
    int last_fallbacks = 0;
    // planes of one fused sub-batch.  Sub-batches that stay L2-resident (96 MB) were the better choice before the
    // programmatic launches; now larger launches win (C3, 64 x 1080p: 96 MB 34.7, 320 MB 38.0, 640 MB 38.8, 1100 MB 39.4 Gpix/s)
    size_t chunk_bytes = (size_t)2 << 30;
    // per-kernel CUDA-event profiling (bench.py roofline attribution); off by default
    bool profiling = false;
    struct ProfRec { const char* name; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
};

// Launch of a kernel that starts with pdl_enter() (pdl.cuh): with programmatic stream serialization the grid may
// become resident while its predecessor on the stream drains; it touches no memory before its griddepcontrol.wait.
// ONLY kernels that begin with pdl_enter() may be launched through this helper.  SSW_PDL=0 -> plain launches.
template <class... KArgs, class... Args>
static void launch_pdl(ssw_ctx* c, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = c->pdl ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);   // errors: cudaGetLastError at the call site
}

// Counts a kernel launch and, when profiling is on, brackets it with CUDA events on the context's
// stream (the stream the kernel is launched on).
struct KScope {
    ssw_ctx* c;
    cudaEvent_t b = nullptr;
    KScope(ssw_ctx* ctx, const char* name, int n_launch = 1) : c(ctx) {
        c->launches += (uint64_t)n_launch;
        if (!c->profiling) return;
        while (c->ev_pool.size() < c->ev_used + 2) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return;
            c->ev_pool.push_back(e);
        }
        cudaEvent_t a = c->ev_pool[c->ev_used++];
        b = c->ev_pool[c->ev_used++];
        cudaEventRecord(a, c->stream);
        c->prof.push_back({name, a, b});
    }
    ~KScope() { if (b) cudaEventRecord(b, c->stream); }
};

static int ctx_bind(ssw_ctx* c) {
    CK(cudaSetDevice(c->device));
    return SSW_OK;
}

extern "C" int ssw_ctx_create_on_stream(int device, void* stream, ssw_ctx** out) {
    if (!out) return fail(SSW_ERR_INVALID, "out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SSW_ERR_CUDA, std::string("no CUDA device available (libssw has no CPU path): ") +
                                      cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(SSW_ERR_INVALID, "device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(SSW_ERR_CUDA, "libssw is built for sm_100a (B200) only; found compute capability " +
                                      std::to_string(prop.major) + "." + std::to_string(prop.minor));
    auto c = std::make_unique<ssw_ctx>();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->max_smem = (int)prop.sharedMemPerBlockOptin;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    {   // stream-ordered allocations come from a pool the context owns (it keeps what it has freed: sub-batches of up to
        // 2 GiB are reused call after call) -- the device's default pool, which other libraries of the process share, is left alone
        cudaMemPoolProps props;
        std::memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        CK(cudaMemPoolCreate(&c->pool, &props));
        uint64_t thr = UINT64_MAX;
        CK(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr));
    }
    CK(cudaHostAlloc((void**)&c->h_flag, 64, cudaHostAllocDefault));
    CK(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->copy_in2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    if (const char* s = getenv("SSW_OVERLAP_TOPK")) c->overlap_topk = atoi(s) != 0;
    if (const char* s = getenv("SSW_ROW_PAIRS")) c->row_pairs = atoi(s);
    if (const char* s = getenv("SSW_COL_PAIRS")) c->col_pairs = atoi(s);
    if (const char* s = getenv("SSW_CHUNK_MB")) c->chunk_bytes = (size_t)atoll(s) << 20;
    if (const char* s = getenv("SSW_NO_FAST")) c->use_fast = atoi(s) == 0;
    if (const char* s = getenv("SSW_PDL")) c->pdl = atoi(s) != 0;
    if (const char* s = getenv("SSW_PDL_MODE")) c->pdl_mode = atoi(s);
    {
        const int small_early = (c->pdl_mode == 1) ? 0 : 1;
        CK(cudaMemcpyToSymbol(g_pdl_small_early, &small_early, sizeof(int)));
    }
    if (const char* s = getenv("SSW_COL_VARIANT")) c->col_variant = atoi(s);
    if (const char* s = getenv("SSW_ROW_VARIANT")) c->row_variant = atoi(s);
    if (const char* s = getenv("SSW_PREFETCH")) c->prefetch = atoi(s) != 0;
    if (const char* s = getenv("SSW_PF_TILES")) c->pf_tiles = atoi(s);
    if (const char* s = getenv("SSW_TOPK_FULL_HIST")) c->topk_full_hist = atoi(s) != 0;
    if (const char* s = getenv("SSW_FORCE_LINE1")) c->force_line1 = atoi(s) != 0;
    if (const char* s = getenv("SSW_COL_PIPE")) c->col_pipe = atoi(s);
    if (const char* s = getenv("SSW_ROW_PIPE")) c->row_pipe = atoi(s);
    if (const char* s = getenv("SSW_ROW_INPLACE")) c->row_inplace = atoi(s);
    if (const char* s = getenv("SSW_PARTIAL_INV")) c->partial_inv = atoi(s);
    if (const char* s = getenv("SSW_HIST_RELIEF")) c->hist_relief = atoi(s);
    if (const char* s = getenv("SSW_TILE_MAX")) c->tile_max.on = atoi(s);
    if (const char* s = getenv("SSW_SIM_EXACT")) c->sim_exact = atoi(s) != 0;
    if (const char* s = getenv("SSW_COL_HIST")) c->col_hist_on = atoi(s) != 0;
    if (const char* s = getenv("SSW_COL_SPLIT")) c->col_split = atoi(s) != 0;
    if (const char* s = getenv("SSW_COL_COLLECT")) c->col_collect = atoi(s) != 0;
    if (const char* s = getenv("SSW_LOWRANK")) c->lowrank = atoi(s) != 0;
    if (const char* s = getenv("SSW_LOWRANK_MMA")) c->lowrank_mma = atoi(s) != 0;
    *out = c.release();
    return SSW_OK;
}

extern "C" int ssw_ctx_create(int device, ssw_ctx** out) { return ssw_ctx_create_on_stream(device, nullptr, out); }

static void topk_scratch_free(ssw_ctx* c) {
    if (!c->ts_batch) return;
    cudaFree(c->ts.hist); cudaFree(c->ts.ticket); cudaFree(c->ts.sel_bin);
    cudaFree(c->tile_max.buf); c->tile_max.buf = nullptr; c->tile_max.cap = 0;
    cudaFree(c->ts.cand_count); cudaFree(c->ts.cand); cudaFree(c->ts.overflow); cudaFree(c->ts.maxrow); cudaFree(c->ts.maxcol);
    c->ts = TopkScratch{};
    c->ts_batch = 0;
}

// Objects created from a context (writers, readers, banks, sharded frames) keep it alive: ssw_ctx_destroy drops the
// owner's reference, the context is torn down when the last object has gone.  Host languages whose finalisers run in
// no particular order (Python at interpreter exit, a Rust thread-local against values that outlive it) stay safe.
static void ctx_free(ssw_ctx* c) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& kv : c->plans) cudaFree(kv.second->tables);
    for (auto& kv : c->fast_tw) cudaFree(kv.second);
    for (auto& kv : c->lr_tab) cudaFree(kv.second);
    topk_scratch_free(c);
    c->general.release();
    if (c->h_flag) cudaFreeHost(c->h_flag);
    if (c->aux) { cudaStreamSynchronize(c->aux); cudaStreamDestroy(c->aux); }
    if (c->copy_in) { cudaStreamSynchronize(c->copy_in); cudaStreamDestroy(c->copy_in); }
    if (c->copy_out) { cudaStreamSynchronize(c->copy_out); cudaStreamDestroy(c->copy_out); }
    if (c->copy_in2) { cudaStreamSynchronize(c->copy_in2); cudaStreamDestroy(c->copy_in2); }
    for (cudaEvent_t e : c->markers) cudaEventDestroy(e);
    for (cudaEvent_t e : c->pipe_events) cudaEventDestroy(e);
    for (cudaEvent_t e : c->range_events) cudaEventDestroy(e);
    for (auto& st : c->stage) {
        cudaFree(st.in); cudaFree(st.out); cudaFree(st.f32);
        if (st.done) cudaEventDestroy(st.done);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    if (c->pool) { cudaDeviceSynchronize(); cudaMemPoolDestroy(c->pool); }   // frees of the other streams have run
    delete c;
}
static void ctx_release(ssw_ctx* c) { if (c && c->refs.fetch_sub(1) == 1) ctx_free(c); }
struct CtxRef {
    ssw_ctx* c = nullptr;
    CtxRef() = default;
    CtxRef(const CtxRef&) = delete;
    CtxRef& operator=(const CtxRef&) = delete;
    CtxRef& operator=(ssw_ctx* ctx) { if (ctx) ctx->refs.fetch_add(1); ctx_release(c); c = ctx; return *this; }
    ~CtxRef() { ctx_release(c); }
    ssw_ctx* operator->() const { return c; }
    operator ssw_ctx*() const { return c; }
};

extern "C" int ssw_ctx_destroy(ssw_ctx* c) {
    if (!c) return SSW_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);   // the caller's work is complete when the call returns, whoever frees the context
    ctx_release(c);
    return SSW_OK;
}

// diagnostics: per-CTA time line of the persistent pipelines (dct_pipe.cuh, trace_tile).  dev_buf: 16 x 1024 x 64 int64
// (8 MiB) of device memory, or NULL to switch it off; launch i of a pipeline kernel writes block (i % 16).
extern "C" int ssw_ctx_set_trace(ssw_ctx* c, void* dev_buf) {
    if (!c) return fail(SSW_ERR_INVALID, "ctx is NULL");
#if !defined(SSW_TRACE)
    if (dev_buf) return fail(SSW_ERR_UNSUPPORTED, "this libssw was built without -DSSW_TRACE (tools/build_tmp.sh trace -DSSW_TRACE)");
#endif
    c->trace = (long long*)dev_buf;
    c->trace_launch = 0;
    return SSW_OK;
}

extern "C" int ssw_ctx_synchronize(ssw_ctx* c) {
    if (!c) return fail(SSW_ERR_INVALID, "ctx is NULL");
    CKS(ctx_bind(c));
    CK(cudaStreamSynchronize(c->stream));
    // the asynchronous host-buffer entry points finish on the copy streams
    if (c->copy_in) CK(cudaStreamSynchronize(c->copy_in));
    if (c->copy_in2) CK(cudaStreamSynchronize(c->copy_in2));
    if (c->copy_out) CK(cudaStreamSynchronize(c->copy_out));
    if (c->aux) CK(cudaStreamSynchronize(c->aux));
    c->pending_d2h.clear();
    c->pending_h2d.clear();
    return SSW_OK;
}
extern "C" void* ssw_ctx_stream(ssw_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" uint64_t ssw_ctx_launch_count(ssw_ctx* c) { return c ? c->launches : 0; }
extern "C" int ssw_ctx_set_tiling(ssw_ctx* c, int row_pairs, int col_pairs) {
    if (!c || row_pairs < 0 || col_pairs < 0) return fail(SSW_ERR_INVALID, "bad tiling");
    c->row_pairs = row_pairs;
    c->col_pairs = col_pairs;
    return SSW_OK;
}

extern "C" int ssw_ctx_profile_begin(ssw_ctx* c) {
    if (!c) return fail(SSW_ERR_INVALID, "ctx is NULL");
    CKS(ctx_bind(c));
    CK(cudaStreamSynchronize(c->stream));
    c->prof.clear();
    c->ev_used = 0;
    c->profiling = true;
    return SSW_OK;
}

extern "C" int ssw_ctx_profile_end(ssw_ctx* c, char* json_out, size_t cap) {
    if (!c || !json_out || cap < 3) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(ctx_bind(c));
    c->profiling = false;
    CK(cudaStreamSynchronize(c->stream));
    std::map<std::string, std::pair<uint64_t, double>> agg;  // name -> (launches, ms)
    for (const auto& r : c->prof) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, r.a, r.b));
        auto& e = agg[r.name];
        e.first += 1;
        e.second += (double)ms;
    }
    c->prof.clear();
    c->ev_used = 0;
    std::string js = "{";
    bool first = true;
    for (const auto& kv : agg) {
        char tmp[256];
        snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"launches\": %llu, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(),
                 (unsigned long long)kv.second.first, kv.second.second);
        js += tmp;
        first = false;
    }
    js += "}";
    if (js.size() + 1 > cap) return fail(SSW_ERR_INVALID, "profile buffer too small");
    std::memcpy(json_out, js.c_str(), js.size() + 1);
    return SSW_OK;
}

extern "C" int ssw_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(SSW_ERR_INVALID, "out is NULL");
    CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return SSW_OK;
}
extern "C" int ssw_host_free(void* p) {
    if (p) CK(cudaFreeHost(p));
    return SSW_OK;
}

// ------------------------------------------------------------------------------------------------
// plans
// ------------------------------------------------------------------------------------------------
static int get_plan(ssw_ctx* c, int n, const DevPlan** out) {
    auto it = c->plans.find(n);
    if (it != c->plans.end()) { *out = it->second.get(); return SSW_OK; }
    auto p = std::make_unique<DevPlan>();
    p->host = make_dct_plan(n);
    if (!p->host.error.empty()) return fail(SSW_ERR_UNSUPPORTED, p->host.error);
    const DctPlanHost& h = p->host;
    if ((size_t)h.npad * sizeof(cplx) > (size_t)c->max_smem)
        return fail(SSW_ERR_UNSUPPORTED, "line length " + std::to_string(n) + " does not fit shared memory");
    const size_t b_tw = h.stage_tw.size() * sizeof(float), b_wn = h.wn.size() * sizeof(float),
                 b_t4 = h.t4.size() * sizeof(float);
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    CK(cudaMalloc(&p->tables, al(b_tw) + al(b_wn) + al(b_t4) + 256));
    char* base = (char*)p->tables;
    if (b_tw) CK(cudaMemcpy(base, h.stage_tw.data(), b_tw, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(base + al(b_tw), h.wn.data(), b_wn, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(base + al(b_tw) + al(b_wn), h.t4.data(), b_t4, cudaMemcpyHostToDevice));
    DctPlanDev& d = p->dev;
    std::memset(&d, 0, sizeof(d));
    d.n = h.n; d.npad = h.npad; d.tp = h.tp; d.nstages = h.nstages;
    for (int i = 0; i < h.nstages; ++i) {
        d.stages[i] = h.stages[i];
        d.ns_magic[i] = h.stages[i].ns > 1 ? (unsigned)((1ull << 32) / (unsigned)h.stages[i].ns + 1) : 0u;
    }
    d.stage_tw = (const cplx*)base;
    d.wn = (const cplx*)(base + al(b_tw));
    d.t4 = (const cplx*)(base + al(b_tw) + al(b_wn));
    *out = p.get();
    c->plans[n] = std::move(p);
    return SSW_OK;
}

// ------------------------------------------------------------------------------------------------
// line-pass launches
// ------------------------------------------------------------------------------------------------
struct Tiling { int P, G, threads; size_t smem; };

static int pick_tiling(ssw_ctx* c, const DctPlanDev& pl, bool column, int lines, Tiling* t) {
    const size_t pair_bytes = (size_t)pl.npad * sizeof(cplx);
    int G = std::max(1, 256 / pl.tp);
    int P = column ? (c->col_pairs ? c->col_pairs : 4) : (c->row_pairs ? c->row_pairs : G);
    const int max_pairs = (lines + 1) / 2;
    P = std::max(1, std::min(P, max_pairs));
    while (P > 1 && (size_t)P * pair_bytes > (size_t)c->max_smem) --P;
    if ((size_t)P * pair_bytes > (size_t)c->max_smem) return fail(SSW_ERR_UNSUPPORTED, "line does not fit shared memory");
    G = std::max(1, std::min(G, P));
    t->P = P; t->G = G; t->threads = G * pl.tp; t->smem = (size_t)P * pair_bytes;
    return SSW_OK;
}

template <class K>
static int launch_line(ssw_ctx* c, const char* name, K kernel, const LineArgs& a, const Tiling& t, long long tiles) {
    const void* key = (const void*)kernel;
    auto it = c->smem_attr.find(key);
    if (it == c->smem_attr.end() || it->second < (int)t.smem) {
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem));
        c->smem_attr[key] = (int)t.smem;
    }
    if (tiles <= 0 || tiles > 0x7FFFFFFFll) return fail(SSW_ERR_INVALID, "tile count out of range");
    {
        KScope ks(c, name);
        kernel<<<(unsigned)tiles, t.threads, t.smem, c->stream>>>(a);
    }
    CK(cudaGetLastError());
    return SSW_OK;
}

static LineArgs base_args(const DctPlanDev& pl, int w, int h, const Tiling& t, long long npix) {
    LineArgs a;
    std::memset(&a, 0, sizeof(a));
    a.plan = pl; a.w = w; a.h = h; a.P = t.P;
    a.scale0 = 1.f; a.scalen = 1.f;
    a.src_stride = npix; a.plane_stride = npix; a.dst_stride = npix;
    return a;
}


// ------------------------------------------------------------------------------------------------
// fast path: compile-time planned kernels (dct_fast.cuh) for the common frame sizes
// ------------------------------------------------------------------------------------------------
template <class P>
static int fast_tables(ssw_ctx* c, const cplx** tw, const cplx** t4) {
    const DevPlan* gp;
    CKS(get_plan(c, P::N, &gp));  // the generic plan owns the exp(-i*pi*k/2N) table
    *t4 = gp->dev.t4;
    auto it = c->fast_tw.find(P::KEY);
    if (it == c->fast_tw.end()) {
        std::vector<float> h(2 * (size_t)P::TW_TOTAL + 2);
        fast::make_stage_twiddles<P>(h.data());
        void* d = nullptr;
        CK(cudaMalloc(&d, h.size() * sizeof(float)));
        CK(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
        it = c->fast_tw.emplace(P::KEY, d).first;
    }
    *tw = (const cplx*)it->second;
    return SSW_OK;
}

// single-line kernels: stage twiddles of the M-point plan + exp(-i*pi*k/2N) for the line length N = 2M
template <class P>
static int line1_tables(ssw_ctx* c, const cplx** tw, const cplx** t4) {
    const int key = -(int)P::N;  // negative keys: line1 tables (pair plans use +N)
    auto it = c->fast_tw.find(key);
    if (it == c->fast_tw.end()) {
        const int n = 2 * P::N;
        std::vector<float> h(2 * (size_t)P::TW_TOTAL + 2 + 2 * (size_t)n);
        fast::make_stage_twiddles<P>(h.data());
        float* t = h.data() + 2 * (size_t)P::TW_TOTAL + 2;
        for (int j = 0; j < n; ++j) {
            const double b = -M_PI * (double)j / (2.0 * (double)n);
            t[2 * j] = (float)std::cos(b);
            t[2 * j + 1] = (float)std::sin(b);
        }
        void* d = nullptr;
        CK(cudaMalloc(&d, h.size() * sizeof(float)));
        CK(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
        it = c->fast_tw.emplace(key, d).first;
    }
    *tw = (const cplx*)it->second;
    *t4 = (const cplx*)((const float*)it->second + 2 * (size_t)P::TW_TOTAL + 2);
    return SSW_OK;
}

template <class K, bool LINE1 = false>
static int launch_fast(ssw_ctx* c, const char* name, fast::FastArgs a, int w, int h, int batch) {
    if (LINE1) CKS(line1_tables<typename K::P>(c, &a.tw, &a.t4));
    else CKS(fast_tables<typename K::P>(c, &a.tw, &a.t4));
    a.tiles_per_image = K::tiles_per_image(w, h);
    a.pdl_late = c->pdl_mode != 0;
    const long long tiles = (long long)a.tiles_per_image * batch;
    if (tiles <= 0 || tiles > 0x7FFFFFFFll) return fail(SSW_ERR_INVALID, "tile count out of range");
    auto kernel = fast::fast_kernel<K>;
    const void* key = (const void*)kernel;
    if (c->smem_attr.find(key) == c->smem_attr.end()) {
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
        c->smem_attr[key] = K::SMEM;
    }
    {
        KScope ks(c, name);
        launch_pdl(c, kernel, (unsigned)tiles, K::THREADS, K::SMEM, c->stream, a);
    }
    CK(cudaGetLastError());
    return SSW_OK;
}

// prefetching kernels (fast_kernel_pf): a CTA works through `per` consecutive tiles
template <class K>
static int launch_fast_pf(ssw_ctx* c, const char* name, fast::FastArgs a, int w, int h, int batch) {
    CKS(fast_tables<typename K::P>(c, &a.tw, &a.t4));
    a.tiles_per_image = K::tiles_per_image(w, h);
    a.pdl_late = c->pdl_mode != 0;
    const long long tiles = (long long)a.tiles_per_image * batch;
    if (tiles <= 0 || tiles > 0x7FFFFFFFll) return fail(SSW_ERR_INVALID, "tile count out of range");
    auto kernel = fast::fast_kernel_pf<K>;
    const void* key = (const void*)kernel;
    auto it = c->smem_attr.find(key);
    if (it == c->smem_attr.end()) {
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, K::THREADS, K::SMEM));
        it = c->smem_attr.emplace(key, std::max(1, occ)).first;   // value: resident CTAs per SM
    }
    // enough tiles per CTA that the grid is about one resident wave, at most 8
    const long long slots = (long long)it->second * c->sm_count;
    long long per = c->pf_tiles > 0 ? c->pf_tiles : std::min<long long>(8, (tiles + slots - 1) / slots);
    per = std::max<long long>(1, per);
    a.total_tiles = (int)tiles;
    a.tiles_per_cta = (int)per;
    {
        KScope ks(c, name);
        launch_pdl(c, kernel, (unsigned)((tiles + per - 1) / per), K::THREADS, K::SMEM, c->stream, a);
    }
    CK(cudaGetLastError());
    return SSW_OK;
}

static fast::FastArgs fast_args(int w, int h) {
    fast::FastArgs a;
    std::memset(&a, 0, sizeof(a));
    a.w = w; a.h = h;
    a.scale0 = 1.f; a.scalen = 1.f;
    a.src_stride = a.plane_stride = a.dst_stride = (long long)w * h;
    a.seg_shift = -1;
    a.dbg_skip = 0;
    a.neg_zero = -0.0f;
    return a;
}

static void apply_seg(const ssw_ctx* c, fast::FastArgs* a) {
    if (!c->seg.active) return;
    a->seg_shift = c->seg.seg_shift; a->chunk_shift = c->seg.chunk_shift;
    a->seg_ranks = c->seg.ranks; a->seg_lines = c->seg.lines;
}

static bool aligned(const void* p, size_t n) { return (((size_t)p) & (n - 1)) == 0; }

#ifdef SSW_TUNE
template <class K, int M> struct WithMinB : K { static constexpr int MINB = M; };
#endif

// ---- persistent bulk-copy row pipelines (dct_pipe.cuh): RGB8 frames <-> coefficient planes -------------------------
template <class K>
static int launch_row_pipe(ssw_ctx* c, const char* name, const void* pix, float* plane, void* out, int w, int h, int batch,
                           float scale0, float scalen) {
    using P = typename K::P;
    fast::RowPipeArgs a;
    std::memset(&a, 0, sizeof(a));
    a.w = w; a.h = h; a.batch = batch; a.scale0 = scale0; a.scalen = scalen;
    a.pix = (const unsigned char*)pix; a.plane = plane; a.out = (unsigned char*)out;
    CKS(fast_tables<P>(c, &a.tw, &a.t4));
    a.tiles_per_image = K::tiles_per_image(w, h);
    const long long tiles = (long long)a.tiles_per_image * batch;
    if (tiles <= 0 || tiles > 0x7FFFFFFFll) return fail(SSW_ERR_INVALID, "tile count out of range");
    a.total_tiles = (int)tiles;
    a.pdl_late = c->pdl_mode != 0;
    a.neg_zero = -0.0f;
    a.trace = c->trace ? c->trace + (size_t)(c->trace_launch++ % 16u) * 1024 * 64 : nullptr;
    if (K::INVERSE && c->row_cut.img) {   // rows after a partial inverse column pass (ssw_embed_batch_rgb8_dev)
        a.col_cut_img = c->row_cut.img; a.col_tile_shift = c->row_cut.shift; a.col_gain = c->row_cut.gain;
        c->row_cut.used = true;
    }
    auto kernel = fast::row_pipe_kernel<K>;
    const void* key = (const void*)kernel;
    auto it = c->smem_attr.find(key);
    if (it == c->smem_attr.end()) {
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, K::THREADS, K::SMEM));
        it = c->smem_attr.emplace(key, std::max(1, occ)).first;   // value: resident CTAs per SM
    }
    const long long slots = (long long)it->second * c->sm_count;
    const unsigned grid = (unsigned)std::min<long long>(tiles, slots);
    {
        KScope ks(c, name);
        launch_pdl(c, kernel, grid, K::THREADS, K::SMEM, c->stream, a);
    }
    CK(cudaGetLastError());
    return SSW_OK;
}

// *done = true when a pipeline ran.  inverse: d_pix = the original frames, d_out = the destination frames.
static int pipe_row(ssw_ctx* c, bool inverse, const void* d_pix, float* d_plane, void* d_out, int w, int h, int batch,
                    float scale0, float scalen, bool* done) {
    *done = false;
    if (c->row_pipe <= 0 || c->seg.active) return SSW_OK;
    // bulk copies need 16-byte aligned tiles on both sides: aligned bases and whole 16-byte frames
    if (!aligned(d_pix, 16) || !aligned(d_plane, 16) || (inverse && !aligned(d_out, 16)) || (((size_t)w * h * 3) & 15)) return SSW_OK;
    int rc = SSW_OK;
    fast::with_plan(w, [&](auto p) {
        using P = decltype(p);
        using Cfg = fast::RowPipeCfg<P>;
        if constexpr (Cfg::OK) {
            if (!Cfg::Fwd::supports(w, h)) return;
            if (inverse && c->row_inplace && Cfg::INPLACE_OK) rc = launch_row_pipe<typename Cfg::InvP>(c, "inv_rows", d_pix, d_plane, d_out, w, h, batch, scale0, scalen);
            else if (inverse) rc = launch_row_pipe<typename Cfg::Inv>(c, "inv_rows", d_pix, d_plane, d_out, w, h, batch, scale0, scalen);
            else rc = launch_row_pipe<typename Cfg::Fwd>(c, "fwd_rows", d_pix, d_plane, nullptr, w, h, batch, scale0, scalen);
            *done = true;
        }
    });
    return rc;
}

// would pipe_row(inverse) run for these frames?  (the partial inverse needs it: only the row pipeline applies RowPipeArgs::col_cut_img)
static bool pipe_row_available(const ssw_ctx* c, const void* d_pix, const void* d_out, int w, int h) {
    if (c->row_pipe <= 0 || c->seg.active || !c->use_fast) return false;
    if (!aligned(d_pix, 16) || !aligned(d_out, 16) || (((size_t)w * h * 3) & 15)) return false;
    bool ok = false;
    fast::with_plan(w, [&](auto p) {
        using Cfg = fast::RowPipeCfg<decltype(p)>;
        if constexpr (Cfg::OK) ok = Cfg::Fwd::supports(w, h);
    });
    return ok;
}

// returns SSW_OK and sets *done when a fast kernel ran; *done = false -> caller uses the generic kernel
static int fast_row_fwd(ssw_ctx* c, int src_type, const void* d_src, int w, int h, int batch, float* d_plane,
                        float scale0, float scalen, bool* done) {
    *done = false;
    if (!c->use_fast) return SSW_OK;
    if (!aligned(d_plane, 16) || !aligned(d_src, src_type == PIX_RGB8 ? 4 : 16)) return SSW_OK;
    int rc = SSW_OK;
    if (src_type == PIX_RGB8 && !c->prefetch && !c->row_variant) {
        CKS(pipe_row(c, false, d_src, d_plane, nullptr, w, h, batch, scale0, scalen, done));
        if (*done) return SSW_OK;
    }
#ifdef SSW_TUNE
    if (w == 3840 && src_type == PIX_RGB8 && c->row_variant) {
        fast::FastArgs a = fast_args(w, h);
        a.src = d_src; a.plane = d_plane; a.scale0 = scale0; a.scalen = scalen;
        *done = true;
        if (c->row_variant >= 10) {   // 10 + mask: default kernel with parts switched off (upper bounds for pipelining)
            a.dbg_skip = c->row_variant - 10;
            return launch_fast<fast::RowFwd<fast::Plan3840, 1, PIX_RGB8>>(c, "fwd_rows", a, w, h, batch);
        }
        switch (c->row_variant) {
            case 1: return launch_fast<fast::RowFwd<fast::Plan3840b, 1, PIX_RGB8>>(c, "fwd_rows", a, w, h, batch);
            case 2: return launch_fast<fast::RowFwd<fast::Plan3840, 2, PIX_RGB8>>(c, "fwd_rows", a, w, h, batch);
            case 3: return launch_fast<fast::RowFwd<fast::Plan3840c, 1, PIX_RGB8>>(c, "fwd_rows", a, w, h, batch);
            case 4: return launch_fast<fast::RowFwd<fast::Plan3840b, 2, PIX_RGB8>>(c, "fwd_rows", a, w, h, batch);
            case 5: return launch_fast<WithMinB<fast::RowFwd<fast::Plan3840b, 1, PIX_RGB8>, 3>>(c, "fwd_rows", a, w, h, batch);
            case 6: return launch_fast<WithMinB<fast::RowFwd<fast::Plan3840, 1, PIX_RGB8>, 5>>(c, "fwd_rows", a, w, h, batch);
            case 7: return launch_fast<WithMinB<fast::RowFwd<fast::Plan3840c, 1, PIX_RGB8>, 3>>(c, "fwd_rows", a, w, h, batch);
            case 8: return launch_fast<WithMinB<fast::RowFwd<fast::Plan3840d, 1, PIX_RGB8>, 4>>(c, "fwd_rows", a, w, h, batch);
            case 9: return launch_fast<WithMinB<fast::RowFwd<fast::Plan3840d, 2, PIX_RGB8>, 2>>(c, "fwd_rows", a, w, h, batch);
            default: *done = false; break;
        }
    }
#endif
    *done = fast::with_plan(w, [&](auto p) {
        using P = decltype(p);
        constexpr int G = fast::RowG<P>::value;
        fast::FastArgs a = fast_args(w, h);
        a.src = d_src; a.plane = d_plane; a.scale0 = scale0; a.scalen = scalen;
        apply_seg(c, &a);
        if (src_type == PIX_RGB8) {
            if constexpr ((3 * P::N) % 16 == 0) {
                if (c->prefetch && aligned(d_src, 16)) {   // cp.async-staged rows, several tiles per CTA
                    rc = launch_fast_pf<fast::RowFwdPF<P, G>>(c, "fwd_rows", a, w, h, batch);
                    return;
                }
            }
            rc = launch_fast<fast::RowFwd<P, G, PIX_RGB8>>(c, "fwd_rows", a, w, h, batch);
        } else if (src_type == PIX_RGB32F) {
            rc = launch_fast<fast::RowFwd<P, G, PIX_RGB32F>>(c, "fwd_rows_rgb32f", a, w, h, batch);
        } else {
            rc = launch_fast<fast::RowFwd<P, G, PIX_PLANE>>(c, "fwd_rows_plane", a, w, h, batch);
        }
    });
    return rc;
}

template <class P, bool INV>
static int launch_col_variant(ssw_ctx* c, const fast::FastArgs& a, int w, int h, int batch) {
    const char* name = INV ? "inv_cols" : "fwd_cols";
#ifdef SSW_TUNE
    switch (c->col_variant) {
        case 1: return launch_fast<fast::ColPass<P, 4, 4, INV>>(c, name, a, w, h, batch);
        case 2: return launch_fast<fast::ColPass<P, 4, 1, INV>>(c, name, a, w, h, batch);
        case 3: return launch_fast<fast::ColPass<P, 8, 4, INV>>(c, name, a, w, h, batch);
        case 4: return launch_fast<fast::ColPass<P, 8, 2, INV>>(c, name, a, w, h, batch);
        case 5: return launch_fast<fast::ColPass<P, 2, 2, INV>>(c, name, a, w, h, batch);
        case 6: return launch_fast<fast::ColPass<P, 4, 2, INV, 3>>(c, name, a, w, h, batch);
        case 7: return launch_fast<fast::ColPass<P, 2, 1, INV>>(c, name, a, w, h, batch);
        default: break;
    }
#endif
    constexpr int G = fast::ColG<P>::value, TEAMS = fast::ColTeams<P>::value;
    if constexpr (G == 0) {
        (void)a; (void)w; (void)h; (void)batch; (void)name;
        return SSW_ERR_UNSUPPORTED;   // tile does not fit shared memory: generic column kernel
    } else {
        return launch_fast<fast::ColPass<P, G, TEAMS, INV, ((P::PAD && TEAMS * P::T <= 256) ? 2 : 0)>>(c, name, a, w, h, batch);
    }
}

static OrderConsts make_order(int ordering, int w, int h);

// ---- persistent TMA column pipelines (dct_pipe.cuh) ---------------------------------------------------------------
// tensor maps of a coefficient plane [batch][h][w] f32 for tiles of 2*g columns:
//   sample side      4-D (column, row parity, row pair, image)  -- even rows / odd rows as separate boxes (Makhoul split)
//   coefficient side 3-D (column, row, image)
// swz: 32-byte swizzle (boxes of exactly 8 floats): the 16-byte halves of a buffer row are swapped in rows 4..7 (mod 8)
static int tma_maps_for(ssw_ctx* c, const float* plane, int w, int h, int batch, int g, int rb_half, int rb_full, bool swz,
                        const std::pair<fast::TmaMap, fast::TmaMap>** out) {
    if (swz && g != 4) return fail(SSW_ERR_INVALID, "the 32-byte swizzle needs boxes of 8 columns");
    const ssw_ctx::MapKey key{plane, w, h, batch, g + (swz ? 100 : 0)};
    const CUtensorMapSwizzle swizzle = swz ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    auto it = c->tma_maps.find(key);
    if (it != c->tma_maps.end()) { *out = &it->second; return SSW_OK; }
    if (!c->encode_tiled) {
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &c->encode_tiled, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !c->encode_tiled)
            return fail(SSW_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    }
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    EncodeFn enc = (EncodeFn)c->encode_tiled;
    static_assert(sizeof(CUtensorMap) == sizeof(fast::TmaMap), "tensor map size");
    std::pair<fast::TmaMap, fast::TmaMap> maps;
    const cuuint32_t ones[4] = {1, 1, 1, 1};
    {
        CUtensorMap m;
        const cuuint64_t dims[4] = {(cuuint64_t)w, 2, (cuuint64_t)h / 2, (cuuint64_t)batch};
        const cuuint64_t strides[3] = {(cuuint64_t)w * 4, (cuuint64_t)w * 8, (cuuint64_t)w * h * 4};
        const cuuint32_t box[4] = {(cuuint32_t)(2 * g), 1, (cuuint32_t)rb_half, 1};
        const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)plane, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SSW_ERR_CUDA, "cuTensorMapEncodeTiled (4-D sample view) failed: " + std::to_string((int)r));
        std::memcpy(&maps.first, &m, sizeof(m));
    }
    {
        CUtensorMap m;
        const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)batch};
        const cuuint64_t strides[2] = {(cuuint64_t)w * 4, (cuuint64_t)w * h * 4};
        const cuuint32_t box[3] = {(cuuint32_t)(2 * g), (cuuint32_t)rb_full, 1};
        const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)plane, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SSW_ERR_CUDA, "cuTensorMapEncodeTiled (3-D coefficient view) failed: " + std::to_string((int)r));
        std::memcpy(&maps.second, &m, sizeof(m));
    }
    *out = &c->tma_maps.emplace(key, maps).first->second;
    return SSW_OK;
}

// d_plane: the plane the pass reads; d_out: the plane it writes (nullptr = in place); col_limit: PipeArgs::col_limit (inverse)
template <class K>
static int launch_col_pipe(ssw_ctx* c, const char* name, int w, int h, int batch, float* d_plane, float scale0, float scalen,
                           float* d_out = nullptr, const unsigned* col_limit = nullptr) {
    using P = typename K::P;
    fast::PipeArgs a;
    std::memset(&a, 0, sizeof(a));
    a.w = w; a.h = h; a.batch = batch; a.scale0 = scale0; a.scalen = scalen;
    CKS(fast_tables<P>(c, &a.tw, &a.t4));
    a.tiles_per_image = K::tiles_per_image(w, h);
    const long long tiles = (long long)a.tiles_per_image * batch;
    if (tiles <= 0 || tiles > 0x7FFFFFFFll) return fail(SSW_ERR_INVALID, "tile count out of range");
    a.total_tiles = (int)tiles;
    a.pdl_late = c->pdl_mode != 0;
    a.trace = c->trace ? c->trace + (size_t)(c->trace_launch++ % 16u) * 1024 * 64 : nullptr;
    if (!K::INVERSE && c->col_hist.want && (unsigned)batch <= c->ts_batch) {
        const bool small = c->col_hist.k <= (unsigned)kSmallMaxK;
        a.ts = c->ts;
        a.hist_k = c->col_hist.k;
        a.hist_rows = std::min(h, small ? kSmallRows : kBlockRows);
        a.hist_cols = std::min(w, small ? kSmallCols : kBlockCols);
        a.oc = make_order(c->col_hist.ordering, w, h);
        c->col_hist.done = true;
    }
    if (c->tma_maps.size() > 4096) c->tma_maps.clear();   // planes come and go (cudaMallocAsync): bounded cache (cleared before any lookup of this launch)
    const std::pair<fast::TmaMap, fast::TmaMap>* maps;
    CKS(tma_maps_for(c, d_plane, w, h, batch, K::G, K::RB_HALF, K::RB_FULL, K::SWZ, &maps));
    const std::pair<fast::TmaMap, fast::TmaMap>* maps2 = maps;   // boxes of G columns for half tiles
    if (K::HALF_OK) CKS(tma_maps_for(c, d_plane, w, h, batch, K::G / 2, K::RB_HALF, K::RB_FULL, false, &maps2));
    // out of place: the side the pass stores through (forward: coefficient view, inverse: sample view) belongs to d_out
    const std::pair<fast::TmaMap, fast::TmaMap>* omaps = maps;
    const std::pair<fast::TmaMap, fast::TmaMap>* omaps2 = maps2;
    if (d_out && d_out != d_plane) {
        if (!aligned(d_out, 16)) return fail(SSW_ERR_INVALID, "column pipeline: destination plane must be 16-byte aligned");
        CKS(tma_maps_for(c, d_out, w, h, batch, K::G, K::RB_HALF, K::RB_FULL, K::SWZ, &omaps));
        omaps2 = omaps;
        if (K::HALF_OK) CKS(tma_maps_for(c, d_out, w, h, batch, K::G / 2, K::RB_HALF, K::RB_FULL, false, &omaps2));
    }
    if (!K::INVERSE && K::TRACK && c->tile_max.want && c->tile_max.on) {
        const size_t need = (size_t)tiles;
        if (need > c->tile_max.cap) {
            CK(cudaStreamSynchronize(c->stream));
            cudaFree(c->tile_max.buf); c->tile_max.buf = nullptr; c->tile_max.cap = 0;
            CK(cudaMalloc(&c->tile_max.buf, need * sizeof(unsigned)));
            CK(cudaMemset(c->tile_max.buf, 0, need * sizeof(unsigned)));
            c->tile_max.cap = need;
        }
        a.tile_max = c->tile_max.buf;
        c->tile_max.tiles = (unsigned)a.tiles_per_image; c->tile_max.cols = 2 * K::G;
        c->tile_max.plane = (d_out && d_out != d_plane) ? d_out : d_plane;   // the coefficient planes the maxima describe
    }
    a.col_limit = K::INVERSE ? col_limit : nullptr;
    a.col_limit_img = a.col_limit ? col_limit + 1 : nullptr;   // (TopkScratch::maxcol layout: [0] the launch, [1 + i] image i)
    const fast::TmaMap& m_s = K::INVERSE ? omaps->first : maps->first;      // sample side: loaded by the forward pass, stored by the inverse
    const fast::TmaMap& m_c = K::INVERSE ? maps->second : omaps->second;    // coefficient side
    const fast::TmaMap& m_s2 = K::INVERSE ? omaps2->first : maps2->first;
    const fast::TmaMap& m_c2 = K::INVERSE ? maps2->second : omaps2->second;
    auto kernel = fast::col_pipe_kernel<K>;
    const void* key = (const void*)kernel;
    auto it = c->smem_attr.find(key);
    if (it == c->smem_attr.end()) {
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, K::THREADS, K::SMEM));
        it = c->smem_attr.emplace(key, std::max(1, occ)).first;   // value: resident CTAs per SM
    }
    const long long slots = (long long)it->second * c->sm_count;
    const unsigned grid = (unsigned)std::min<long long>(tiles, slots);
    // schedule: whole rounds of whole tiles; a remainder that would keep less than half of the CTAs busy for one more
    // round is cut into half tiles, one per CTA (single frames: 480 tiles on 148 CTAs = 3 rounds + 72 half tiles)
    a.full_tiles = a.total_tiles; a.half_tiles = 0;
    const long long rounds = tiles / grid, rem = tiles % grid;
    if (K::HALF_OK && c->col_split && batch == 1 && rounds >= 1 && rem > 0 && 2 * rem <= (long long)grid && (w % (2 * K::G)) == 0) {
        a.full_tiles = (int)(rounds * grid);
        a.half_tiles = (int)(2 * rem);
    }
    if (K::HALF_OK && !K::INVERSE && a.ts.hist && batch == 1 && a.half_tiles > 0 && c->hist_relief) {
        const long long nh = (a.hist_cols + 2 * K::G - 1) / (2 * K::G);
        if (rounds >= 2 && 3 * nh <= (long long)grid - a.half_tiles) a.hist_relief = (int)nh;
    }
    if (!K::INVERSE && a.ts.hist && batch == 1 && a.half_tiles == 0) {
        // one frame: `rot` CTAs get one tile more than the others and finish last -- keep the histogram tiles off them
        const long long nh = (a.hist_cols + 2 * K::G - 1) / (2 * K::G);
        if (rem + nh <= grid) a.tile_rot = (int)rem;
    }
    // the candidates of the ordering straight from the tiles (Energy ordering; see PipeArgs::collect)
    if (K::COLLECT_OK && !K::INVERSE && a.ts.hist && c->col_collect && !c->lowrank && c->col_hist.ordering == 0 && batch == 1) { a.collect = 1; c->col_hist.collected = true; }
    a.tab_bulk = aligned(a.tw, 16) && aligned(a.t4, 16);
    {
        KScope ks(c, name);
        launch_pdl(c, kernel, grid, K::THREADS, K::SMEM, c->stream, a, m_s, m_c, m_s2, m_c2);
    }
    CK(cudaGetLastError());
    return SSW_OK;
}

// *done = true when a pipeline ran
static bool pipe_col_available(const ssw_ctx* c, int w, int h) { return c->col_pipe > 0 && (w % 4) == 0 && (h == 2160 || h == 1080); }

static int pipe_col(ssw_ctx* c, bool inverse, int w, int h, int batch, float* d_plane, float scale0, float scalen, bool* done,
                    float* d_out = nullptr, const unsigned* col_limit = nullptr) {
    *done = false;
    if (c->col_pipe <= 0 || !aligned(d_plane, 16) || (w % 4) || (h & 1)) return SSW_OK;
    const char* name = inverse ? (col_limit ? "inv_cols_part" : "inv_cols") : "fwd_cols";
    int rc = SSW_OK;
    auto run = [&](auto k) { using K = decltype(k); rc = launch_col_pipe<K>(c, name, w, h, batch, d_plane, scale0, scalen, d_out, col_limit); *done = true; };
    if (h == 2160) {
        using P = fast::Plan2160;
        // SSW_COL_PIPE=1 (default) / 2: 2 teams x 2 rounds per tile (416 threads, registers to spare) with the split schedule
        // (half tiles) and 32-byte swizzled tile buffers; =4: 4 teams, one round per tile (800 threads at the 72-register cap
        // of 25 warps per SM); =3: 4-column tiles, two CTAs per SM.  Measured on one 4K frame with the spill-free build
        // (profiles/r2b_*): step 0.193 ms (2 teams) / 0.198 ms (4 teams).
        if (c->col_pipe == 4) { if (inverse) run(fast::ColPipe<P, 4, 4, true>{}); else run(fast::ColPipe<P, 4, 4, false>{}); }
        else if (c->col_pipe == 3) { if (inverse) run(fast::ColPipe<P, 2, 2, true, 2>{}); else run(fast::ColPipe<P, 2, 2, false, 2>{}); }
        else { if (inverse) run(fast::ColPipe<P, 4, 2, true>{}); else run(fast::ColPipe<P, 4, 2, false>{}); }   // SSW_COL_PIPE=1 (default), 2
    } else if (h == 1080) {
        using P = fast::Plan1080;
        if (c->col_pipe == 2) { if (inverse) run(fast::ColPipe<P, 4, 4, true, 1>{}); else run(fast::ColPipe<P, 4, 4, false, 1>{}); }
        else { if (inverse) run(fast::ColPipe<P, 4, 4, true, 2>{}); else run(fast::ColPipe<P, 4, 4, false, 2>{}); }
    }
    return rc;
}

static int fast_col(ssw_ctx* c, bool inverse, int w, int h, int batch, float* d_plane, float scale0, float scalen,
                    bool* done) {
    *done = false;
    if (!c->use_fast || (w % 4) || !aligned(d_plane, 16)) return SSW_OK;
    CKS(pipe_col(c, inverse, w, h, batch, d_plane, scale0, scalen, done));
    if (*done) return SSW_OK;
    int rc = SSW_OK;
    *done = fast::with_plan(h, [&](auto p) {
        using P = decltype(p);
        fast::FastArgs a = fast_args(w, h);
        a.plane = d_plane; a.scale0 = scale0; a.scalen = scalen;
        if (inverse) rc = launch_col_variant<P, true>(c, a, w, h, batch);
        else rc = launch_col_variant<P, false>(c, a, w, h, batch);
    });
    if (rc == SSW_ERR_UNSUPPORTED) { *done = false; rc = SSW_OK; }
    return rc;
}

static int fast_row_inv(ssw_ctx* c, float* d_plane, int src_type, const void* d_src, int w, int h, int batch,
                        int dst_type, void* d_dst, float scale, bool* done) {
    *done = false;
    if (!c->use_fast || !aligned(d_plane, 16)) return SSW_OK;
    if (dst_type == PIX_PLANE) {
        if (!aligned(d_dst, 16)) return SSW_OK;
    } else {
        if (src_type == PIX_PLANE) return SSW_OK;
        if (!aligned(d_dst, dst_type == PIX_RGB8 ? 4 : 16) || !aligned(d_src, src_type == PIX_RGB8 ? 4 : 16)) return SSW_OK;
    }
    int rc = SSW_OK;
    if (dst_type == PIX_RGB8 && src_type == PIX_RGB8) {
        CKS(pipe_row(c, true, d_src, d_plane, d_dst, w, h, batch, scale, scale, done));
        if (*done) return SSW_OK;
    }
    *done = fast::with_plan(w, [&](auto p) {
        using P = decltype(p);
        constexpr int G = fast::RowG<P>::value;
        fast::FastArgs a = fast_args(w, h);
        a.src = d_src; a.plane = d_plane; a.dst = d_dst; a.scale0 = scale;
        apply_seg(c, &a);
        if (dst_type == PIX_PLANE) rc = launch_fast<fast::RowInv<P, G, PIX_PLANE, PIX_PLANE>>(c, "inv_rows_plane", a, w, h, batch);
        else if (dst_type == PIX_RGB8 && src_type == PIX_RGB8) rc = launch_fast<fast::RowInv<P, G, PIX_RGB8, PIX_RGB8>>(c, "inv_rows", a, w, h, batch);
        else if (dst_type == PIX_RGB8) rc = launch_fast<fast::RowInv<P, G, PIX_RGB8, PIX_RGB32F>>(c, "inv_rows_src32f", a, w, h, batch);
        else if (src_type == PIX_RGB8) rc = launch_fast<fast::RowInv<P, G, PIX_RGB32F, PIX_RGB8>>(c, "inv_rows_rgb32f_src8", a, w, h, batch);
        else rc = launch_fast<fast::RowInv<P, G, PIX_RGB32F, PIX_RGB32F>>(c, "inv_rows_rgb32f", a, w, h, batch);
    });
    return rc;
}

// forward: pixels/plane -> coefficient plane (rows then columns)
// single-line kernels (lines too long for a packed pair, e.g. 32768; SSW_FORCE_LINE1=1 prefers them wherever
// a plan exists so that the GPU tests can compare them with the pair kernels)
static int line1_row_fwd(ssw_ctx* c, int src_type, const void* d_src, int w, int h, int batch, float* d_plane,
                         float scale0, float scalen, bool* done) {
    *done = false;
    if (src_type != PIX_RGB8 && src_type != PIX_PLANE) return SSW_OK;
    if (!aligned(d_plane, 16) || !aligned(d_src, src_type == PIX_RGB8 ? 4 : 16)) return SSW_OK;
    int rc = SSW_OK;
    *done = fast::with_line1_plan(w, [&](auto p) {
        using P = decltype(p);
        fast::FastArgs a = fast_args(w, h);
        a.src = d_src; a.plane = d_plane; a.scale0 = scale0; a.scalen = scalen;
        apply_seg(c, &a);
        if (src_type == PIX_RGB8) rc = launch_fast<fast::Line1Fwd<P, PIX_RGB8>, true>(c, "fwd_line1", a, w, h, batch);
        else rc = launch_fast<fast::Line1Fwd<P, PIX_PLANE>, true>(c, "fwd_line1_plane", a, w, h, batch);
    });
    return rc;
}

static int line1_row_inv(ssw_ctx* c, float* d_plane, int src_type, const void* d_src, int w, int h, int batch,
                         int dst_type, void* d_dst, float scale, bool* done) {
    *done = false;
    if (!aligned(d_plane, 16)) return SSW_OK;
    const bool rgb8 = dst_type == PIX_RGB8 && src_type == PIX_RGB8 && aligned(d_src, 4) && aligned(d_dst, 4);
    const bool plane = dst_type == PIX_PLANE && aligned(d_dst, 16);
    if (!rgb8 && !plane) return SSW_OK;
    int rc = SSW_OK;
    *done = fast::with_line1_plan(w, [&](auto p) {
        using P = decltype(p);
        fast::FastArgs a = fast_args(w, h);
        a.src = d_src; a.plane = d_plane; a.dst = d_dst; a.scale0 = scale;
        apply_seg(c, &a);
        if (rgb8) rc = launch_fast<fast::Line1Inv<P, PIX_RGB8, PIX_RGB8>, true>(c, "inv_line1", a, w, h, batch);
        else rc = launch_fast<fast::Line1Inv<P, PIX_PLANE, PIX_PLANE>, true>(c, "inv_line1_plane", a, w, h, batch);
    });
    return rc;
}

// row pass of the forward transform: pixels / plane rows -> DCT-II along x (x scale0/scalen) -> plane
static int run_rows_forward(ssw_ctx* c, int src_type, const void* d_src, int w, int h, int batch, float* d_plane,
                            float rs0, float rsn) {
    const long long npix = (long long)w * h;
    bool done = false;
    if (c->force_line1) {
        CKS(line1_row_fwd(c, src_type, d_src, w, h, batch, d_plane, rs0, rsn, &done));
        if (done) return SSW_OK;
    }
    CKS(fast_row_fwd(c, src_type, d_src, w, h, batch, d_plane, rs0, rsn, &done));
    if (done) return SSW_OK;
    if (w > 16384 || c->seg.active) {
        CKS(line1_row_fwd(c, src_type, d_src, w, h, batch, d_plane, rs0, rsn, &done));
        if (done) return SSW_OK;
    }
    if (c->seg.active) return fail(SSW_ERR_UNSUPPORTED, "segmented source lines need a planned line length");
    const DevPlan* pw;
    CKS(get_plan(c, w, &pw));
    Tiling tr;
    CKS(pick_tiling(c, pw->dev, false, h, &tr));
    LineArgs ar = base_args(pw->dev, w, h, tr, npix);
    ar.src = d_src; ar.plane = d_plane; ar.scale0 = rs0; ar.scalen = rsn;
    ar.tiles_per_image = (h + 2 * tr.P - 1) / (2 * tr.P);
    const long long ntr = (long long)ar.tiles_per_image * batch;
    switch (src_type) {
        case PIX_RGB8: return launch_line(c, "row_fwd_rgb8", row_fwd_kernel<PIX_RGB8>, ar, tr, ntr);
        case PIX_RGB32F: return launch_line(c, "row_fwd_rgb32f", row_fwd_kernel<PIX_RGB32F>, ar, tr, ntr);
        default: return launch_line(c, "row_fwd_plane", row_fwd_kernel<PIX_PLANE>, ar, tr, ntr);
    }
}

static int run_forward(ssw_ctx* c, int src_type, const void* d_src, int w, int h, int batch, float* d_plane,
                       int dct_type) {
    float rs0 = 1.f, rsn = 1.f, cs0 = 1.f, csn = 1.f;
    if (dct_type == SSW_DCT2_ORTHOGONAL) {  // src/dct2d.rs:153-162,189-198
        rs0 = std::sqrt(1.0f / (4.0f * (float)w)); rsn = std::sqrt(1.0f / (2.0f * (float)w));
        cs0 = std::sqrt(1.0f / (4.0f * (float)h)); csn = std::sqrt(1.0f / (2.0f * (float)h));
    }
    const long long npix = (long long)w * h;
    CKS(run_rows_forward(c, src_type, d_src, w, h, batch, d_plane, rs0, rsn));
    bool done = false;
    CKS(fast_col(c, false, w, h, batch, d_plane, cs0, csn, &done));
    if (!done) {
        const DevPlan* ph;
        CKS(get_plan(c, h, &ph));
        Tiling tc;
        CKS(pick_tiling(c, ph->dev, true, w, &tc));
        LineArgs ac = base_args(ph->dev, w, h, tc, npix);
        ac.plane = d_plane; ac.scale0 = cs0; ac.scalen = csn;
        ac.tiles_per_image = (w + 2 * tc.P - 1) / (2 * tc.P);
        CKS(launch_line(c, "col_fwd", col_fwd_kernel, ac, tc, (long long)ac.tiles_per_image * batch));
    }
    return SSW_OK;
}

// inverse: coefficient plane (destroyed) -> pixels/plane (columns then rows)
// row pass of the inverse transform: plane rows -> DCT-III along x -> x out_scale -> plane | pixels
static int run_rows_inverse(ssw_ctx* c, float* d_plane, int src_type, const void* d_src, int w, int h, int batch,
                            int dst_type, void* d_dst, float out_scale) {
    const long long npix = (long long)w * h;
    bool done = false;
    if (c->force_line1) {
        CKS(line1_row_inv(c, d_plane, src_type, d_src, w, h, batch, dst_type, d_dst, out_scale, &done));
        if (done) return SSW_OK;
    }
    CKS(fast_row_inv(c, d_plane, src_type, d_src, w, h, batch, dst_type, d_dst, out_scale, &done));
    if (done) return SSW_OK;
    if (w > 16384 || c->seg.active) {
        CKS(line1_row_inv(c, d_plane, src_type, d_src, w, h, batch, dst_type, d_dst, out_scale, &done));
        if (done) return SSW_OK;
    }
    if (c->seg.active) return fail(SSW_ERR_UNSUPPORTED, "segmented source lines need a planned line length");
    const DevPlan* pw;
    CKS(get_plan(c, w, &pw));
    Tiling tr;
    CKS(pick_tiling(c, pw->dev, false, h, &tr));
    LineArgs ar = base_args(pw->dev, w, h, tr, npix);
    ar.plane = d_plane; ar.src = d_src; ar.dst = d_dst;
    ar.scale0 = out_scale;
    ar.tiles_per_image = (h + 2 * tr.P - 1) / (2 * tr.P);
    const long long ntr = (long long)ar.tiles_per_image * batch;
    if (dst_type == PIX_PLANE) return launch_line(c, "row_inv_plane", row_inv_kernel<PIX_PLANE, PIX_PLANE>, ar, tr, ntr);
    if (dst_type == PIX_RGB8 && src_type == PIX_RGB8) return launch_line(c, "row_inv_rgb8", row_inv_kernel<PIX_RGB8, PIX_RGB8>, ar, tr, ntr);
    if (dst_type == PIX_RGB8 && src_type == PIX_RGB32F) return launch_line(c, "row_inv_rgb8_src32f", row_inv_kernel<PIX_RGB8, PIX_RGB32F>, ar, tr, ntr);
    if (dst_type == PIX_RGB32F && src_type == PIX_RGB8) return launch_line(c, "row_inv_rgb32f_src8", row_inv_kernel<PIX_RGB32F, PIX_RGB8>, ar, tr, ntr);
    if (dst_type == PIX_RGB32F && src_type == PIX_RGB32F) return launch_line(c, "row_inv_rgb32f", row_inv_kernel<PIX_RGB32F, PIX_RGB32F>, ar, tr, ntr);
    return fail(SSW_ERR_INVALID, "bad pixel type combination");
}

static int run_inverse(ssw_ctx* c, float* d_plane, int src_type, const void* d_src, int w, int h, int batch,
                       int dst_type, void* d_dst) {
    const long long npix = (long long)w * h;
    const float out_scale = 4.0f / (float)((size_t)w * (size_t)h);  // src/dct2d.rs:213-217
    bool done = false;
    CKS(fast_col(c, true, w, h, batch, d_plane, 1.f, 1.f, &done));
    if (!done) {
        const DevPlan* ph;
        CKS(get_plan(c, h, &ph));
        Tiling tc;
        CKS(pick_tiling(c, ph->dev, true, w, &tc));
        LineArgs ac = base_args(ph->dev, w, h, tc, npix);
        ac.plane = d_plane;
        ac.tiles_per_image = (w + 2 * tc.P - 1) / (2 * tc.P);
        CKS(launch_line(c, "col_inv", col_inv_kernel, ac, tc, (long long)ac.tiles_per_image * batch));
    }
    return run_rows_inverse(c, d_plane, src_type, d_src, w, h, batch, dst_type, d_dst, out_scale);
}

// ------------------------------------------------------------------------------------------------
// top-k
// ------------------------------------------------------------------------------------------------
static int ensure_topk_scratch(ssw_ctx* c, unsigned batch) {
    if (batch <= c->ts_batch) return SSW_OK;
    CK(cudaStreamSynchronize(c->stream));
    topk_scratch_free(c);
    const unsigned b = std::max(batch, 16u);
    CK(cudaMalloc(&c->ts.hist, (size_t)b * kHistBins * sizeof(unsigned)));
    CK(cudaMalloc(&c->ts.ticket, b * sizeof(unsigned)));
    CK(cudaMalloc(&c->ts.sel_bin, b * sizeof(unsigned)));
    CK(cudaMalloc(&c->ts.cand_count, b * sizeof(unsigned)));
    CK(cudaMalloc(&c->ts.cand, (size_t)b * kTopkCap * sizeof(unsigned long long)));
    CK(cudaMalloc(&c->ts.overflow, sizeof(unsigned)));
    CK(cudaMalloc(&c->ts.maxrow, b * sizeof(unsigned)));
    CK(cudaMemset(c->ts.maxrow, 0, b * sizeof(unsigned)));
    CK(cudaMalloc(&c->ts.maxcol, (1 + (size_t)b) * sizeof(unsigned)));   // [0]: the launch, [1 + i]: image i
    CK(cudaMemset(c->ts.maxcol, 0, (1 + (size_t)b) * sizeof(unsigned)));
    c->ts.maxcol_img = c->ts.maxcol + 1;
    CK(cudaMemset(c->ts.sel_bin, 0, b * sizeof(unsigned)));   // bit 31 = "bin published" (collecting column pipeline)
    CK(cudaMemset(c->ts.hist, 0, (size_t)b * kHistBins * sizeof(unsigned)));
    CK(cudaMemset(c->ts.ticket, 0, b * sizeof(unsigned)));
    CK(cudaMemset(c->ts.cand_count, 0, b * sizeof(unsigned)));
    CK(cudaMemset(c->ts.overflow, 0, sizeof(unsigned)));
    c->ts_batch = b;
    return SSW_OK;
}

static OrderConsts make_order(int ordering, int w, int h) {
    OrderConsts oc;
    oc.mode = ordering; oc.w = w;
    oc.t_ld = 0u; oc.t_col0 = 0u;
    oc.s_k0_w = std::sqrt(1.0f / (4.0f * (float)w));
    oc.s_k0_h = std::sqrt(1.0f / (4.0f * (float)h));
    oc.s_w = std::sqrt(1.0f / (2.0f * (float)w));
    oc.s_h = std::sqrt(1.0f / (2.0f * (float)h));
    return oc;
}

// fast path: k + (one histogram bin of elements) must fit kTopkCap; no host synchronisation.
// full_hist = false: selection bin from the low-frequency block -- from the histogram the forward column pipeline left
//                    in ts.hist (hist_ready, see ssw_ctx::col_hist) or from topk_block_bin -- one pass over the plane;
// full_hist = true : selection bin from a histogram of the whole plane (two passes) -- the repair path.
// ap: what the ranking kernel does with (rank, index) beyond storing the index list (embed / extract [+ score]).
static int run_topk_fast(ssw_ctx* c, const float* d_planes, int w, int h, unsigned batch, int ordering,
                         unsigned k, unsigned* d_idx, long long idx_stride, bool full_hist, int hist_ready = 0,
                         const TopkApply* ap = nullptr, cudaEvent_t join_before_apply = nullptr) {
    CKS(ensure_topk_scratch(c, batch));
    const unsigned n = (unsigned)((size_t)w * h);
    const OrderConsts oc = make_order(ordering, w, h);
    const long long stride = (long long)n;
    unsigned blocks = (unsigned)std::min<size_t>(((size_t)n / 4 + 511) / 512, (size_t)std::max(1u, (unsigned)(c->sm_count * 4) / std::min(batch, (unsigned)(c->sm_count * 4))));
    blocks = std::max(1u, blocks);
    // the collecting scan runs as exactly one wave of its own occupancy (its loop keeps four 16-byte loads in flight per thread)
    if (!c->collect_occ) {
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, topk_collect_kernel, 512, 0));
        c->collect_occ = std::max(1, occ);
    }
    const unsigned wave = (unsigned)(c->sm_count * c->collect_occ);
    const unsigned cblocks = std::max(1u, (unsigned)std::min<size_t>(((size_t)n / 4 + 511) / 512, (size_t)std::max(1u, wave / std::min(batch, wave))));
    // hist_ready: the forward column pipeline has already left the selection bin of the low-frequency block in ts.sel_bin
    if (full_hist && hist_ready) return fail(SSW_ERR_STATE, "selection bin requested from both the block and the full plane");
    unsigned* rank_tile_max = nullptr;   // tile maxima consumed by the tile-wise scan: the ranking kernel clears them
    unsigned rank_tile_count = 0;
    for (unsigned b0 = 0; b0 < batch; b0 += 65535) {
        const unsigned nb = std::min(65535u, batch - b0);
        TopkScratch ts = c->ts;
        ts.hist += (size_t)b0 * kHistBins; ts.ticket += b0; ts.sel_bin += b0; ts.cand_count += b0; ts.maxrow += b0; ts.maxcol_img += b0;
        ts.cand += (size_t)b0 * kTopkCap;
        if (full_hist) {
            KScope ks(c, "topk_hist");
            launch_pdl(c, topk_hist_kernel, dim3(blocks, nb), 512, 0, c->stream, d_planes + (size_t)b0 * n, stride, n, k, oc, ts);
        } else if (!hist_ready) {
            KScope ks(c, "topk_block_bin");
            launch_pdl(c, topk_block_bin_kernel, dim3(nb), kBinThreads, 0, c->stream, d_planes + (size_t)b0 * n, stride, (unsigned)w, (unsigned)h, k, oc, ts);
        }
        const bool by_tiles = !full_hist && c->tile_max.tiles && c->tile_max.tiles <= (unsigned)kTileMaxTiles && c->tile_max.buf && c->tile_max.plane == d_planes && ordering == 0 && batch <= 65535u && (w % 4) == 0 &&
                              aligned(d_planes, 16) && (size_t)c->tile_max.tiles * batch <= c->tile_max.cap;
        if (hist_ready < 2 && by_tiles) {
            // the forward column pipeline left the largest |coefficient| of every column tile: only tiles that can hold a candidate are read
            KScope ks(c, "topk_collect_tiles");
            launch_pdl(c, topk_collect_tiles_kernel, dim3(kTileSplit, nb), 256, 0, c->stream, d_planes + (size_t)b0 * n, stride, (unsigned)w, (unsigned)h,
                       c->tile_max.cols, c->tile_max.tiles, ts, (const unsigned*)(c->tile_max.buf + (size_t)b0 * c->tile_max.tiles));
            rank_tile_max = c->tile_max.buf + (size_t)b0 * c->tile_max.tiles; rank_tile_count = c->tile_max.tiles;
        } else if (hist_ready < 2) {   // 2: the forward column pipeline has appended the candidates as well (PipeArgs::collect)
            KScope ks(c, "topk_collect"); launch_pdl(c, topk_collect_kernel, dim3(cblocks, nb), 512, 0, c->stream, d_planes + (size_t)b0 * n, stride, n, oc, ts);
        }
        CK(cudaGetLastError());
    }
    c->tile_max.tiles = 0;   // (consumed, or not applicable: the next forward pass has to produce them again)
    if (join_before_apply) CK(cudaStreamWaitEvent(c->stream, join_before_apply, 0));   // extract: the derived planes
    const void* key = (const void*)topk_rank_kernel;
    const int smem = kTopkCap * (int)sizeof(unsigned long long);
    if (c->smem_attr.find(key) == c->smem_attr.end()) {
        CK(cudaFuncSetAttribute(topk_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        c->smem_attr[key] = smem;
    }
    for (unsigned b0 = 0; b0 < batch; b0 += 65535) {
        const unsigned nb = std::min(65535u, batch - b0);
        TopkScratch ts = c->ts;
        ts.hist += (size_t)b0 * kHistBins; ts.ticket += b0; ts.sel_bin += b0; ts.cand_count += b0; ts.cand += (size_t)b0 * kTopkCap; ts.maxrow += b0; ts.maxcol_img += b0;
        ts.tile_max = rank_tile_max; ts.tile_count = rank_tile_count;   // (set only by the single-launch tile-wise scan above)
        TopkApply a;
        std::memset(&a, 0, sizeof(a));
        if (ap) {
            a = *ap;
            if (a.planes) a.planes += (size_t)b0 * a.plane_stride;
            if (a.derived) a.derived += (size_t)b0 * a.plane_stride;
            if (a.marks) a.marks += (size_t)b0 * a.mark_stride;
            if (a.out) a.out += (size_t)b0 * a.out_stride;
            if (a.sim) a.sim += b0;
        }
        KScope ks(c, ap && (ap->mode == 1 || ap->mode == 3) ? "topk_rank_embed" : (ap && ap->mode == 2 ? "topk_rank_extract" : "topk_rank"));
        launch_pdl(c, topk_rank_kernel, dim3(kRankCtas, nb), kRankThreads, smem, c->stream, ts, k, d_idx + (long long)b0 * idx_stride, idx_stride, a);
    }
    CK(cudaGetLastError());
    return SSW_OK;
}

// read-and-clear the sticky overflow counter (synchronises the stream)
static int take_overflow(ssw_ctx* c, unsigned* out) {
    *out = 0;
    if (!c->ts_batch) return SSW_OK;
    CK(cudaMemcpyAsync(c->h_flag, c->ts.overflow, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemsetAsync(c->ts.overflow, 0, sizeof(unsigned), c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *out = *c->h_flag;
    return SSW_OK;
}

// exact top-k of ONE plane for any k (host-synchronising): fast path when it applies, otherwise the
// general radix-sort path; a fast-path overflow is detected and repaired here.
static int run_topk_exact(ssw_ctx* c, const float* d_plane, int w, int h, int ordering, size_t k,
                          unsigned* d_idx) {
    const size_t n = (size_t)w * h;
    if (k == 0) return SSW_OK;
    if (k > n - 1) return fail(SSW_ERR_INVALID, "k exceeds the number of AC coefficients");
    if (k <= (size_t)kTopkCap / 2) {
        for (int attempt = c->topk_full_hist ? 1 : 0; attempt < 2; ++attempt) {
            CKS(run_topk_fast(c, d_plane, w, h, 1, ordering, (unsigned)k, d_idx, 0, attempt == 1));
            unsigned ov = 0;
            CKS(take_overflow(c, &ov));
            if (!ov) return SSW_OK;
        }
        c->last_fallbacks += 1;
    }
    const OrderConsts oc = make_order(ordering, w, h);
    int rc;
    { KScope ks(c, "general_select", 0); rc = c->general.run(c->stream, d_plane, (unsigned)n, oc, k, d_idx, &c->launches); }
    if (rc != SSW_OK) return fail(rc, c->general.error);
    return SSW_OK;
}

// ------------------------------------------------------------------------------------------------
// dct2d::dct2_2d
// ------------------------------------------------------------------------------------------------
static int check_dims(uint32_t w, uint32_t h) {
    if (w == 0 || h == 0) return fail(SSW_ERR_INVALID, "image dimensions must be non-zero");
    if ((uint64_t)w * h >= 0xFFFFFFFFull) return fail(SSW_ERR_UNSUPPORTED, "more than 2^32-1 pixels per frame");
    return SSW_OK;
}

extern "C" int ssw_dct2_2d_dev(ssw_ctx* c, int type, uint32_t w, uint32_t h, float* d) {
    if (!c || !d) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_dims(w, h));
    CKS(ctx_bind(c));
    if (type == SSW_DCT2 || type == SSW_DCT2_ORTHOGONAL) return run_forward(c, PIX_PLANE, d, w, h, 1, d, type);
    if (type == SSW_DCT3) return run_inverse(c, d, PIX_PLANE, nullptr, w, h, 1, PIX_PLANE, d);
    return fail(SSW_ERR_INVALID, "unknown transform type");
}

extern "C" int ssw_dct2_2d(ssw_ctx* c, int type, uint32_t w, uint32_t h, float* data) {
    if (!c || !data) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_dims(w, h));
    CKS(ctx_bind(c));
    const size_t bytes = (size_t)w * h * sizeof(float);
    float* d = nullptr;
    CK(cudaMallocFromPoolAsync(&d, bytes, c->pool, c->stream));
    CK(cudaMemcpyAsync(d, data, bytes, cudaMemcpyHostToDevice, c->stream));
    int rc = ssw_dct2_2d_dev(c, type, w, h, d);
    if (rc == SSW_OK) {
        cudaError_t e = cudaMemcpyAsync(data, d, bytes, cudaMemcpyDeviceToHost, c->stream);
        if (e != cudaSuccess) rc = fail(SSW_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFreeAsync(d, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (rc == SSW_OK && e != cudaSuccess) rc = fail(SSW_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}

// ------------------------------------------------------------------------------------------------
// yiq planes
// ------------------------------------------------------------------------------------------------
extern "C" int ssw_rgb32f_to_yiq(ssw_ctx* c, const float* rgb, uint32_t w, uint32_t h, float* y, float* i, float* q) {
    if (!c || !rgb || !y || !i || !q) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_dims(w, h));
    CKS(ctx_bind(c));
    const size_t np = (size_t)w * h;
    float* d = nullptr;
    CK(cudaMallocFromPoolAsync(&d, np * 6 * sizeof(float), c->pool, c->stream));
    CK(cudaMemcpyAsync(d, rgb, np * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    float* dy = d + 3 * np;
    { KScope ks(c, "rgb32f_to_yiq"); rgb32f_to_yiq_kernel<<<(unsigned)((np + 255) / 256), 256, 0, c->stream>>>(d, np, dy, dy + np, dy + 2 * np); }
    CK(cudaMemcpyAsync(y, dy, np * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(i, dy + np, np * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(q, dy + 2 * np, np * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaFreeAsync(d, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SSW_OK;
}

extern "C" int ssw_yiq_to_rgb32f(ssw_ctx* c, const float* y, const float* i, const float* q, uint32_t w, uint32_t h, float* rgb) {
    if (!c || !rgb || !y || !i || !q) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_dims(w, h));
    CKS(ctx_bind(c));
    const size_t np = (size_t)w * h;
    float* d = nullptr;
    CK(cudaMallocFromPoolAsync(&d, np * 6 * sizeof(float), c->pool, c->stream));
    CK(cudaMemcpyAsync(d, y, np * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d + np, i, np * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d + 2 * np, q, np * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    { KScope ks(c, "yiq_to_rgb32f"); yiq_to_rgb32f_kernel<<<(unsigned)((np + 255) / 256), 256, 0, c->stream>>>(d, d + np, d + 2 * np, np, d + 3 * np); }
    CK(cudaMemcpyAsync(rgb, d + 3 * np, np * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaFreeAsync(d, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SSW_OK;
}

// ------------------------------------------------------------------------------------------------
// config validation
// ------------------------------------------------------------------------------------------------
static int check_cfg(const ssw_config* cfg) {
    if (!cfg) return fail(SSW_ERR_INVALID, "config is NULL");
    if (cfg->method < 1 || cfg->method > 3)
        return fail(SSW_ERR_UNSUPPORTED, "only Option1/2/3 insertion/extraction run on the device (Custom closures are host code)");
    if (cfg->ordering < 0 || cfg->ordering > 2)
        return fail(SSW_ERR_UNSUPPORTED, "only Energy/EnergyOrthogonal/Legacy orderings run on the device (Custom closures are host code)");
    return SSW_OK;
}

// ------------------------------------------------------------------------------------------------
// Writer
// ------------------------------------------------------------------------------------------------
struct ssw_writer {
    CtxRef ctx;
    uint32_t w, h;
    ssw_config cfg;
    int src_type;
    void* d_src;  // original pixels (owned unless borrowed)
    bool own_src;
    float* d_plane;
    unsigned* d_idx;
    size_t k_cached;
    bool consumed;
    bool embedded;   // an embed has modified the coefficients: the cached ordering can no longer be extended
};

static int writer_new(ssw_ctx* c, int src_type, const void* src, bool src_on_device, uint32_t w, uint32_t h,
                      const ssw_config* cfg, ssw_writer** out) {
    if (!c || !src || !out) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_cfg(cfg));
    CKS(check_dims(w, h));
    CKS(ctx_bind(c));
    auto wr = std::make_unique<ssw_writer>();
    wr->ctx = c; wr->w = w; wr->h = h; wr->cfg = *cfg; wr->src_type = src_type;
    wr->d_idx = nullptr; wr->k_cached = 0; wr->consumed = false; wr->embedded = false;
    const size_t np = (size_t)w * h;
    const size_t src_bytes = np * 3 * (src_type == PIX_RGB8 ? 1 : 4);
    if (src_on_device) {
        wr->d_src = const_cast<void*>(src);
        wr->own_src = false;
    } else {
        CK(cudaMallocFromPoolAsync(&wr->d_src, src_bytes, c->pool, c->stream));
        wr->own_src = true;
        CK(cudaMemcpyAsync(wr->d_src, src, src_bytes, cudaMemcpyHostToDevice, c->stream));
    }
    CK(cudaMallocFromPoolAsync(&wr->d_plane, np * sizeof(float), c->pool, c->stream));
    int rc = run_forward(c, src_type, wr->d_src, w, h, 1, wr->d_plane, SSW_DCT2);
    if (!src_on_device) {
        // the upload reads caller memory asynchronously: finish before returning control
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (rc == SSW_OK && e != cudaSuccess) rc = fail(SSW_ERR_CUDA, cudaGetErrorString(e));
    }
    if (rc != SSW_OK) {
        if (wr->own_src) cudaFreeAsync(wr->d_src, c->stream);
        cudaFreeAsync(wr->d_plane, c->stream);
        return rc;
    }
    *out = wr.release();
    return SSW_OK;
}

extern "C" int ssw_writer_new_rgb8(ssw_ctx* c, const uint8_t* rgb, uint32_t w, uint32_t h, const ssw_config* cfg, ssw_writer** out) {
    return writer_new(c, PIX_RGB8, rgb, false, w, h, cfg, out);
}
extern "C" int ssw_writer_new_rgb32f(ssw_ctx* c, const float* rgb, uint32_t w, uint32_t h, const ssw_config* cfg, ssw_writer** out) {
    return writer_new(c, PIX_RGB32F, rgb, false, w, h, cfg, out);
}
extern "C" int ssw_writer_new_rgb8_dev(ssw_ctx* c, const uint8_t* rgb, uint32_t w, uint32_t h, const ssw_config* cfg, ssw_writer** out) {
    return writer_new(c, PIX_RGB8, rgb, true, w, h, cfg, out);
}

static int ensure_indices(ssw_ctx* c, const float* d_plane, uint32_t w, uint32_t h, int ordering, size_t k,
                          unsigned** d_idx, size_t* k_cached) {
    if (k <= *k_cached) return SSW_OK;
    if (*d_idx) { CK(cudaFreeAsync(*d_idx, c->stream)); *d_idx = nullptr; *k_cached = 0; }
    CK(cudaMallocFromPoolAsync(d_idx, k * sizeof(unsigned), c->pool, c->stream));
    CKS(run_topk_exact(c, d_plane, w, h, ordering, k, *d_idx));
    *k_cached = k;
    return SSW_OK;
}

extern "C" int ssw_writer_embed(ssw_writer* wr, const float* const* marks, const size_t* lens, size_t n_marks) {
    if (!wr || (n_marks && (!marks || !lens))) return fail(SSW_ERR_INVALID, "NULL argument");
    if (wr->consumed) return fail(SSW_ERR_STATE, "writer already consumed by result()");
    if (n_marks == 0) return SSW_OK;
    ssw_ctx* c = wr->ctx;
    CKS(ctx_bind(c));
    const size_t nac = (size_t)wr->w * wr->h - 1;
    size_t kmax = 0;
    std::vector<unsigned> hl(n_marks);
    for (size_t m = 0; m < n_marks; ++m) {
        if (lens[m] && !marks[m]) return fail(SSW_ERR_INVALID, "mark pointer is NULL");
        const size_t l = std::min(lens[m], nac);  // zip truncation, src/algorithm.rs:396
        hl[m] = (unsigned)l;
        kmax = std::max(kmax, l);
    }
    if (kmax == 0) return SSW_OK;
    // the reference orders once, in Writer::new, from the original coefficients (src/algorithm.rs:324-327); here the ordering
    // is computed for the length first asked for -- a later, longer mark would be ordered on modified coefficients
    if (wr->embedded && kmax > wr->k_cached)
        return fail(SSW_ERR_STATE, "a second embed needs more ordered coefficients than the first one fixed; embed the longest mark first "
                                   "or call ssw_writer_indices(n) before the first embed");
    CKS(ensure_indices(c, wr->d_plane, wr->w, wr->h, wr->cfg.ordering, kmax, &wr->d_idx, &wr->k_cached));
    wr->embedded = true;
    // stage marks as [n_marks][kmax] zero-padded + lens
    std::vector<float> stage(n_marks * kmax, 0.f);
    for (size_t m = 0; m < n_marks; ++m) std::memcpy(&stage[m * kmax], marks[m], hl[m] * sizeof(float));
    float* d_marks = nullptr;
    unsigned* d_lens = nullptr;
    CK(cudaMallocFromPoolAsync(&d_marks, stage.size() * sizeof(float), c->pool, c->stream));
    CK(cudaMallocFromPoolAsync(&d_lens, n_marks * sizeof(unsigned), c->pool, c->stream));
    CK(cudaMemcpyAsync(d_marks, stage.data(), stage.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_lens, hl.data(), n_marks * sizeof(unsigned), cudaMemcpyHostToDevice, c->stream));
    {
        KScope ks(c, "embed_scatter");
        launch_pdl(c, embed_scatter_kernel, dim3((unsigned)((kmax + 255) / 256), 1), 256, 0, c->stream, 
            wr->d_plane, 0, wr->d_idx, 0, (unsigned)kmax, d_marks, (long long)kmax, (int)n_marks, d_lens,
            wr->cfg.method, wr->cfg.alpha);
    }
    CK(cudaGetLastError());
    CK(cudaFreeAsync(d_marks, c->stream));
    CK(cudaFreeAsync(d_lens, c->stream));
    CK(cudaStreamSynchronize(c->stream));  // `stage` / `hl` are pageable host memory
    return SSW_OK;
}

extern "C" int ssw_writer_coefficients(ssw_writer* wr, float* out) {
    if (!wr || !out) return fail(SSW_ERR_INVALID, "NULL argument");
    if (wr->consumed) return fail(SSW_ERR_STATE, "writer already consumed by result()");
    ssw_ctx* c = wr->ctx;
    CKS(ctx_bind(c));
    CK(cudaMemcpyAsync(out, wr->d_plane, (size_t)wr->w * wr->h * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SSW_OK;
}

static int download_indices(ssw_ctx* c, const unsigned* d_idx, uint64_t* out, size_t n) {
    std::vector<unsigned> tmp(n);
    CK(cudaMemcpyAsync(tmp.data(), d_idx, n * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n; ++i) out[i] = tmp[i];
    return SSW_OK;
}

extern "C" int ssw_writer_indices(ssw_writer* wr, uint64_t* out, size_t n) {
    if (!wr || (!out && n)) return fail(SSW_ERR_INVALID, "NULL argument");
    if (n == 0) return SSW_OK;
    if (n > (size_t)wr->w * wr->h - 1) return fail(SSW_ERR_INVALID, "more indices requested than AC coefficients");
    if (wr->consumed) return fail(SSW_ERR_STATE, "writer already consumed by result()");
    ssw_ctx* c = wr->ctx;
    CKS(ctx_bind(c));
    if (n > wr->k_cached && wr->k_cached)
        return fail(SSW_ERR_STATE, "indices beyond the embedded length are undefined after embed (coefficients were modified)");
    CKS(ensure_indices(c, wr->d_plane, wr->w, wr->h, wr->cfg.ordering, n, &wr->d_idx, &wr->k_cached));
    return download_indices(c, wr->d_idx, out, n);
}

static int writer_result(ssw_writer* wr, int dst_type, void* out, bool out_on_device) {
    if (!wr || !out) return fail(SSW_ERR_INVALID, "NULL argument");
    if (wr->consumed) return fail(SSW_ERR_STATE, "writer already consumed by result()");
    ssw_ctx* c = wr->ctx;
    CKS(ctx_bind(c));
    const size_t np = (size_t)wr->w * wr->h;
    const size_t bytes = np * 3 * (dst_type == PIX_RGB8 ? 1 : 4);
    void* d_out = out;
    if (!out_on_device) CK(cudaMallocFromPoolAsync(&d_out, bytes, c->pool, c->stream));
    int rc = run_inverse(c, wr->d_plane, wr->src_type, wr->d_src, wr->w, wr->h, 1, dst_type, d_out);
    wr->consumed = true;
    if (!out_on_device) {
        if (rc == SSW_OK) {
            cudaError_t e = cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, c->stream);
            if (e != cudaSuccess) rc = fail(SSW_ERR_CUDA, cudaGetErrorString(e));
        }
        cudaFreeAsync(d_out, c->stream);
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (rc == SSW_OK && e != cudaSuccess) rc = fail(SSW_ERR_CUDA, cudaGetErrorString(e));
    }
    return rc;
}

extern "C" int ssw_writer_result_rgb8(ssw_writer* wr, uint8_t* out) { return writer_result(wr, PIX_RGB8, out, false); }
extern "C" int ssw_writer_result_rgb32f(ssw_writer* wr, float* out) { return writer_result(wr, PIX_RGB32F, out, false); }
extern "C" int ssw_writer_result_rgb8_dev(ssw_writer* wr, uint8_t* out) { return writer_result(wr, PIX_RGB8, out, true); }

extern "C" int ssw_writer_destroy(ssw_writer* wr) {
    if (!wr) return SSW_OK;
    ssw_ctx* c = wr->ctx;
    cudaSetDevice(c->device);
    if (wr->own_src && wr->d_src) cudaFreeAsync(wr->d_src, c->stream);
    if (wr->d_plane) cudaFreeAsync(wr->d_plane, c->stream);
    if (wr->d_idx) cudaFreeAsync(wr->d_idx, c->stream);
    delete wr;
    return SSW_OK;
}

// ------------------------------------------------------------------------------------------------
// Reader
// ------------------------------------------------------------------------------------------------
struct ssw_reader {
    CtxRef ctx;
    uint32_t w, h;
    bool is_base;
    ssw_config cfg;
    float* d_plane;
    unsigned* d_idx;
    size_t k_cached;
};

static int reader_new(ssw_ctx* c, int src_type, const void* src, bool src_on_device, uint32_t w, uint32_t h,
                      bool is_base, const ssw_config* cfg, ssw_reader** out) {
    if (!c || !src || !out) return fail(SSW_ERR_INVALID, "NULL argument");
    if (is_base) CKS(check_cfg(cfg));
    CKS(check_dims(w, h));
    CKS(ctx_bind(c));
    auto rd = std::make_unique<ssw_reader>();
    rd->ctx = c; rd->w = w; rd->h = h; rd->is_base = is_base;
    if (is_base) rd->cfg = *cfg; else rd->cfg = ssw_config{2, 0.1f, 0};
    rd->d_idx = nullptr; rd->k_cached = 0;
    const size_t np = (size_t)w * h;
    const size_t src_bytes = np * 3 * (src_type == PIX_RGB8 ? 1 : 4);
    void* d_src = const_cast<void*>(src);
    if (!src_on_device) {
        CK(cudaMallocFromPoolAsync(&d_src, src_bytes, c->pool, c->stream));
        CK(cudaMemcpyAsync(d_src, src, src_bytes, cudaMemcpyHostToDevice, c->stream));
    }
    CK(cudaMallocFromPoolAsync(&rd->d_plane, np * sizeof(float), c->pool, c->stream));
    int rc = run_forward(c, src_type, d_src, w, h, 1, rd->d_plane, SSW_DCT2);
    if (!src_on_device) {
        cudaFreeAsync(d_src, c->stream);
        // the upload reads caller memory asynchronously: finish before returning control
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (rc == SSW_OK && e != cudaSuccess) rc = fail(SSW_ERR_CUDA, cudaGetErrorString(e));
    }
    if (rc != SSW_OK) { cudaFreeAsync(rd->d_plane, c->stream); return rc; }
    *out = rd.release();
    return SSW_OK;
}

extern "C" int ssw_reader_base_rgb8(ssw_ctx* c, const uint8_t* rgb, uint32_t w, uint32_t h, const ssw_config* cfg, ssw_reader** out) {
    return reader_new(c, PIX_RGB8, rgb, false, w, h, true, cfg, out);
}
extern "C" int ssw_reader_base_rgb32f(ssw_ctx* c, const float* rgb, uint32_t w, uint32_t h, const ssw_config* cfg, ssw_reader** out) {
    return reader_new(c, PIX_RGB32F, rgb, false, w, h, true, cfg, out);
}
extern "C" int ssw_reader_derived_rgb8(ssw_ctx* c, const uint8_t* rgb, uint32_t w, uint32_t h, ssw_reader** out) {
    return reader_new(c, PIX_RGB8, rgb, false, w, h, false, nullptr, out);
}
extern "C" int ssw_reader_derived_rgb32f(ssw_ctx* c, const float* rgb, uint32_t w, uint32_t h, ssw_reader** out) {
    return reader_new(c, PIX_RGB32F, rgb, false, w, h, false, nullptr, out);
}
extern "C" int ssw_reader_base_rgb8_dev(ssw_ctx* c, const uint8_t* rgb, uint32_t w, uint32_t h, const ssw_config* cfg, ssw_reader** out) {
    return reader_new(c, PIX_RGB8, rgb, true, w, h, true, cfg, out);
}
extern "C" int ssw_reader_derived_rgb8_dev(ssw_ctx* c, const uint8_t* rgb, uint32_t w, uint32_t h, ssw_reader** out) {
    return reader_new(c, PIX_RGB8, rgb, true, w, h, false, nullptr, out);
}

static int reader_extract(ssw_reader* base, ssw_reader* derived, float* out, size_t n, bool out_on_device) {
    if (!base || !derived || (!out && n)) return fail(SSW_ERR_INVALID, "NULL argument");
    if (!base->is_base) return fail(SSW_ERR_STATE, "extract called on a derived reader (reference: unwrap on None, src/algorithm.rs:530)");
    if (base->ctx != derived->ctx) return fail(SSW_ERR_INVALID, "readers belong to different contexts");
    if ((size_t)base->w * base->h != (size_t)derived->w * derived->h)
        return fail(SSW_ERR_INVALID, "Derived coefficient length not equal to base coefficient length.");
    if (n >= (size_t)base->w * base->h)
        return fail(SSW_ERR_INVALID, "Desired extraction length exceeds available coefficients.");
    if (n == 0) return SSW_OK;
    ssw_ctx* c = base->ctx;
    CKS(ctx_bind(c));
    CKS(ensure_indices(c, base->d_plane, base->w, base->h, base->cfg.ordering, n, &base->d_idx, &base->k_cached));
    float* d_out = out;
    if (!out_on_device) CK(cudaMallocFromPoolAsync(&d_out, n * sizeof(float), c->pool, c->stream));
    {
        KScope ks(c, "extract_gather");
        launch_pdl(c, extract_gather_kernel, dim3((unsigned)((n + 255) / 256), 1), 256, 0, c->stream, 
            base->d_plane, derived->d_plane, 0, base->d_idx, 0, (unsigned)n, base->cfg.method, base->cfg.alpha, d_out, 0);
    }
    CK(cudaGetLastError());
    if (!out_on_device) {
        CK(cudaMemcpyAsync(out, d_out, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaFreeAsync(d_out, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return SSW_OK;
}

extern "C" int ssw_reader_extract(ssw_reader* b, ssw_reader* d, float* out, size_t n) { return reader_extract(b, d, out, n, false); }
extern "C" int ssw_reader_extract_dev(ssw_reader* b, ssw_reader* d, float* out, size_t n) { return reader_extract(b, d, out, n, true); }

extern "C" int ssw_reader_coefficients(ssw_reader* r, float* out) {
    if (!r || !out) return fail(SSW_ERR_INVALID, "NULL argument");
    ssw_ctx* c = r->ctx;
    CKS(ctx_bind(c));
    CK(cudaMemcpyAsync(out, r->d_plane, (size_t)r->w * r->h * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SSW_OK;
}

extern "C" int ssw_reader_indices(ssw_reader* r, uint64_t* out, size_t n) {
    if (!r || (!out && n)) return fail(SSW_ERR_INVALID, "NULL argument");
    if (!r->is_base) return fail(SSW_ERR_STATE, "indices() called on a derived reader (reference: unwrap on None, src/algorithm.rs:507)");
    if (n == 0) return SSW_OK;
    if (n > (size_t)r->w * r->h - 1) return fail(SSW_ERR_INVALID, "more indices requested than AC coefficients");
    ssw_ctx* c = r->ctx;
    CKS(ctx_bind(c));
    CKS(ensure_indices(c, r->d_plane, r->w, r->h, r->cfg.ordering, n, &r->d_idx, &r->k_cached));
    return download_indices(c, r->d_idx, out, n);
}

extern "C" int ssw_reader_destroy(ssw_reader* r) {
    if (!r) return SSW_OK;
    ssw_ctx* c = r->ctx;
    cudaSetDevice(c->device);
    if (r->d_plane) cudaFreeAsync(r->d_plane, c->stream);
    if (r->d_idx) cudaFreeAsync(r->d_idx, c->stream);
    delete r;
    return SSW_OK;
}

// ------------------------------------------------------------------------------------------------
// similarity / bank / marks
// ------------------------------------------------------------------------------------------------
struct ssw_bank {
    CtxRef ctx;
    float* d_marks;
    size_t n_marks, n;
};

// exact: scores in the reference's sequential f32 order (bit-identical to src/algorithm.rs:696-714) -- always for
// Tester::similarity (one mark), on request (SSW_SIM_EXACT=1) for bank searches, which default to the HBM-speed kernel
static int launch_similarity(ssw_ctx* c, const float* d_bank, size_t n_marks, size_t n, const float* d_ext,
                             size_t n_ext, bool pair_mode, float* d_out, bool exact = false) {
    if (n_marks == 0 || n_ext == 0) return SSW_OK;
    if (pair_mode) {
        if (n_marks > 0x7FFFFFFFull || n > 0xFFFFFFFFull) return fail(SSW_ERR_INVALID, "similarity problem too large");
        {
            KScope ks(c, "similarity_pairs");
            launch_pdl(c, similarity_pairs_kernel, (unsigned)((n_marks + kPairsPerCta - 1) / kPairsPerCta), kPairsPerCta * 64, 0, c->stream, 
                d_bank, d_ext, (unsigned)n, (long long)n, (unsigned)n_marks, d_out);
        }
        CK(cudaGetLastError());
        return SSW_OK;
    }
    const size_t gx = (n_marks + kSimMarks - 1) / kSimMarks;
    if (gx > 0x7FFFFFFFull || n_ext > 65535 || n > 0xFFFFFFFFull) return fail(SSW_ERR_INVALID, "similarity problem too large");
    float* d_den = nullptr;
    CK(cudaMallocFromPoolAsync(&d_den, n_ext * sizeof(float), c->pool, c->stream));
    {
        KScope ks(c, "similarity_den");
        launch_pdl(c, similarity_den_kernel, (unsigned)((n_ext + 3) / 4), 128, 0, c->stream, d_ext, (unsigned)n, (long long)n, (unsigned)n_ext, d_den);
    }
    if (!exact && !c->sim_exact && n <= (size_t)kWarpSimMaxN) {
        // default: one warp per stored mark, 16-byte coalesced streaming reads, fixed-shape reduction (HBM speed)
        const size_t ctas = (n_marks + kWarpSimThreads / 32 - 1) / (kWarpSimThreads / 32);
        const unsigned gxw = (unsigned)std::max<size_t>(1, std::min<size_t>(ctas, (size_t)c->sm_count * 6));
        KScope ks(c, "similarity_bank");
        launch_pdl(c, similarity_bank_warp_kernel, dim3(gxw, (unsigned)n_ext), kWarpSimThreads, 0, c->stream,
            d_bank, n_marks, (unsigned)n, d_ext, (long long)n, d_den, d_out, (long long)n_marks);
    } else {
        // SSW_SIM_EXACT=1 (or very long marks): one thread per mark in the reference's sequential order, bit-identical
        KScope ks(c, "similarity_bank_seq");
        launch_pdl(c, similarity_bank_kernel, dim3((unsigned)gx, (unsigned)n_ext), kSimMarks, 0, c->stream, 
            d_bank, n_marks, (unsigned)n, d_ext, (long long)n, d_den, d_out, (long long)n_marks);
    }
    CK(cudaGetLastError());
    CK(cudaFreeAsync(d_den, c->stream));
    return SSW_OK;
}

extern "C" int ssw_similarity(ssw_ctx* c, const float* extracted, const float* mark, size_t n, float* out) {
    if (!c || !out || (n && (!extracted || !mark))) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(ctx_bind(c));
    float* d = nullptr;
    CK(cudaMallocFromPoolAsync(&d, (2 * n + 1) * sizeof(float), c->pool, c->stream));
    if (n) {
        CK(cudaMemcpyAsync(d, extracted, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(d + n, mark, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    CKS(launch_similarity(c, d + n, 1, n, d, 1, false, d + 2 * n, true));
    CK(cudaMemcpyAsync(out, d + 2 * n, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaFreeAsync(d, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SSW_OK;
}

static int fill_normal(ssw_ctx* c, float* d_out, size_t n, uint64_t seed) {
    if (n == 0) return SSW_OK;
    if (seed == 0) {
        std::random_device rd;
        seed = ((uint64_t)rd() << 32) ^ rd();
        if (seed == 0) seed = 0x9E3779B97F4A7C15ull;
    }
    const size_t quads = (n + 3) / 4;
    { KScope ks(c, "normal_fill"); normal_fill_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, c->stream>>>(d_out, n, seed, 0ull); }
    CK(cudaGetLastError());
    return SSW_OK;
}

extern "C" int ssw_bank_create(ssw_ctx* c, const float* marks, size_t n_marks, size_t n, ssw_bank** out) {
    if (!c || !out || (!marks && n_marks * n)) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(ctx_bind(c));
    auto b = std::make_unique<ssw_bank>();
    b->ctx = c; b->n_marks = n_marks; b->n = n; b->d_marks = nullptr;
    CK(cudaMalloc(&b->d_marks, std::max<size_t>(n_marks * n, 1) * sizeof(float)));
    if (n_marks * n) {
        CK(cudaMemcpyAsync(b->d_marks, marks, n_marks * n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    *out = b.release();
    return SSW_OK;
}

extern "C" int ssw_bank_create_normal(ssw_ctx* c, uint64_t seed, size_t n_marks, size_t n, ssw_bank** out) {
    if (!c || !out) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(ctx_bind(c));
    auto b = std::make_unique<ssw_bank>();
    b->ctx = c; b->n_marks = n_marks; b->n = n; b->d_marks = nullptr;
    CK(cudaMalloc(&b->d_marks, std::max<size_t>(n_marks * n, 1) * sizeof(float)));
    CKS(fill_normal(c, b->d_marks, n_marks * n, seed ? seed : 1));
    *out = b.release();
    return SSW_OK;
}

extern "C" int ssw_bank_row(ssw_bank* b, size_t index, float* out) {
    if (!b || !out) return fail(SSW_ERR_INVALID, "NULL argument");
    if (index >= b->n_marks) return fail(SSW_ERR_INVALID, "bank row out of range");
    ssw_ctx* c = b->ctx;
    CKS(ctx_bind(c));
    CK(cudaMemcpyAsync(out, b->d_marks + index * b->n, b->n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SSW_OK;
}

extern "C" int ssw_bank_similarity_dev(ssw_bank* b, const float* d_ext, size_t n_ext, float* d_out) {
    if (!b || !d_ext || !d_out) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(ctx_bind(b->ctx));
    return launch_similarity(b->ctx, b->d_marks, b->n_marks, b->n, d_ext, n_ext, false, d_out);
}

extern "C" int ssw_bank_similarity(ssw_bank* b, const float* ext, size_t n_ext, float* out) {
    if (!b || !ext || !out) return fail(SSW_ERR_INVALID, "NULL argument");
    ssw_ctx* c = b->ctx;
    CKS(ctx_bind(c));
    float *d_ext = nullptr, *d_out = nullptr;
    CK(cudaMallocFromPoolAsync(&d_ext, std::max<size_t>(n_ext * b->n, 1) * sizeof(float), c->pool, c->stream));
    CK(cudaMallocFromPoolAsync(&d_out, std::max<size_t>(n_ext * b->n_marks, 1) * sizeof(float), c->pool, c->stream));
    CK(cudaMemcpyAsync(d_ext, ext, n_ext * b->n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CKS(launch_similarity(c, b->d_marks, b->n_marks, b->n, d_ext, n_ext, false, d_out));
    CK(cudaMemcpyAsync(out, d_out, n_ext * b->n_marks * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaFreeAsync(d_ext, c->stream));
    CK(cudaFreeAsync(d_out, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SSW_OK;
}

extern "C" int ssw_bank_destroy(ssw_bank* b) {
    if (!b) return SSW_OK;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    cudaFree(b->d_marks);
    delete b;
    return SSW_OK;
}

extern "C" int ssw_mark_generate_normal(ssw_ctx* c, uint64_t seed, size_t n, float* out) {
    if (!c || (!out && n)) return fail(SSW_ERR_INVALID, "NULL argument");
    if (n == 0) return SSW_OK;
    CKS(ctx_bind(c));
    float* d = nullptr;
    CK(cudaMallocFromPoolAsync(&d, n * sizeof(float), c->pool, c->stream));
    CKS(fill_normal(c, d, n, seed));
    CK(cudaMemcpyAsync(out, d, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaFreeAsync(d, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SSW_OK;
}

// ------------------------------------------------------------------------------------------------
// fused device-resident pipelines
// ------------------------------------------------------------------------------------------------
static unsigned chunk_images(ssw_ctx* c, size_t np, unsigned batch, int planes_per_image) {
    const size_t per = np * sizeof(float) * planes_per_image;
    size_t n = std::max<size_t>(1, c->chunk_bytes / per);
    return (unsigned)std::min<size_t>(n, batch);
}

static int lowrank_table(ssw_ctx* c, unsigned n, const float** out) {
    auto it = c->lr_tab.find(n);
    if (it == c->lr_tab.end()) {
        float* t = nullptr;
        CK(cudaMalloc(&t, (size_t)kLrTab * n * sizeof(float)));
        { KScope ks(c, "lowrank_table"); lowrank_table_kernel<<<dim3((n + 255) / 256, kLrTab), 256, 0, c->stream>>>(t, n); }
        CK(cudaGetLastError());
        it = c->lr_tab.emplace(n, t).first;
    }
    *out = it->second;
    return SSW_OK;
}

extern "C" int ssw_embed_batch_rgb8_dev(ssw_ctx* c, const uint8_t* rgb, uint32_t w, uint32_t h, uint32_t batch,
                                        const ssw_config* cfg, const float* marks, size_t n, uint8_t* out_rgb) {
    if (!c || !rgb || !out_rgb || (n && !marks)) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_cfg(cfg));
    CKS(check_dims(w, h));
    if (batch == 0) return SSW_OK;
    CKS(ctx_bind(c));
    const size_t np = (size_t)w * h;
    const size_t k = std::min(n, np - 1);
    if (k > (size_t)kTopkCap / 2) return fail(SSW_ERR_UNSUPPORTED, "fused pipeline supports mark lengths up to 4096; use the Writer API");
    // Partial inverse (SSW_PARTIAL_INV, default on; needs the column pipelines): the forward column pass runs out of place and
    // keeps the row-transformed plane R; only the columns that hold a modified coefficient go back through the inverse
    // column pass (C -> R), every other column of R already is the inverse column transform of its unchanged coefficients
    // (up to the round-off of the two passes, ~1e-7 relative: RGB8 within +-1 LSB of the full inverse).
    const bool partial = c->partial_inv && !c->lowrank && k > 0 && c->use_fast && pipe_col_available(c, (int)w, (int)h) && !c->force_line1 &&
                         pipe_row_available(c, rgb, out_rgb, (int)w, (int)h);
    const unsigned cb = chunk_images(c, np, batch, partial ? 2 : 1);
    CKS(ensure_topk_scratch(c, cb));
    float* d_planes = nullptr;
    unsigned* d_idx = nullptr;
    float* d_delta = nullptr;
    CK(cudaMallocFromPoolAsync(&d_planes, (size_t)cb * np * sizeof(float) * (partial ? 2 : 1), c->pool, c->stream));
    float* d_rows = partial ? d_planes + (size_t)cb * np : nullptr;   // R: row-transformed frames (partial inverse)
    CK(cudaMallocFromPoolAsync(&d_idx, (size_t)cb * std::max<size_t>(k, 1) * sizeof(unsigned), c->pool, c->stream));
    CK(cudaMallocFromPoolAsync(&d_delta, (size_t)cb * std::max<size_t>(k, 1) * sizeof(float), c->pool, c->stream));
    int rc = SSW_OK;
    for (unsigned b0 = 0; b0 < batch && rc == SSW_OK; b0 += cb) {
        const unsigned nb = std::min(cb, batch - b0);
        const uint8_t* src = rgb + (size_t)b0 * np * 3;
        // the forward column pipeline leaves the low-frequency-block histogram of every frame for the ordering
        // (one frame per launch: the in-pipeline histogram takes a kernel off the latency chain; on batched launches it
        // was measured slower than the separate topk_block_bin kernel, whose cost is shared by the whole batch)
        c->col_hist.want = k > 0 && !c->topk_full_hist && c->col_hist_on && nb <= 4; c->col_hist.done = c->col_hist.collected = false;
        c->col_hist.k = (unsigned)k; c->col_hist.ordering = cfg->ordering;
        c->tile_max.want = k > 0 && cfg->ordering == 0 && !c->topk_full_hist; c->tile_max.tiles = 0;
        if (partial) {
            rc = run_rows_forward(c, PIX_RGB8, src, w, h, nb, d_rows, 1.f, 1.f);
            bool done = false;
            if (rc == SSW_OK) rc = pipe_col(c, false, w, h, nb, d_rows, 1.f, 1.f, &done, d_planes);
            if (rc == SSW_OK && !done) rc = fail(SSW_ERR_STATE, "partial inverse: no column pipeline for this frame");
        } else {
            rc = run_forward(c, PIX_RGB8, src, w, h, nb, d_planes, SSW_DCT2);
        }
        const int hist_ready = c->col_hist.done ? (c->col_hist.collected ? 2 : 1) : 0;
        c->col_hist.want = false; c->tile_max.want = false;
        const bool lowrank = c->lowrank && k > 0 && (w % 4u) == 0 && w <= 65535u && h <= 65535u && aligned(src, 4) && aligned(out_rgb, 4) &&
                             ((np * 3) % 4) == 0;
        if (rc == SSW_OK && k) {
            // ordering + embedding: the ranking kernel applies mark value r to the rank-r coefficient in place -- or, for
            // the low-rank inverse, stores the change D_r = f(c, w_r) - c of that coefficient (lowrank.cuh)
            TopkApply ap;
            std::memset(&ap, 0, sizeof(ap));
            ap.mode = lowrank ? 3 : 1; ap.method = cfg->method; ap.alpha = cfg->alpha; ap.width = (lowrank || partial) ? w : 0u;
            ap.planes = d_planes; ap.plane_stride = (long long)np;
            ap.marks = marks + (size_t)b0 * n; ap.mark_stride = (long long)n;
            ap.out = d_delta; ap.out_stride = (long long)k;
            rc = run_topk_fast(c, d_planes, w, h, nb, cfg->ordering, (unsigned)k, d_idx, (long long)k, c->topk_full_hist, hist_ready, &ap);
        }
        if (rc == SSW_OK && lowrank) {
            // Y' = Y + IDCT(D): Kr x W strip products, then Kr FMAs per pixel on top of the original frame
            const size_t ent_bytes = k * sizeof(unsigned long long);
            const float *cx_tab, *cy_tab;
            CKS(lowrank_table(c, w, &cx_tab));
            CKS(lowrank_table(c, h, &cy_tab));
            const void* key = (const void*)lowrank_rows_kernel;
            auto it = c->smem_attr.find(key);
            if (it == c->smem_attr.end() || it->second < (int)ent_bytes) {
                CK(cudaFuncSetAttribute(lowrank_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(ent_bytes, 16384)));
                c->smem_attr[key] = (int)std::max<size_t>(ent_bytes, 16384);
            }
            for (unsigned i0 = 0; i0 < nb && rc == SSW_OK; i0 += 65535) {
                const unsigned ni = std::min(65535u, nb - i0);
                {
                    KScope ks(c, "lowrank_rows");
                    launch_pdl(c, lowrank_rows_kernel, dim3((w + 31) / 32, ni), 256, ent_bytes, c->stream, d_planes + (size_t)i0 * np, (long long)np, w, h,
                               (const unsigned*)(d_idx + (size_t)i0 * k), (const float*)(d_delta + (size_t)i0 * k), (unsigned)k,
                               (const unsigned*)(c->ts.maxrow + i0), cx_tab);
                }
                {
                    KScope ks(c, "lowrank_apply");
                    launch_pdl(c, c->lowrank_mma ? lowrank_apply_mma_kernel : lowrank_apply_kernel,
                               dim3((w + kLrTileX - 1) / kLrTileX, (h + kLrTileY - 1) / kLrTileY, ni), 256, 0, c->stream,
                               (const float*)(d_planes + (size_t)i0 * np), (long long)np, w, h, (const unsigned char*)(src + (size_t)i0 * np * 3),
                               (unsigned char*)(out_rgb + ((size_t)b0 + i0) * np * 3), (const unsigned*)(c->ts.maxrow + i0), cy_tab, -0.0f);
                }
                CK(cudaGetLastError());
            }
        } else if (rc == SSW_OK && partial) {
            bool done = false;
            rc = pipe_col(c, true, w, h, nb, d_planes, 1.f, 1.f, &done, d_rows, c->ts.maxcol);
            if (rc == SSW_OK && !done) rc = fail(SSW_ERR_STATE, "partial inverse: no column pipeline for this frame");
            if (rc == SSW_OK) {
                // the columns that skipped the column passes lack their gain h/2; the row pipeline applies it as it reads them
                c->row_cut.img = c->ts.maxcol + 1;
                c->row_cut.shift = (c->col_pipe == 3 && h == 2160) ? 2 : 3;   // log2(columns per tile of the column pipeline that ran)
                c->row_cut.gain = 0.5f * (float)h;
                c->row_cut.used = false;
                rc = run_rows_inverse(c, d_rows, PIX_RGB8, src, w, h, nb, PIX_RGB8, out_rgb + (size_t)b0 * np * 3, 4.0f / (float)np);
                const bool used = c->row_cut.used;
                c->row_cut.img = nullptr;
                if (rc == SSW_OK && !used) rc = fail(SSW_ERR_STATE, "partial inverse: the row pass did not take the pipeline");
            }
        } else if (rc == SSW_OK) {
            rc = run_inverse(c, d_planes, PIX_RGB8, src, w, h, nb, PIX_RGB8, out_rgb + (size_t)b0 * np * 3);
        }
    }
    cudaFreeAsync(d_delta, c->stream);
    cudaFreeAsync(d_planes, c->stream);
    cudaFreeAsync(d_idx, c->stream);
    if (rc == SSW_OK) CK(cudaGetLastError());
    return rc;
}

extern "C" int ssw_extract_batch_rgb8_dev(ssw_ctx* c, const uint8_t* base_rgb, const uint8_t* derived_rgb, uint32_t w,
                                          uint32_t h, uint32_t batch, const ssw_config* cfg, size_t n,
                                          float* extracted, const float* marks, float* sim) {
    if (!c || !base_rgb || !derived_rgb || !extracted) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_cfg(cfg));
    CKS(check_dims(w, h));
    if (sim && !marks) return fail(SSW_ERR_INVALID, "similarity requested without marks");
    const size_t np = (size_t)w * h;
    if (n >= np) return fail(SSW_ERR_INVALID, "Desired extraction length exceeds available coefficients.");
    if (n > (size_t)kTopkCap / 2) return fail(SSW_ERR_UNSUPPORTED, "fused pipeline supports mark lengths up to 4096; use the Reader API");
    if (batch == 0 || n == 0) return SSW_OK;
    CKS(ctx_bind(c));
    const unsigned cb = chunk_images(c, np, batch, 2);
    CKS(ensure_topk_scratch(c, cb));
    float* d_planes = nullptr;
    unsigned* d_idx = nullptr;
    CK(cudaMallocFromPoolAsync(&d_planes, (size_t)cb * np * 2 * sizeof(float), c->pool, c->stream));
    CK(cudaMallocFromPoolAsync(&d_idx, (size_t)cb * n * sizeof(unsigned), c->pool, c->stream));
    int rc = SSW_OK;
    for (unsigned b0 = 0; b0 < batch && rc == SSW_OK; b0 += cb) {
        const unsigned nb = std::min(cb, batch - b0);
        float* pb = d_planes;
        float* pd = d_planes + (size_t)cb * np;
        // ordering + extraction [+ 1:1 score]: the ranking kernel reads the rank-r coefficient pair and stores x*_r
        TopkApply ap;
        std::memset(&ap, 0, sizeof(ap));
        ap.mode = 2; ap.method = cfg->method; ap.alpha = cfg->alpha;
        ap.planes = pb; ap.derived = pd; ap.plane_stride = (long long)np;
        ap.marks = marks ? marks + (size_t)b0 * n : nullptr; ap.mark_stride = (long long)n;
        ap.out = extracted + (size_t)b0 * n; ap.out_stride = (long long)n;
        ap.sim = (sim && !c->sim_exact) ? sim + b0 : nullptr;
        auto base_forward = [&]() -> int {
            c->col_hist.want = !c->topk_full_hist && c->col_hist_on && nb <= 4; c->col_hist.done = c->col_hist.collected = false;
            c->col_hist.k = (unsigned)n; c->col_hist.ordering = cfg->ordering;
            c->tile_max.want = cfg->ordering == 0 && !c->topk_full_hist; c->tile_max.tiles = 0;
            const int r = run_forward(c, PIX_RGB8, base_rgb + (size_t)b0 * np * 3, w, h, nb, pb, SSW_DCT2);
            c->col_hist.want = false; c->tile_max.want = false;
            return r;
        };
        int hist_ready = 0;
        if (c->overlap_topk && !c->profiling) {   // per-kernel profiling times every kernel alone, on one stream
            // fork: the derived frame's forward transform runs on the side stream beside the base frame's forward
            // transform and the candidate collection; CTAs of the two transforms fill each other's partial waves.
            // The join sits before the ranking kernel, the first consumer of the derived planes.
            cudaStream_t main_stream = c->stream;
            CK(cudaEventRecord(c->ev_fork, main_stream));
            CK(cudaStreamWaitEvent(c->aux, c->ev_fork, 0));
            c->stream = c->aux;
            rc = run_forward(c, PIX_RGB8, derived_rgb + (size_t)b0 * np * 3, w, h, nb, pd, SSW_DCT2);
            c->stream = main_stream;
            CK(cudaEventRecord(c->ev_join, c->aux));
            if (rc == SSW_OK) { rc = base_forward(); hist_ready = c->col_hist.done ? (c->col_hist.collected ? 2 : 1) : 0; }
            if (rc == SSW_OK) rc = run_topk_fast(c, pb, w, h, nb, cfg->ordering, (unsigned)n, d_idx, (long long)n, c->topk_full_hist, hist_ready, &ap, c->ev_join);
            else CK(cudaStreamWaitEvent(main_stream, c->ev_join, 0));   // error paths: keep the streams ordered
        } else {
            rc = base_forward(); hist_ready = c->col_hist.done ? (c->col_hist.collected ? 2 : 1) : 0;
            if (rc == SSW_OK) rc = run_forward(c, PIX_RGB8, derived_rgb + (size_t)b0 * np * 3, w, h, nb, pd, SSW_DCT2);
            if (rc == SSW_OK) rc = run_topk_fast(c, pb, w, h, nb, cfg->ordering, (unsigned)n, d_idx, (long long)n, c->topk_full_hist, hist_ready, &ap);
        }
        if (rc == SSW_OK && sim && c->sim_exact)
            rc = launch_similarity(c, marks + (size_t)b0 * n, nb, n, extracted + (size_t)b0 * n, nb, true, sim + b0);
    }
    cudaFreeAsync(d_planes, c->stream);
    cudaFreeAsync(d_idx, c->stream);
    if (rc == SSW_OK) CK(cudaGetLastError());
    return rc;
}

// host-buffer entry points: the batch is cut into chunks of >= ~32 MB; chunk i+1 is uploaded (copy-in stream) while
// chunk i is transformed (context stream) and chunk i-1 is downloaded (copy-out stream) -- PCIe is full duplex, so a
// batch approaches the one-directional line rate instead of the sum of both directions.
static int pipe_event(ssw_ctx* c, size_t i, cudaEvent_t* ev) {
    while (c->pipe_events.size() <= i) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->pipe_events.push_back(e);
    }
    *ev = c->pipe_events[i];
    return SSW_OK;
}

static uint32_t pipe_chunk_frames(size_t frame_bytes, uint32_t batch) {
    if (batch <= 1) return std::max(1u, batch);
    const size_t target = 32u << 20;
    uint32_t cb = (uint32_t)std::max<size_t>(1, target / std::max<size_t>(1, frame_bytes));
    cb = std::min(cb, (batch + 1) / 2);   // at least two chunks so that the directions overlap
    return std::max(1u, cb);
}

extern "C" int ssw_embed_batch_rgb8(ssw_ctx* c, const uint8_t* rgb, uint32_t w, uint32_t h, uint32_t batch,
                                    const ssw_config* cfg, const float* marks, size_t n, uint8_t* out_rgb) {
    if (!c || !rgb || !out_rgb || (n && !marks)) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_dims(w, h));
    if (batch == 0) return SSW_OK;
    CKS(ctx_bind(c));
    const size_t fbytes = (size_t)w * h * 3, bytes = fbytes * batch;
    uint8_t *d_in = nullptr, *d_out = nullptr;
    float* d_marks = nullptr;
    CK(cudaMallocFromPoolAsync(&d_in, bytes, c->pool, c->stream));
    CK(cudaMallocFromPoolAsync(&d_out, bytes, c->pool, c->stream));
    CK(cudaMallocFromPoolAsync(&d_marks, std::max<size_t>(n * batch, 1) * sizeof(float), c->pool, c->stream));
    if (n) CK(cudaMemcpyAsync(d_marks, marks, n * batch * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    const uint32_t cb = pipe_chunk_frames(fbytes, batch);
    const uint32_t nchunks = (batch + cb - 1) / cb;
    // first attempt: threshold bin from the low-frequency block; a candidate overflow (noise-like
    // spectrum) is repaired by one more run with the full-plane histogram
    int rc = SSW_OK;
    unsigned ov = 0;
    const bool saved = c->topk_full_hist;
    for (int attempt = 0; attempt < 2; ++attempt) {
        cudaEvent_t e0;
        CKS(pipe_event(c, 0, &e0));
        CK(cudaEventRecord(e0, c->stream));                 // allocations (and a previous attempt) are ordered before the copies
        CK(cudaStreamWaitEvent(c->copy_in, e0, 0));
        CK(cudaStreamWaitEvent(c->copy_out, e0, 0));
        for (uint32_t ci = 0; ci < nchunks && rc == SSW_OK; ++ci) {
            const uint32_t f0 = ci * cb, nf = std::min(cb, batch - f0);
            cudaEvent_t e_in, e_cmp;
            CKS(pipe_event(c, 1 + 2 * (size_t)ci, &e_in));
            CKS(pipe_event(c, 2 + 2 * (size_t)ci, &e_cmp));
            CK(cudaMemcpyAsync(d_in + f0 * fbytes, rgb + f0 * fbytes, nf * fbytes, cudaMemcpyHostToDevice, c->copy_in));
            CK(cudaEventRecord(e_in, c->copy_in));
            CK(cudaStreamWaitEvent(c->stream, e_in, 0));
            rc = ssw_embed_batch_rgb8_dev(c, d_in + f0 * fbytes, w, h, nf, cfg, d_marks + (size_t)f0 * n, n, d_out + f0 * fbytes);
            if (rc != SSW_OK) break;
            CK(cudaEventRecord(e_cmp, c->stream));
            CK(cudaStreamWaitEvent(c->copy_out, e_cmp, 0));
            CK(cudaMemcpyAsync(out_rgb + f0 * fbytes, d_out + f0 * fbytes, nf * fbytes, cudaMemcpyDeviceToHost, c->copy_out));
        }
        CK(cudaEventRecord(e0, c->copy_out));               // join the copy streams back into the context stream
        CK(cudaStreamWaitEvent(c->stream, e0, 0));
        CK(cudaEventRecord(e0, c->copy_in));
        CK(cudaStreamWaitEvent(c->stream, e0, 0));
        int rs = take_overflow(c, &ov);  // synchronises
        if (rc == SSW_OK) rc = rs;
        if (rc != SSW_OK || !ov || c->topk_full_hist) break;
        c->topk_full_hist = true;
    }
    c->topk_full_hist = saved;
    cudaFreeAsync(d_in, c->stream); cudaFreeAsync(d_out, c->stream); cudaFreeAsync(d_marks, c->stream);
    c->last_fallbacks = (int)ov;
    if (rc == SSW_OK && ov) rc = fail(SSW_ERR_UNSUPPORTED, "top-k candidate overflow in the fused pipeline (degenerate spectrum): use the Writer API for these frames");
    return rc;
}

extern "C" int ssw_extract_batch_rgb8(ssw_ctx* c, const uint8_t* base_rgb, const uint8_t* derived_rgb, uint32_t w,
                                      uint32_t h, uint32_t batch, const ssw_config* cfg, size_t n, float* extracted,
                                      const float* marks, float* sim) {
    if (!c || !base_rgb || !derived_rgb || !extracted) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_dims(w, h));
    if (batch == 0 || n == 0) return SSW_OK;
    CKS(ctx_bind(c));
    const size_t fbytes = (size_t)w * h * 3, bytes = fbytes * batch;
    uint8_t *d_b = nullptr, *d_d = nullptr;
    float *d_ext = nullptr, *d_marks = nullptr, *d_sim = nullptr;
    CK(cudaMallocFromPoolAsync(&d_b, bytes, c->pool, c->stream));
    CK(cudaMallocFromPoolAsync(&d_d, bytes, c->pool, c->stream));
    CK(cudaMallocFromPoolAsync(&d_ext, n * batch * sizeof(float), c->pool, c->stream));
    if (sim && marks) {
        CK(cudaMallocFromPoolAsync(&d_marks, n * batch * sizeof(float), c->pool, c->stream));
        CK(cudaMallocFromPoolAsync(&d_sim, batch * sizeof(float), c->pool, c->stream));
        CK(cudaMemcpyAsync(d_marks, marks, n * batch * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    const uint32_t cb = pipe_chunk_frames(2 * fbytes, batch);
    const uint32_t nchunks = (batch + cb - 1) / cb;
    int rc = SSW_OK;
    unsigned ov = 0;
    const bool saved = c->topk_full_hist;
    for (int attempt = 0; attempt < 2; ++attempt) {
        cudaEvent_t e0;
        CKS(pipe_event(c, 0, &e0));
        CK(cudaEventRecord(e0, c->stream));
        CK(cudaStreamWaitEvent(c->copy_in, e0, 0));
        for (uint32_t ci = 0; ci < nchunks && rc == SSW_OK; ++ci) {
            const uint32_t f0 = ci * cb, nf = std::min(cb, batch - f0);
            cudaEvent_t e_in;
            CKS(pipe_event(c, 1 + (size_t)ci, &e_in));
            CK(cudaMemcpyAsync(d_b + f0 * fbytes, base_rgb + f0 * fbytes, nf * fbytes, cudaMemcpyHostToDevice, c->copy_in));
            CK(cudaMemcpyAsync(d_d + f0 * fbytes, derived_rgb + f0 * fbytes, nf * fbytes, cudaMemcpyHostToDevice, c->copy_in));
            CK(cudaEventRecord(e_in, c->copy_in));
            CK(cudaStreamWaitEvent(c->stream, e_in, 0));
            rc = ssw_extract_batch_rgb8_dev(c, d_b + f0 * fbytes, d_d + f0 * fbytes, w, h, nf, cfg, n, d_ext + (size_t)f0 * n,
                                            d_marks ? d_marks + (size_t)f0 * n : nullptr, d_sim ? d_sim + f0 : nullptr);
        }
        if (rc == SSW_OK) {
            cudaError_t e = cudaMemcpyAsync(extracted, d_ext, n * batch * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess && d_sim) e = cudaMemcpyAsync(sim, d_sim, batch * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
            if (e != cudaSuccess) rc = fail(SSW_ERR_CUDA, cudaGetErrorString(e));
        }
        CK(cudaEventRecord(e0, c->copy_in));
        CK(cudaStreamWaitEvent(c->stream, e0, 0));
        int rs = take_overflow(c, &ov);  // synchronises
        if (rc == SSW_OK) rc = rs;
        if (rc != SSW_OK || !ov || c->topk_full_hist) break;
        c->topk_full_hist = true;
    }
    c->topk_full_hist = saved;
    cudaFreeAsync(d_b, c->stream); cudaFreeAsync(d_d, c->stream); cudaFreeAsync(d_ext, c->stream);
    if (d_marks) cudaFreeAsync(d_marks, c->stream);
    if (d_sim) cudaFreeAsync(d_sim, c->stream);
    c->last_fallbacks = (int)ov;
    if (rc == SSW_OK && ov) rc = fail(SSW_ERR_UNSUPPORTED, "top-k candidate overflow in the fused pipeline (degenerate spectrum): use the Reader API for these frames");
    return rc;
}

// ------------------------------------------------------------------------------------------------
// asynchronous host-buffer entry points: enqueue and return; results are valid after ssw_ctx_synchronize.
// Consecutive calls overlap -- the upload of call i+1 runs on the copy-in stream beside the kernels (context stream)
// and the download (copy-out stream) of call i; PCIe is full duplex, so a stream of frames moves at the line rate of
// the busier direction instead of the sum of both.  Host buffers may be reused freely between
// calls: a copy that touches a host range with an earlier copy still in flight is ordered behind that copy (the
// watermarked frames of an embed call can be handed to an extract call without a synchronize in between).
// Like the _dev entry points they do not repair a top-k candidate overflow (no host synchronisation): such frames come
// back unmarked / with a zero vector and are reported by ssw_ctx_last_topk_fallbacks.
// ------------------------------------------------------------------------------------------------
static bool ranges_overlap(const void* a, size_t na, const char* b, size_t nb) {
    const char* pa = (const char*)a;
    return na && nb && pa < b + nb && b < pa + na;
}

static int range_event(ssw_ctx* c, cudaEvent_t* ev) {
    constexpr size_t kRing = 32;
    if (c->range_events.size() < kRing) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->range_events.push_back(e);
        *ev = e;
        return SSW_OK;
    }
    cudaEvent_t e = c->range_events[c->range_next++ % kRing];
    CK(cudaEventSynchronize(e));   // 16 calls old: long complete; its ranges leave the pending lists
    auto drop = [&](std::vector<ssw_ctx::HostRange>& v) {
        v.erase(std::remove_if(v.begin(), v.end(), [&](const ssw_ctx::HostRange& r) { return r.ev == e; }), v.end());
    };
    drop(c->pending_d2h); drop(c->pending_h2d);
    *ev = e;
    return SSW_OK;
}

// the next staging set, large enough, with every stream that will touch it ordered behind its previous user
static int acquire_stage(ssw_ctx* c, cudaStream_t up, size_t in_bytes, size_t out_bytes, size_t n_f32, ssw_ctx::Staging** out) {
    ssw_ctx::Staging& st = c->stage[c->stage_next++ & 3];
    if (!st.done) CK(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
    if (st.in_cap < in_bytes || st.out_cap < out_bytes || st.f32_cap < n_f32) {
        CKS(ssw_ctx_synchronize(c));   // growth is rare (first calls): everything drains, then the set is reallocated
        if (st.in_cap < in_bytes) { cudaFree(st.in); st.in = nullptr; st.in_cap = 0; CK(cudaMalloc(&st.in, in_bytes)); st.in_cap = in_bytes; }
        if (st.out_cap < out_bytes) { cudaFree(st.out); st.out = nullptr; st.out_cap = 0; CK(cudaMalloc(&st.out, out_bytes)); st.out_cap = out_bytes; }
        if (st.f32_cap < n_f32) { cudaFree(st.f32); st.f32 = nullptr; st.f32_cap = 0; CK(cudaMalloc(&st.f32, n_f32 * sizeof(float))); st.f32_cap = n_f32; }
        st.used = false;
    }
    if (st.used) {
        CK(cudaStreamWaitEvent(up, st.done, 0));
        CK(cudaStreamWaitEvent(c->stream, st.done, 0));
    }
    st.used = true;
    *out = &st;
    return SSW_OK;
}

// order `stream` behind every pending copy of the other direction that touches [p, p+n)
static int wait_host_range(ssw_ctx* c, cudaStream_t stream, const std::vector<ssw_ctx::HostRange>& pending, const void* p, size_t n) {
    for (const auto& r : pending)
        if (ranges_overlap(p, n, r.p, r.n)) CK(cudaStreamWaitEvent(stream, r.ev, 0));
    return SSW_OK;
}

extern "C" int ssw_embed_batch_rgb8_async(ssw_ctx* c, const uint8_t* rgb, uint32_t w, uint32_t h, uint32_t batch,
                                          const ssw_config* cfg, const float* marks, size_t n, uint8_t* out_rgb) {
    if (!c || !rgb || !out_rgb || (n && !marks)) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_cfg(cfg));
    CKS(check_dims(w, h));
    if (batch == 0) return SSW_OK;
    CKS(ctx_bind(c));
    const size_t fbytes = (size_t)w * h * 3, bytes = fbytes * batch;
    cudaStream_t up = (c->async_calls++ & 1) ? c->copy_in2 : c->copy_in;   // consecutive calls alternate upload streams
    ssw_ctx::Staging* st;
    CKS(acquire_stage(c, up, bytes, bytes, std::max<size_t>(n * batch, 1), &st));
    cudaEvent_t ev_up, ev_down;
    CKS(range_event(c, &ev_up));
    CKS(range_event(c, &ev_down));
    // uploads wait for downloads still writing their source ranges; downloads for uploads still reading their target
    CKS(wait_host_range(c, up, c->pending_d2h, rgb, bytes));
    CKS(wait_host_range(c, up, c->pending_d2h, marks, n * batch * sizeof(float)));
    CKS(wait_host_range(c, c->copy_out, c->pending_h2d, out_rgb, bytes));
    CKS(wait_host_range(c, c->copy_out, c->pending_d2h, out_rgb, bytes));
    if (n) CK(cudaMemcpyAsync(st->f32, marks, n * batch * sizeof(float), cudaMemcpyHostToDevice, up));
    const uint32_t cb = pipe_chunk_frames(fbytes, batch);
    const uint32_t nchunks = (batch + cb - 1) / cb;
    int rc = SSW_OK;
    for (uint32_t ci = 0; ci < nchunks && rc == SSW_OK; ++ci) {
        const uint32_t f0 = ci * cb, nf = std::min(cb, batch - f0);
        cudaEvent_t e_in, e_cmp;
        CKS(pipe_event(c, 1 + 2 * (size_t)ci, &e_in));
        CKS(pipe_event(c, 2 + 2 * (size_t)ci, &e_cmp));
        CK(cudaMemcpyAsync(st->in + f0 * fbytes, rgb + f0 * fbytes, nf * fbytes, cudaMemcpyHostToDevice, up));
        CK(cudaEventRecord(e_in, up));
        CK(cudaStreamWaitEvent(c->stream, e_in, 0));
        rc = ssw_embed_batch_rgb8_dev(c, st->in + f0 * fbytes, w, h, nf, cfg, st->f32 + (size_t)f0 * n, n, st->out + f0 * fbytes);
        if (rc != SSW_OK) break;
        CK(cudaEventRecord(e_cmp, c->stream));
        CK(cudaStreamWaitEvent(c->copy_out, e_cmp, 0));
        CK(cudaMemcpyAsync(out_rgb + f0 * fbytes, st->out + f0 * fbytes, nf * fbytes, cudaMemcpyDeviceToHost, c->copy_out));
    }
    CK(cudaEventRecord(ev_up, up));
    CK(cudaEventRecord(ev_down, c->copy_out));
    CK(cudaEventRecord(st->done, c->copy_out));
    c->pending_h2d.push_back({(const char*)rgb, bytes, ev_up});
    if (n) c->pending_h2d.push_back({(const char*)marks, n * batch * sizeof(float), ev_up});
    c->pending_d2h.push_back({(const char*)out_rgb, bytes, ev_down});
    return rc;
}

extern "C" int ssw_extract_batch_rgb8_async(ssw_ctx* c, const uint8_t* base_rgb, const uint8_t* derived_rgb, uint32_t w,
                                            uint32_t h, uint32_t batch, const ssw_config* cfg, size_t n, float* extracted,
                                            const float* marks, float* sim) {
    if (!c || !base_rgb || !derived_rgb || !extracted) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_cfg(cfg));
    CKS(check_dims(w, h));
    if (sim && !marks) return fail(SSW_ERR_INVALID, "similarity requested without marks");
    if (batch == 0 || n == 0) return SSW_OK;
    CKS(ctx_bind(c));
    const size_t fbytes = (size_t)w * h * 3, bytes = fbytes * batch, vbytes = n * batch * sizeof(float);
    cudaStream_t up = (c->async_calls++ & 1) ? c->copy_in2 : c->copy_in;
    ssw_ctx::Staging* st;
    // out: extracted vectors followed by the scores; f32: the marks
    CKS(acquire_stage(c, up, 2 * bytes, vbytes + batch * sizeof(float), n * batch, &st));
    uint8_t *d_b = st->in, *d_d = st->in + bytes;
    float* d_ext = (float*)st->out;
    float* d_sim = sim ? d_ext + n * batch : nullptr;
    cudaEvent_t ev_up, ev_down;
    CKS(range_event(c, &ev_up));
    CKS(range_event(c, &ev_down));
    CKS(wait_host_range(c, up, c->pending_d2h, base_rgb, bytes));
    if (marks) CKS(wait_host_range(c, up, c->pending_d2h, marks, vbytes));
    for (const void* q : {(const void*)extracted, (const void*)sim}) {
        if (!q) continue;
        const size_t qn = q == (const void*)extracted ? vbytes : batch * sizeof(float);
        CKS(wait_host_range(c, c->copy_out, c->pending_h2d, q, qn));
        CKS(wait_host_range(c, c->copy_out, c->pending_d2h, q, qn));
    }
    if (marks) CK(cudaMemcpyAsync(st->f32, marks, vbytes, cudaMemcpyHostToDevice, up));
    const uint32_t cb = pipe_chunk_frames(2 * fbytes, batch);
    const uint32_t nchunks = (batch + cb - 1) / cb;
    int rc = SSW_OK;
    for (uint32_t ci = 0; ci < nchunks && rc == SSW_OK; ++ci) {
        const uint32_t f0 = ci * cb, nf = std::min(cb, batch - f0);
        cudaEvent_t e_in;
        CKS(pipe_event(c, 1 + (size_t)ci, &e_in));
        CK(cudaMemcpyAsync(d_b + f0 * fbytes, base_rgb + f0 * fbytes, nf * fbytes, cudaMemcpyHostToDevice, up));
        // the derived frames are typically the output of an embed call whose download is still in flight: the base
        // frames above go up beside that download, the derived frames behind it
        if (ci == 0) CKS(wait_host_range(c, up, c->pending_d2h, derived_rgb, bytes));
        CK(cudaMemcpyAsync(d_d + f0 * fbytes, derived_rgb + f0 * fbytes, nf * fbytes, cudaMemcpyHostToDevice, up));
        CK(cudaEventRecord(e_in, up));
        CK(cudaStreamWaitEvent(c->stream, e_in, 0));
        rc = ssw_extract_batch_rgb8_dev(c, d_b + f0 * fbytes, d_d + f0 * fbytes, w, h, nf, cfg, n, d_ext + (size_t)f0 * n,
                                        marks ? st->f32 + (size_t)f0 * n : nullptr, d_sim ? d_sim + f0 : nullptr);
    }
    if (rc == SSW_OK) {
        cudaEvent_t e_cmp;
        CKS(pipe_event(c, 0, &e_cmp));
        CK(cudaEventRecord(e_cmp, c->stream));
        CK(cudaStreamWaitEvent(c->copy_out, e_cmp, 0));
        CK(cudaMemcpyAsync(extracted, d_ext, vbytes, cudaMemcpyDeviceToHost, c->copy_out));
        if (d_sim) CK(cudaMemcpyAsync(sim, d_sim, batch * sizeof(float), cudaMemcpyDeviceToHost, c->copy_out));
    }
    CK(cudaEventRecord(ev_up, up));
    CK(cudaEventRecord(ev_down, c->copy_out));
    CK(cudaEventRecord(st->done, c->copy_out));
    c->pending_h2d.push_back({(const char*)base_rgb, bytes, ev_up});
    c->pending_h2d.push_back({(const char*)derived_rgb, bytes, ev_up});
    if (marks) c->pending_h2d.push_back({(const char*)marks, vbytes, ev_up});
    c->pending_d2h.push_back({(const char*)extracted, vbytes, ev_down});
    if (sim) c->pending_d2h.push_back({(const char*)sim, batch * sizeof(float), ev_down});
    return rc;
}

// completion markers of the asynchronous calls: ssw_ctx_marker returns a ticket for "everything enqueued so far";
// ssw_ctx_wait_marker blocks the host until that point is reached (results of the calls before the marker are in host
// memory) WITHOUT draining later calls -- a caller keeps one or two calls in flight and reads the results of the previous
// one.  Tickets are valid for the 64 most recent markers.
extern "C" int ssw_ctx_marker(ssw_ctx* c, uint64_t* marker) {
    if (!c || !marker) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(ctx_bind(c));
    constexpr size_t kRing = 64;
    if (c->markers.size() < kRing) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));
        c->markers.push_back(e);
    }
    cudaEvent_t e = c->markers[c->marker_next % kRing];
    // downloads are the last stage of every asynchronous call; calls without a download end on the context stream
    CK(cudaEventRecord(c->ev_join, c->stream));
    CK(cudaStreamWaitEvent(c->copy_out, c->ev_join, 0));
    CK(cudaEventRecord(e, c->copy_out));
    *marker = c->marker_next++;
    return SSW_OK;
}

extern "C" int ssw_ctx_wait_marker(ssw_ctx* c, uint64_t marker) {
    if (!c) return fail(SSW_ERR_INVALID, "ctx is NULL");
    if (marker >= c->marker_next || marker + 64 < c->marker_next) return fail(SSW_ERR_INVALID, "unknown or expired marker");
    CKS(ctx_bind(c));
    CK(cudaEventSynchronize(c->markers[marker % 64]));
    return SSW_OK;
}

extern "C" int ssw_ctx_last_topk_fallbacks(ssw_ctx* c) {
    if (!c) return 0;
    if (ctx_bind(c) != SSW_OK) return -1;
    unsigned ov = 0;
    if (take_overflow(c, &ov) != SSW_OK) return -1;
    const int r = c->last_fallbacks + (int)ov;
    c->last_fallbacks = 0;
    return r;
}

// ------------------------------------------------------------------------------------------------
// device self-test of the packed RGB8 output conversion (pack_u8x4x2, dct_fast.cuh) against the reference-faithful
// round(clamp(v, 0, 1) * 255) (color.cuh) over ALL 2^32 float bit patterns; returns the number of mismatching values
// ------------------------------------------------------------------------------------------------
__global__ void selftest_pack_u8_kernel(unsigned long long* bad, float nz) {
#if defined(__CUDA_ARCH__)   // the packed helpers exist in the device pass only
    unsigned long long local = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x * 8ull;
    for (unsigned long long b = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 8ull; b < (1ull << 32); b += stride) {
        float2 o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = make_float2(__uint_as_float((unsigned)(b + 2 * i)), __uint_as_float((unsigned)(b + 2 * i + 1)));
        unsigned wa, wb;
        fast::pack_u8x4x2(o, nz, wa, wb);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            local += ((wa >> (8 * i)) & 255u) != unit_to_u8(clamp01(o[i].x));
            local += ((wb >> (8 * i)) & 255u) != unit_to_u8(clamp01(o[i].y));
        }
    }
    if (local) atomicAdd(bad, local);
#endif
}

extern "C" int ssw_selftest_pack_u8(ssw_ctx* c, uint64_t* mismatches) {
    if (!c || !mismatches) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(ctx_bind(c));
    unsigned long long* d = nullptr;
    CK(cudaMallocFromPoolAsync(&d, sizeof(unsigned long long), c->pool, c->stream));
    CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long), c->stream));
    { KScope ks(c, "selftest_pack_u8"); selftest_pack_u8_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(d, -0.0f); }
    CK(cudaGetLastError());
    unsigned long long hv = 0;
    CK(cudaMemcpyAsync(&hv, d, sizeof(hv), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaFreeAsync(d, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *mismatches = hv;
    return SSW_OK;
}

// ------------------------------------------------------------------------------------------------
// synthetic frames + stage hooks
// ------------------------------------------------------------------------------------------------
extern "C" int ssw_synth_frame_rgb8_dev(ssw_ctx* c, uint32_t w, uint32_t h, uint64_t seed, uint32_t first_image,
                                        uint32_t n_images, uint8_t* out) {
    if (!c || !out) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_dims(w, h));
    if (h > 65535) return fail(SSW_ERR_UNSUPPORTED, "synthetic frames are limited to 65535 rows");
    CKS(ctx_bind(c));
    for (uint32_t i0 = 0; i0 < n_images; i0 += 65535) {
        const uint32_t nb = std::min<uint32_t>(65535, n_images - i0);
        { KScope ks(c, "synth_frame"); synth_frame_kernel<<<dim3((w + 127) / 128, h, nb), 128, 0, c->stream>>>(out + (size_t)i0 * w * h * 3, w, h, seed, first_image + i0, 0u); }
        CK(cudaGetLastError());
    }
    return SSW_OK;
}

// rows [row0, row0 + n_rows) of synthetic frame `image` (the shard of one rank of a sharded frame)
extern "C" int ssw_synth_rows_rgb8_dev(ssw_ctx* c, uint32_t w, uint64_t seed, uint32_t image, uint32_t row0, uint32_t n_rows,
                                       uint8_t* out) {
    if (!c || !out) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_dims(w, n_rows));
    CKS(ctx_bind(c));
    for (uint32_t r = 0; r < n_rows; r += 32768) {
        const uint32_t nr = std::min<uint32_t>(32768, n_rows - r);
        { KScope ks(c, "synth_frame"); synth_frame_kernel<<<dim3((w + 127) / 128, nr, 1), 128, 0, c->stream>>>(out + (size_t)r * w * 3, w, nr, seed, image, row0 + r); }
        CK(cudaGetLastError());
    }
    return SSW_OK;
}

extern "C" int ssw_stage_forward_rgb8_dev(ssw_ctx* c, const uint8_t* rgb, uint32_t w, uint32_t h, uint32_t batch, float* plane) {
    if (!c || !rgb || !plane) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_dims(w, h));
    CKS(ctx_bind(c));
    return run_forward(c, PIX_RGB8, rgb, w, h, batch, plane, SSW_DCT2);
}

extern "C" int ssw_stage_topk_dev(ssw_ctx* c, const float* plane, uint32_t w, uint32_t h, uint32_t batch, int ordering,
                                  size_t k, uint32_t* idx) {
    if (!c || !plane || !idx) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_dims(w, h));
    if (k == 0 || k > (size_t)kTopkCap / 2 || k > (size_t)w * h - 1) return fail(SSW_ERR_INVALID, "k out of range for the fused top-k");
    CKS(ctx_bind(c));
    return run_topk_fast(c, plane, w, h, batch, ordering, (unsigned)k, idx, (long long)k, c->topk_full_hist);
}

extern "C" int ssw_stage_inverse_rgb8_dev(ssw_ctx* c, float* plane, const uint8_t* rgb_src, uint32_t w, uint32_t h,
                                          uint32_t batch, uint8_t* out_rgb) {
    if (!c || !plane || !rgb_src || !out_rgb) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_dims(w, h));
    CKS(ctx_bind(c));
    return run_inverse(c, plane, PIX_RGB8, rgb_src, w, h, batch, PIX_RGB8, out_rgb);
}

// ------------------------------------------------------------------------------------------------
// sharded frames: per-rank building blocks.  The exchange steps between them (the
// all-to-all transposes, the max / all-gather of the distributed top-k, the sum of the extracted
// vector) are issued by the host layer (spread_spectrum_watermarking_b200/sharded.py) through
// torch.distributed -- NCCL over NVLink on the GPUs.
// ------------------------------------------------------------------------------------------------
extern "C" int ssw_lines_forward_dev(ssw_ctx* c, int src_type, const void* src, uint32_t n, uint32_t n_lines, float* plane) {
    if (!c || !src || !plane) return fail(SSW_ERR_INVALID, "NULL argument");
    if (src_type < PIX_RGB8 || src_type > PIX_PLANE) return fail(SSW_ERR_INVALID, "bad pixel type");
    CKS(check_dims(n, n_lines));
    CKS(ctx_bind(c));
    return run_rows_forward(c, src_type, src, (int)n, (int)n_lines, 1, plane, 1.f, 1.f);
}

// the same over lines assembled from all-to-all blocks, read in place: src layout [chunks][ranks][n_lines][seg_len],
// sample m of a line in segment s = m / seg_len, which came from rank s / chunks, chunk s % chunks.
// SSW_ERR_UNSUPPORTED (caller falls back to an interleaving copy) unless seg_len and chunks are powers of two,
// seg_len % 4 == 0 and the line length has a compile-time plan.
extern "C" int ssw_lines_forward_seg_dev(ssw_ctx* c, const float* src, uint32_t n, uint32_t n_lines, uint32_t seg_len,
                                         uint32_t chunks, uint32_t ranks, float* plane) {
    if (!c || !src || !plane) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_dims(n, n_lines));
    if (seg_len == 0 || chunks == 0 || ranks == 0 || (uint64_t)seg_len * chunks * ranks != n)
        return fail(SSW_ERR_INVALID, "segments do not tile the line");
    int ss = 0, cs = 0;
    while ((1u << ss) < seg_len) ++ss;
    while ((1u << cs) < chunks) ++cs;
    if ((1u << ss) != seg_len || (1u << cs) != chunks || (seg_len & 3u) || src == plane)
        return fail(SSW_ERR_UNSUPPORTED, "segmented source lines: seg_len and chunks must be powers of two (seg_len >= 4), out of place");
    if (!fast::has_plan((int)n) && !fast::has_line1_plan((int)n))
        return fail(SSW_ERR_UNSUPPORTED, "segmented source lines need a planned line length");
    CKS(ctx_bind(c));
    c->seg.active = true; c->seg.seg_shift = ss; c->seg.chunk_shift = cs; c->seg.ranks = (int)ranks; c->seg.lines = (int)n_lines;
    const int rc = run_rows_forward(c, PIX_PLANE, src, (int)n, (int)n_lines, 1, plane, 1.f, 1.f);
    c->seg.active = false;
    return rc;
}

// inverse counterpart: the coefficient lines are all-to-all blocks [chunks][ranks][n_lines][seg_len], read in place
extern "C" int ssw_lines_inverse_seg_dev(ssw_ctx* c, const float* src, uint32_t n, uint32_t n_lines, uint32_t seg_len,
                                         uint32_t chunks, uint32_t ranks, float scale, int dst_type, void* dst, int pix_type,
                                         const void* pixels) {
    if (!c || !src || !dst) return fail(SSW_ERR_INVALID, "NULL argument");
    if (dst_type != PIX_PLANE && !pixels) return fail(SSW_ERR_INVALID, "pixel output needs the original pixels");
    CKS(check_dims(n, n_lines));
    if (seg_len == 0 || chunks == 0 || ranks == 0 || (uint64_t)seg_len * chunks * ranks != n)
        return fail(SSW_ERR_INVALID, "segments do not tile the line");
    int ss = 0, cs = 0;
    while ((1u << ss) < seg_len) ++ss;
    while ((1u << cs) < chunks) ++cs;
    if ((1u << ss) != seg_len || (1u << cs) != chunks || (const void*)src == dst)
        return fail(SSW_ERR_UNSUPPORTED, "segmented source lines: seg_len and chunks must be powers of two, out of place");
    if (!fast::has_plan((int)n) && !fast::has_line1_plan((int)n))
        return fail(SSW_ERR_UNSUPPORTED, "segmented source lines need a planned line length");
    CKS(ctx_bind(c));
    c->seg.active = true; c->seg.seg_shift = ss; c->seg.chunk_shift = cs; c->seg.ranks = (int)ranks; c->seg.lines = (int)n_lines;
    const int rc = run_rows_inverse(c, const_cast<float*>(src), dst_type == PIX_PLANE ? PIX_PLANE : pix_type, pixels, (int)n,
                                    (int)n_lines, 1, dst_type, dst, scale);
    c->seg.active = false;
    return rc;
}

extern "C" int ssw_lines_inverse_dev(ssw_ctx* c, float* plane, uint32_t n, uint32_t n_lines, float scale, int dst_type,
                                     void* dst, int src_type, const void* src) {
    if (!c || !plane || !dst) return fail(SSW_ERR_INVALID, "NULL argument");
    if (dst_type != PIX_PLANE && !src) return fail(SSW_ERR_INVALID, "pixel output needs the original pixels");
    CKS(check_dims(n, n_lines));
    CKS(ctx_bind(c));
    return run_rows_inverse(c, plane, dst_type == PIX_PLANE ? PIX_PLANE : src_type, src, (int)n, (int)n_lines, 1, dst_type, dst, scale);
}

extern "C" int ssw_transpose_dev(ssw_ctx* c, const float* src, uint32_t rows, uint32_t cols, int64_t src_ld, int64_t src_bstride,
                                 float* dst, int64_t dst_ld, int64_t dst_bstride, uint32_t batch) {
    if (!c || !src || !dst) return fail(SSW_ERR_INVALID, "NULL argument");
    if (rows == 0 || cols == 0 || batch == 0) return SSW_OK;
    if (batch > 65535 || (rows + 31) / 32 > 65535) return fail(SSW_ERR_INVALID, "transpose grid out of range");
    CKS(ctx_bind(c));
    {
        KScope ks(c, "transpose");
        launch_pdl(c, transpose_kernel, dim3((cols + 31) / 32, (rows + 31) / 32, batch), 256, 0, c->stream, 
            src, rows, cols, src_ld, src_bstride, dst, dst_ld, dst_bstride);
    }
    CK(cudaGetLastError());
    return SSW_OK;
}

static int shard_order(const ssw_shard* sh, int ordering, OrderConsts* oc) {
    if (!sh || sh->ncols == 0 || sh->col0 + sh->ncols > sh->width || sh->height == 0)
        return fail(SSW_ERR_INVALID, "bad shard layout");
    if ((uint64_t)sh->width * sh->height >= 0xFFFFFFFFull) return fail(SSW_ERR_UNSUPPORTED, "more than 2^32-1 pixels per frame");
    if (ordering < 0 || ordering > 2) return fail(SSW_ERR_UNSUPPORTED, "only Energy/EnergyOrthogonal/Legacy orderings run on the device");
    *oc = make_order(ordering, (int)sh->width, (int)sh->height);
    oc->t_ld = sh->height;
    oc->t_col0 = sh->col0;
    return SSW_OK;
}

// lower bound (histogram bin) of the frame's k-th largest key from this rank's low-frequency block
extern "C" int ssw_shard_topk_bin_dev(ssw_ctx* c, const float* plane, const ssw_shard* sh, int ordering, size_t k, uint32_t* bin_dev) {
    if (!c || !plane || !bin_dev) return fail(SSW_ERR_INVALID, "NULL argument");
    OrderConsts oc;
    CKS(shard_order(sh, ordering, &oc));
    if (k == 0 || k > (size_t)kTopkCap / 2) return fail(SSW_ERR_UNSUPPORTED, "sharded top-k supports mark lengths up to 4096");
    CKS(ctx_bind(c));
    CKS(ensure_topk_scratch(c, 1));
    TopkScratch ts = c->ts;
    ts.sel_bin = bin_dev;
    {
        KScope ks(c, "topk_block_bin");
        // local plane [ncols][height]: "width" of the block kernel is the local line length
        launch_pdl(c, topk_block_bin_kernel, dim3(1), kBinThreads, 0, c->stream, plane, 0, sh->height, sh->ncols, (unsigned)k, oc, ts);
    }
    CK(cudaGetLastError());
    return SSW_OK;
}

// every local coefficient whose key bin is >= *bin_dev -> cand_dev[0 .. *count_dev) as (key << 32 | ~flat index);
// cand_dev holds SSW_TOPK_CAP entries, *count_dev may exceed it (overflow, reported by the merge)
extern "C" int ssw_shard_topk_collect_dev(ssw_ctx* c, const float* plane, const ssw_shard* sh, int ordering,
                                          const uint32_t* bin_dev, uint64_t* cand_dev, uint32_t* count_dev) {
    if (!c || !plane || !bin_dev || !cand_dev || !count_dev) return fail(SSW_ERR_INVALID, "NULL argument");
    OrderConsts oc;
    CKS(shard_order(sh, ordering, &oc));
    CKS(ctx_bind(c));
    const size_t n = (size_t)sh->ncols * sh->height;
    if (n >= 0xFFFFFFFFull) return fail(SSW_ERR_UNSUPPORTED, "shard too large");
    TopkScratch ts{};
    ts.sel_bin = const_cast<uint32_t*>(bin_dev);
    ts.cand_count = count_dev;
    ts.cand = (unsigned long long*)cand_dev;
    CK(cudaMemsetAsync(count_dev, 0, sizeof(uint32_t), c->stream));
    const unsigned blocks = (unsigned)std::max<size_t>(1, std::min<size_t>((n / 4 + 511) / 512, (size_t)c->sm_count * 4));
    {
        KScope ks(c, "topk_collect");
        launch_pdl(c, topk_collect_kernel, dim3(blocks, 1), 512, 0, c->stream, plane, 0, (unsigned)n, oc, ts);
    }
    CK(cudaGetLastError());
    return SSW_OK;
}

// merge the candidate lists gathered from all ranks ([n_lists][SSW_TOPK_CAP] + counts) into the first k
// ordered flat indices; *overflow_dev != 0 afterwards means the lists did not fit (caller repairs)
extern "C" int ssw_shard_topk_merge_dev(ssw_ctx* c, const uint64_t* lists_dev, const uint32_t* counts_dev, uint32_t n_lists,
                                        size_t k, uint32_t* idx_dev, uint32_t* overflow_dev) {
    if (!c || !lists_dev || !counts_dev || !idx_dev || !overflow_dev) return fail(SSW_ERR_INVALID, "NULL argument");
    if (n_lists == 0 || n_lists > 64) return fail(SSW_ERR_INVALID, "1..64 candidate lists");
    if (k == 0 || k > (size_t)kTopkCap / 2) return fail(SSW_ERR_UNSUPPORTED, "sharded top-k supports mark lengths up to 4096");
    CKS(ctx_bind(c));
    CKS(ensure_topk_scratch(c, 1));
    TopkScratch ts = c->ts;
    ts.overflow = overflow_dev;
    CK(cudaMemsetAsync(overflow_dev, 0, sizeof(uint32_t), c->stream));
    {
        KScope ks(c, "topk_concat");
        launch_pdl(c, topk_concat_kernel, 1, 256, 0, c->stream, (const unsigned long long*)lists_dev, counts_dev, n_lists, (unsigned)kTopkCap, (unsigned)kTopkCap, ts);
    }
    const void* key = (const void*)topk_rank_kernel;
    const int smem = kTopkCap * (int)sizeof(unsigned long long);
    if (c->smem_attr.find(key) == c->smem_attr.end()) {
        CK(cudaFuncSetAttribute(topk_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        c->smem_attr[key] = smem;
    }
    { KScope ks(c, "topk_rank"); TopkApply none; std::memset(&none, 0, sizeof(none));
      launch_pdl(c, topk_rank_kernel, dim3(kRankCtas, 1), kRankThreads, smem, c->stream, ts, (unsigned)k, idx_dev, 0, none); }
    CK(cudaGetLastError());
    return SSW_OK;
}

static ShardLayout shard_layout(const ssw_shard* sh) {
    ShardLayout L;
    L.width = sh->width; L.height = sh->height; L.col0 = sh->col0; L.ncols = sh->ncols;
    return L;
}

extern "C" int ssw_shard_embed_dev(ssw_ctx* c, float* plane, const ssw_shard* sh, const uint32_t* idx_dev, size_t k,
                                   const float* marks_dev, size_t mark_stride, size_t n_marks, const uint32_t* lens_dev,
                                   const ssw_config* cfg) {
    if (!c || !plane || !idx_dev || !marks_dev) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_cfg(cfg));
    OrderConsts oc;
    CKS(shard_order(sh, cfg->ordering, &oc));
    if (k == 0 || n_marks == 0) return SSW_OK;
    CKS(ctx_bind(c));
    {
        KScope ks(c, "embed_scatter");
        launch_pdl(c, embed_scatter_shard_kernel, (unsigned)((k + 255) / 256), 256, 0, c->stream, 
            plane, shard_layout(sh), idx_dev, (unsigned)k, marks_dev, (long long)mark_stride, (int)n_marks, lens_dev,
            cfg->method, cfg->alpha);
    }
    CK(cudaGetLastError());
    return SSW_OK;
}

extern "C" int ssw_shard_extract_dev(ssw_ctx* c, const float* base_plane, const float* derived_plane, const ssw_shard* sh,
                                     const uint32_t* idx_dev, size_t n, const ssw_config* cfg, float* out_dev) {
    if (!c || !base_plane || !derived_plane || !idx_dev || !out_dev) return fail(SSW_ERR_INVALID, "NULL argument");
    CKS(check_cfg(cfg));
    OrderConsts oc;
    CKS(shard_order(sh, cfg->ordering, &oc));
    if (n >= (size_t)sh->width * sh->height) return fail(SSW_ERR_INVALID, "Desired extraction length exceeds available coefficients.");
    if (n == 0) return SSW_OK;
    CKS(ctx_bind(c));
    {
        KScope ks(c, "extract_gather");
        launch_pdl(c, extract_gather_shard_kernel, (unsigned)((n + 255) / 256), 256, 0, c->stream, 
            base_plane, derived_plane, shard_layout(sh), idx_dev, (unsigned)n, cfg->method, cfg->alpha, out_dev);
    }
    CK(cudaGetLastError());
    return SSW_OK;
}

#include "sharded_api.cuh"
