// TEST INFRASTRUCTURE ONLY -- CPU emulation of the CUDA kernel *bodies* for index-math debugging.
//
// The build container has no GPU.  The DCT kernel bodies in csrc/dct_kernels.cuh are written as
// __host__ __device__ functions of (tile, tid, nthreads); this file compiles them with g++ and runs
// one std::thread per CUDA thread with a std::barrier standing in for __syncthreads(), so the very
// same index arithmetic is exercised by the CPU-only test-suite (tests/test_emul_dct.py).
// It is never linked into libssw.so and is not a fallback path: the product has no CPU path.
#include <barrier>
#include <cstring>
#include <thread>
#include <vector>

#include "../../spread_spectrum_watermarking_b200/csrc/dct_kernels.cuh"

namespace ssw {
static thread_local std::barrier<>* tl_barrier = nullptr;
void host_barrier() { tl_barrier->arrive_and_wait(); }
}  // namespace ssw

using namespace ssw;

namespace {

struct PlanHolder {
    DctPlanHost h;
    DctPlanDev d;
    explicit PlanHolder(int n) : h(make_dct_plan(n)) {
        std::memset(&d, 0, sizeof(d));
        d.n = h.n; d.npad = h.npad; d.tp = h.tp; d.nstages = h.nstages;
        for (int i = 0; i < h.nstages; ++i) {
            d.stages[i] = h.stages[i];
            d.ns_magic[i] = h.stages[i].ns > 1 ? (unsigned)((1ull << 32) / (unsigned)h.stages[i].ns + 1) : 0u;
        }
        d.stage_tw = (const cplx*)h.stage_tw.data();
        d.wn = (const cplx*)h.wn.data();
        d.t4 = (const cplx*)h.t4.data();
    }
};

template <class Body>
void launch(int ntiles, int nthreads, size_t smem_elems, Body body) {
    for (int tile = 0; tile < ntiles; ++tile) {
        std::vector<cplx> smem(smem_elems);
        std::barrier<> bar(nthreads);
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t] {
                tl_barrier = &bar;
                body(smem.data(), tile, t, nthreads);
            });
        for (auto& x : th) x.join();
    }
}

LineArgs make_args(const PlanHolder& ph, int w, int h, int P) {
    LineArgs a;
    std::memset(&a, 0, sizeof(a));
    a.plan = ph.d; a.w = w; a.h = h; a.P = P;
    a.scale0 = 1.f; a.scalen = 1.f;
    a.tiles_per_image = 1 << 30;
    return a;
}

}  // namespace

extern "C" {

int emul_plan(int n, int* radices, int* tp, int* npad) {
    DctPlanHost p = make_dct_plan(n);
    if (!p.error.empty()) return -1;
    for (int i = 0; i < p.nstages; ++i) radices[i] = p.stages[i].radix;
    *tp = p.tp; *npad = p.npad;
    return p.nstages;
}

void emul_dft(int R, float* x) {
    cplx* c = (cplx*)x;
    switch (R) {
        case 2: Dft<2>::run(c); break;
        case 3: Dft<3>::run(c); break;
        case 4: Dft<4>::run(c); break;
        case 5: Dft<5>::run(c); break;
        case 6: Dft<6>::run(c); break;
        case 8: Dft<8>::run(c); break;
        case 9: Dft<9>::run(c); break;
        case 10: Dft<10>::run(c); break;
        case 12: Dft<12>::run(c); break;
        case 15: Dft<15>::run(c); break;
        case 16: Dft<16>::run(c); break;
    }
}

// kind: 0 = DCT2, 1 = DCT2Orthogonal, 2 = DCT3 (reference src/dct2d.rs Type).  data: [h][w] in place.
// Pr / Pc = line pairs per tile for the row / column pass, G = pairs transformed concurrently.
int emul_dct2d(int kind, int w, int h, float* data, int Pr, int Pc, int G) {
    PlanHolder pw(w), phh(h);
    if (!pw.h.error.empty() || !phh.h.error.empty()) return -1;
    LineArgs ar = make_args(pw, w, h, Pr);
    LineArgs ac = make_args(phh, w, h, Pc);
    ar.plane = data; ac.plane = data; ar.src = data; ar.dst = data;
    const int tr = (h + 2 * Pr - 1) / (2 * Pr), tc = (w + 2 * Pc - 1) / (2 * Pc);
    if (kind == 0 || kind == 1) {
        if (kind == 1) {
            ar.scale0 = std::sqrt(1.0f / (4.0f * (float)w)); ar.scalen = std::sqrt(1.0f / (2.0f * (float)w));
            ac.scale0 = std::sqrt(1.0f / (4.0f * (float)h)); ac.scalen = std::sqrt(1.0f / (2.0f * (float)h));
        }
        launch(tr, G * pw.d.tp, (size_t)Pr * pw.d.npad, [&](cplx* s, int tile, int t, int nt) { row_fwd_body<PIX_PLANE>(ar, s, tile, t, nt); });
        launch(tc, G * phh.d.tp, (size_t)Pc * phh.d.npad, [&](cplx* s, int tile, int t, int nt) { col_fwd_body(ac, s, tile, t, nt); });
    } else {
        ac.scale0 = 1.f;
        ar.scale0 = 4.0f / (float)(w * h);
        launch(tc, G * phh.d.tp, (size_t)Pc * phh.d.npad, [&](cplx* s, int tile, int t, int nt) { col_inv_body(ac, s, tile, t, nt); });
        launch(tr, G * pw.d.tp, (size_t)Pr * pw.d.npad, [&](cplx* s, int tile, int t, int nt) { row_inv_body<PIX_PLANE, PIX_PLANE>(ar, s, tile, t, nt); });
    }
    return 0;
}

// RGB8 -> coefficient plane (fused colour conversion + forward 2-D DCT)
int emul_rgb8_forward(int w, int h, const unsigned char* rgb, float* plane, int Pr, int Pc, int G) {
    PlanHolder pw(w), phh(h);
    if (!pw.h.error.empty() || !phh.h.error.empty()) return -1;
    LineArgs ar = make_args(pw, w, h, Pr);
    LineArgs ac = make_args(phh, w, h, Pc);
    ar.src = rgb; ar.plane = plane; ac.plane = plane;
    const int tr = (h + 2 * Pr - 1) / (2 * Pr), tc = (w + 2 * Pc - 1) / (2 * Pc);
    launch(tr, G * pw.d.tp, (size_t)Pr * pw.d.npad, [&](cplx* s, int tile, int t, int nt) { row_fwd_body<PIX_RGB8>(ar, s, tile, t, nt); });
    launch(tc, G * phh.d.tp, (size_t)Pc * phh.d.npad, [&](cplx* s, int tile, int t, int nt) { col_fwd_body(ac, s, tile, t, nt); });
    return 0;
}

// coefficient plane (destroyed) + original RGB8 -> RGB8 (inverse 2-D DCT + fused YIQ->RGB8)
int emul_rgb8_inverse(int w, int h, float* plane, const unsigned char* rgb_src, unsigned char* rgb_out, int Pr, int Pc, int G) {
    PlanHolder pw(w), phh(h);
    if (!pw.h.error.empty() || !phh.h.error.empty()) return -1;
    LineArgs ar = make_args(pw, w, h, Pr);
    LineArgs ac = make_args(phh, w, h, Pc);
    ac.plane = plane; ac.scale0 = 1.f;
    ar.plane = plane; ar.src = rgb_src; ar.dst = rgb_out; ar.scale0 = 4.0f / (float)(w * h);
    const int tr = (h + 2 * Pr - 1) / (2 * Pr), tc = (w + 2 * Pc - 1) / (2 * Pc);
    launch(tc, G * phh.d.tp, (size_t)Pc * phh.d.npad, [&](cplx* s, int tile, int t, int nt) { col_inv_body(ac, s, tile, t, nt); });
    launch(tr, G * pw.d.tp, (size_t)Pr * pw.d.npad, [&](cplx* s, int tile, int t, int nt) { row_inv_body<PIX_RGB8, PIX_RGB8>(ar, s, tile, t, nt); });
    return 0;
}

}  // extern "C"
