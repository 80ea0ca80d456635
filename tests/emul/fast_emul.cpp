// TEST INFRASTRUCTURE ONLY -- CPU execution of the fast-path kernel PHASES (csrc/dct_fast.cuh).
//
// The fast kernels are sequences of barrier-separated phases over a per-thread register state; the
// __global__ wrapper runs phase p for all threads, __syncthreads(), phase p+1, ...  Here the same
// phase functions are called for tid = 0..THREADS-1 in turn, with the per-thread states kept in an
// array, which is an exact model of that execution.  Never linked into libssw.so; not a fallback.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../spread_spectrum_watermarking_b200/csrc/fast_dispatch.h"
#include "../../spread_spectrum_watermarking_b200/csrc/dct_pipe.cuh"

using namespace ssw;
using namespace ssw::fast;

namespace {

template <class K>
void emulate(const FastArgs& a, int ntiles) {
    std::vector<cplx> smem(K::SMEM / sizeof(cplx) + 8);
    std::vector<typename K::Thread> th(K::THREADS);
    for (int tile = 0; tile < ntiles; ++tile) {
        static_for<K::NPH>([&](auto ph) {
            constexpr int p = decltype(ph)::value;
            for (int tid = 0; tid < K::THREADS; ++tid) K::template phase<p>(a, smem.data(), tile, tid, th[tid]);
        });
    }
}

// prefetching kernels: same schedule as fast_kernel_pf (prefetch == immediate copy on the CPU)
template <class K>
void emulate_pf(FastArgs a, int ntiles, int per) {
    std::vector<cplx> smem(K::SMEM / sizeof(cplx) + 8);
    std::vector<typename K::Thread> th(K::THREADS);
    a.total_tiles = ntiles; a.tiles_per_cta = per;
    for (int cta = 0; cta * per < ntiles; ++cta) {
        int tile = cta * per;
        const int end = std::min(tile + per, ntiles);
        for (int tid = 0; tid < K::THREADS; ++tid) K::prefetch(a, smem.data(), tile, tid);
        for (; tile < end; ++tile) {
            for (int tid = 0; tid < K::THREADS; ++tid) K::template phase<0>(a, smem.data(), tile, tid, th[tid]);
            if (tile + 1 < end)
                for (int tid = 0; tid < K::THREADS; ++tid) K::prefetch(a, smem.data(), tile + 1, tid);
            static_for<K::NPH - 1>([&](auto ph) {
                constexpr int p = decltype(ph)::value + 1;
                for (int tid = 0; tid < K::THREADS; ++tid) K::template phase<p>(a, smem.data(), tile, tid, th[tid]);
            });
        }
    }
}

template <class P>
struct Tables {
    std::vector<float> tw, t4;
    Tables() : tw(2 * (size_t)P::TW_TOTAL + 2), t4(2 * (size_t)P::N) {
        make_stage_twiddles<P>(tw.data());
        for (int j = 0; j < P::N; ++j) {
            const double b = -M_PI * (double)j / (2.0 * (double)P::N);
            t4[2 * j] = (float)std::cos(b);
            t4[2 * j + 1] = (float)std::sin(b);
        }
    }
};

FastArgs base_args(int w, int h) {
    FastArgs a;
    std::memset(&a, 0, sizeof(a));
    a.w = w; a.h = h;
    a.scale0 = 1.f; a.scalen = 1.f;
    a.src_stride = a.plane_stride = a.dst_stride = (long long)w * h;
    a.seg_shift = -1;
    a.dbg_skip = 0;
    return a;
}

// persistent TMA column pipelines (dct_pipe.cuh): memcpy stands in for the tensor-map copies -- the sample side of a
// tile is [even rows | odd rows] (the 4-D parity view), the coefficient side the rows in natural order; columns past
// the frame read as zero and are not written back (TMA out-of-bounds rules)
// grid > 0: the kernel's split schedule for `grid` CTAs -- whole rounds of whole tiles, the remainder as half tiles (G
// columns, one round over G/2 line pairs, buffer layout [row][G floats]) when it would occupy at most half of the CTAs
template <class K>
void emulate_col_pipe(PipeArgs a, float* plane, const cplx* tw, const cplx* t4, int grid = 0) {
    constexpr int N = K::N, G = K::G;
    std::vector<cplx> buf((size_t)N * G), fft((size_t)K::TEAMS * K::PITCH + 8);
    std::vector<typename K::Thread> th(K::NC);
    a.tiles_per_image = K::tiles_per_image(a.w, a.h);
    a.total_tiles = a.tiles_per_image * a.batch;
    auto sample_row = [](int n) { return n < N / 2 ? 2 * n : 2 * (n - N / 2) + 1; };   // buffer row -> frame row (sample side)
    int full_tiles = a.total_tiles;
    if constexpr (K::HALF_OK) {
        if (grid > 0 && a.batch == 1) {
            const int rounds = a.total_tiles / grid, rem = a.total_tiles % grid;
            if (rounds >= 1 && rem > 0 && 2 * rem <= grid && a.w % (2 * G) == 0) full_tiles = rounds * grid;
        }
        for (int hh = 0; hh < 2 * (a.total_tiles - full_tiles); ++hh) {
            constexpr int GH = K::GH;
            const int tile = full_tiles + (hh >> 1), img = tile / a.tiles_per_image;
            const int c0 = (tile - img * a.tiles_per_image) * 2 * G + (hh & 1) * G;
            float* pl = plane + (size_t)img * a.w * a.h;
            for (int n = 0; n < N; ++n) {
                const int r = K::INVERSE ? n : sample_row(n);
                for (int x = 0; x < G; ++x) ((float*)&buf[(size_t)n * GH])[x] = (c0 + x < a.w) ? pl[(size_t)r * a.w + c0 + x] : 0.f;
            }
            static_for<K::NPHASES>([&](auto ph) {
                constexpr int p = decltype(ph)::value;
                for (int c = 0; c < K::NC; ++c) K::template phase<p, GH>(a, buf.data(), fft.data(), tw, t4, 0, c, th[c]);
            });
            for (int n = 0; n < N; ++n) {
                const int r = K::INVERSE ? sample_row(n) : n;
                for (int x = 0; x < G; ++x)
                    if (c0 + x < a.w) pl[(size_t)r * a.w + c0 + x] = ((float*)&buf[(size_t)n * GH])[x];
            }
        }
    }
    for (int tile = 0; tile < full_tiles; ++tile) {
        const int img = tile / a.tiles_per_image, c0 = (tile - img * a.tiles_per_image) * 2 * G;
        float* pl = plane + (size_t)img * a.w * a.h;
        for (int n = 0; n < N; ++n) {
            const int r = K::INVERSE ? n : sample_row(n);
            for (int x = 0; x < 2 * G; ++x) {
                const float v = (c0 + x < a.w) ? pl[(size_t)r * a.w + c0 + x] : 0.f;
                ((float*)&buf[(size_t)n * G])[K::SWZ ? (x ^ (n & 4)) : x] = v;   // 32-byte swizzle of the tensor map
            }
        }
        for (int rd = 0; rd < K::ROUNDS; ++rd)
            static_for<K::NPHASES>([&](auto ph) {
                constexpr int p = decltype(ph)::value;
                for (int c = 0; c < K::NC; ++c) K::template phase<p>(a, buf.data(), fft.data(), tw, t4, rd, c, th[c]);
            });
        for (int n = 0; n < N; ++n) {
            const int r = K::INVERSE ? sample_row(n) : n;
            for (int x = 0; x < 2 * G; ++x)
                if (c0 + x < a.w) pl[(size_t)r * a.w + c0 + x] = ((float*)&buf[(size_t)n * G])[K::SWZ ? (x ^ (n & 4)) : x];
        }
    }
}

// persistent bulk-copy row pipelines (dct_pipe.cuh): memcpy stands in for the linear bulk copies of a tile (2*TEAMS
// adjacent rows); the phases run for all compute threads one after the other
template <class K>
void emulate_row_pipe(RowPipeArgs a, int kcut_arg = -1) {
    std::vector<unsigned char> A(K::A_BYTES + 16), B(K::B_BYTES + 16);
    std::vector<cplx> fft((size_t)K::TEAMS * K::P::PITCH + 8);
    std::vector<typename K::Thread> th(K::NC);
    a.tiles_per_image = K::tiles_per_image(a.w, a.h);
    a.total_tiles = a.tiles_per_image * a.batch;
    const size_t frame_px = (size_t)a.w * a.h;
    for (int tile = 0; tile < a.total_tiles; ++tile) {
        const int img = tile / a.tiles_per_image;
        const size_t px0 = img * frame_px + (size_t)(tile - img * a.tiles_per_image) * K::ROWS * K::N;
        if (K::INVERSE) {
            if constexpr (K::INPLACE) {   // one row pair per team, straight into its FFT buffer
                for (int g = 0; g < K::TEAMS; ++g)
                    std::memcpy((unsigned char*)(fft.data() + (size_t)g * K::P::PITCH), a.plane + px0 + (size_t)g * 2 * K::N, 2 * K::COEF_ROW);
            } else {
                std::memcpy(A.data(), a.plane + px0, K::A_BYTES);
            }
            std::memcpy(B.data(), a.pix + 3 * px0, K::B_BYTES);
        } else {
            std::memcpy(A.data(), a.pix + 3 * px0, K::A_BYTES);
        }
        static_for<K::NPH>([&](auto ph) {
            constexpr int p = decltype(ph)::value;
            for (int c = 0; c < K::NC; ++c)
                K::template phase<p>(a, A.data(), fft.data(), B.data(), K::INVERSE ? nullptr : a.plane + px0, c, th[c], kcut_arg >= 0 ? kcut_arg : K::N + 1);
        });
        if (K::INVERSE) std::memcpy(a.out + 3 * px0, B.data(), K::B_BYTES);
    }
}

}  // namespace

extern "C" {

// row pipelines: forward (pix -> plane) or inverse (plane + original pix -> out); -2 when the length has no pipeline
int emul_row_pipe_cut(int inverse, const unsigned char* pix, int w, int h, int batch, float* plane, unsigned char* out, float scale0, float scalen,
                      int kcut, float gain);
int emul_row_pipe(int inverse, const unsigned char* pix, int w, int h, int batch, float* plane, unsigned char* out, float scale0, float scalen) {
    return emul_row_pipe_cut(inverse, pix, w, h, batch, plane, out, scale0, scalen, -1, 1.f);
}
// kcut >= 0 (inverse): coefficients of columns >= kcut are multiplied by `gain` as they are read (RowPipeArgs::col_cut_img)
int emul_row_pipe_cut(int inverse, const unsigned char* pix, int w, int h, int batch, float* plane, unsigned char* out, float scale0, float scalen,
                      int kcut, float gain) {
    bool ran = false;
    with_plan(w, [&](auto p) {
        using P = decltype(p);
        using Cfg = RowPipeCfg<P>;
        if constexpr (Cfg::OK) {
            if (!Cfg::Fwd::supports(w, h)) return;
            Tables<P> tb;
            RowPipeArgs a;
            std::memset(&a, 0, sizeof(a));
            a.w = w; a.h = h; a.batch = batch; a.pix = pix; a.plane = plane; a.out = out; a.scale0 = scale0; a.scalen = scalen;
            a.tw = (const cplx*)tb.tw.data(); a.t4 = (const cplx*)tb.t4.data();
            a.col_gain = gain;
            // inverse == 2: the in-place shape (RowPipeCfg::InvP); lengths without one report -2
            if (inverse == 2) { if constexpr (Cfg::INPLACE_OK) emulate_row_pipe<typename Cfg::InvP>(a, kcut); else return; }
            else if (inverse) emulate_row_pipe<typename Cfg::Inv>(a, kcut);
            else emulate_row_pipe<typename Cfg::Fwd>(a);
            ran = true;
        }
    });
    return ran ? 0 : -2;
}

// variant: 1 = 4 pairs x 4 teams, 2 = 4 pairs x 2 teams (2 rounds), 3 = 2 pairs x 2 teams, 4 = variant 2 with the split
// schedule of a 2-CTA grid (remainder tiles as half tiles; batch 1 only)
int emul_col_pipe(int variant, int inverse, int w, int h, int batch, float* plane, float scale0, float scalen) {
    if ((w % 4) || (h & 1)) return -2;
    bool ran = false;
    return with_plan(h, [&](auto p) {
        using P = decltype(p);
        Tables<P> tb;
        PipeArgs a;
        std::memset(&a, 0, sizeof(a));
        a.w = w; a.h = h; a.batch = batch; a.scale0 = scale0; a.scalen = scalen;
        const cplx* tw = (const cplx*)tb.tw.data();
        const cplx* t4 = (const cplx*)tb.t4.data();
        auto run = [&](auto k) { emulate_col_pipe<decltype(k)>(a, plane, tw, t4, variant >= 4 ? 2 : 0); ran = true; };
        if constexpr (ColPipe<P, 4, 4, false>::FITS && (box_rows(P::N / 2, 256) * 16) % 128 == 0) {
            if (variant == 3) { if (inverse) run(ColPipe<P, 2, 2, true, 2>{}); else run(ColPipe<P, 2, 2, false, 2>{}); return; }
        }
        if constexpr (ColPipe<P, 4, 4, false>::FITS) {
            if (variant == 2 || variant == 4) { if (inverse) run(ColPipe<P, 4, 2, true>{}); else run(ColPipe<P, 4, 2, false>{}); }
            else { if (inverse) run(ColPipe<P, 4, 4, true>{}); else run(ColPipe<P, 4, 4, false>{}); }   // 1, and 5 = 1 with the split schedule
        }
    }) && ran ? 0 : -2;
}

int emul_fast_has_plan(int n) { return has_plan(n) ? 1 : 0; }
int emul_fast_col_pairs(int n) { int g = -1; with_plan(n, [&](auto p) { g = ColG<decltype(p)>::value; }); return g; }

// src_type: 0 = RGB8, 1 = RGB32F, 2 = plane (PIX_*).
// plane: [batch][h][w].  scale0/scalen: DCT2Orthogonal factors.
int emul_fast_row_fwd(int src_type, const void* src, int w, int h, int batch, float* plane, float scale0, float scalen) {
    return with_plan(w, [&](auto p) {
        using P = decltype(p);
        constexpr int G = RowG<P>::value;
        Tables<P> tb;
        FastArgs a = base_args(w, h);
        a.src = src; a.plane = plane; a.scale0 = scale0; a.scalen = scalen;
        a.tw = (const cplx*)tb.tw.data(); a.t4 = (const cplx*)tb.t4.data();
        auto run = [&](auto k) {
            using K = decltype(k);
            a.tiles_per_image = K::tiles_per_image(w, h);
            emulate<K>(a, a.tiles_per_image * batch);
        };
        if (src_type == PIX_RGB8) run(RowFwd<P, G, PIX_RGB8>{});
        else if (src_type == PIX_RGB32F) run(RowFwd<P, G, PIX_RGB32F>{});
        else run(RowFwd<P, G, PIX_PLANE>{});
    }) ? 0 : -2;
}

int emul_fast_col(int inverse, int w, int h, int batch, float* plane, float scale0, float scalen) {
    if (w % 4) return -2;
    return with_plan(h, [&](auto p) {
        using P = decltype(p);
        Tables<P> tb;
        FastArgs a = base_args(w, h);
        a.plane = plane; a.scale0 = scale0; a.scalen = scalen;
        a.tw = (const cplx*)tb.tw.data(); a.t4 = (const cplx*)tb.t4.data();
        constexpr int G = ColG<P>::value, TEAMS = ColTeams<P>::value;
        if constexpr (G > 0) {
            if (inverse) {
                using K = ColPass<P, G, TEAMS, true>;
                a.tiles_per_image = K::tiles_per_image(w, h);
                emulate<K>(a, a.tiles_per_image * batch);
            } else {
                using K = ColPass<P, G, TEAMS, false>;
                a.tiles_per_image = K::tiles_per_image(w, h);
                emulate<K>(a, a.tiles_per_image * batch);
            }
        }
    }) ? 0 : -2;
}

// dst_type / src_type: 0 = RGB8, 1 = RGB32F (src = original pixels), dst_type 2 = plane
int emul_fast_row_inv(int dst_type, int src_type, float* plane, const void* src, int w, int h, int batch, void* dst, float scale) {
    return with_plan(w, [&](auto p) {
        using P = decltype(p);
        constexpr int G = RowG<P>::value;
        Tables<P> tb;
        FastArgs a = base_args(w, h);
        a.src = src; a.plane = plane; a.dst = dst; a.scale0 = scale;
        a.tw = (const cplx*)tb.tw.data(); a.t4 = (const cplx*)tb.t4.data();
        auto run = [&](auto k) {
            using K = decltype(k);
            a.tiles_per_image = K::tiles_per_image(w, h);
            emulate<K>(a, a.tiles_per_image * batch);
        };
        if (dst_type == PIX_PLANE) run(RowInv<P, G, PIX_PLANE, PIX_PLANE>{});
        else if (dst_type == PIX_RGB8 && src_type == PIX_RGB8) run(RowInv<P, G, PIX_RGB8, PIX_RGB8>{});
        else if (dst_type == PIX_RGB8) run(RowInv<P, G, PIX_RGB8, PIX_RGB32F>{});
        else if (src_type == PIX_RGB8) run(RowInv<P, G, PIX_RGB32F, PIX_RGB8>{});
        else run(RowInv<P, G, PIX_RGB32F, PIX_RGB32F>{});
    }) ? 0 : -2;
}

// single-line kernels: n_lines lines of length n (plan of n/2 points); src_type / dst_type as above
int emul_line1_fwd(int src_type, const void* src, int n, int n_lines, float* plane, float scale0, float scalen) {
    return with_line1_plan(n, [&](auto p) {
        using P = decltype(p);
        Tables<P> tb;   // stage twiddles of the M-point plan; t4 below must have N = 2M entries
        std::vector<float> t4(2 * (size_t)n);
        for (int j = 0; j < n; ++j) {
            const double b = -M_PI * (double)j / (2.0 * (double)n);
            t4[2 * j] = (float)std::cos(b); t4[2 * j + 1] = (float)std::sin(b);
        }
        FastArgs a = base_args(n, n_lines);
        a.src = src; a.plane = plane; a.scale0 = scale0; a.scalen = scalen;
        a.tw = (const cplx*)tb.tw.data(); a.t4 = (const cplx*)t4.data();
        a.tiles_per_image = n_lines;
        if (src_type == PIX_RGB8) emulate<Line1Fwd<P, PIX_RGB8>>(a, n_lines);
        else emulate<Line1Fwd<P, PIX_PLANE>>(a, n_lines);
    }) ? 0 : -2;
}

int emul_line1_inv(int dst_type, float* plane, const void* src, int n, int n_lines, void* dst, float scale) {
    return with_line1_plan(n, [&](auto p) {
        using P = decltype(p);
        Tables<P> tb;
        std::vector<float> t4(2 * (size_t)n);
        for (int j = 0; j < n; ++j) {
            const double b = -M_PI * (double)j / (2.0 * (double)n);
            t4[2 * j] = (float)std::cos(b); t4[2 * j + 1] = (float)std::sin(b);
        }
        FastArgs a = base_args(n, n_lines);
        a.src = src; a.plane = plane; a.dst = dst; a.scale0 = scale;
        a.tw = (const cplx*)tb.tw.data(); a.t4 = (const cplx*)t4.data();
        a.tiles_per_image = n_lines;
        if (dst_type == PIX_RGB8) emulate<Line1Inv<P, PIX_RGB8, PIX_RGB8>>(a, n_lines);
        else emulate<Line1Inv<P, PIX_PLANE, PIX_PLANE>>(a, n_lines);
    }) ? 0 : -2;
}

// forward pass over SEGMENTED f32 lines (layout [chunks][ranks][n_lines][seg_len]); line1 != 0: single-line kernels
int emul_fwd_segmented(int line1, const float* src, int n, int n_lines, int seg_len, int chunks, int ranks, float* plane) {
    int seg_shift = 0, chunk_shift = 0;
    while ((1 << seg_shift) < seg_len) ++seg_shift;
    while ((1 << chunk_shift) < chunks) ++chunk_shift;
    if ((1 << seg_shift) != seg_len || (1 << chunk_shift) != chunks || seg_len * chunks * ranks != n) return -3;
    auto fill = [&](FastArgs& a) {
        a.src = src; a.plane = plane;
        a.seg_shift = seg_shift; a.chunk_shift = chunk_shift; a.seg_ranks = ranks; a.seg_lines = n_lines;
    };
    if (line1) {
        return with_line1_plan(n, [&](auto p) {
            using P = decltype(p);
            Tables<P> tb;
            std::vector<float> t4(2 * (size_t)n);
            for (int j = 0; j < n; ++j) {
                const double b = -M_PI * (double)j / (2.0 * (double)n);
                t4[2 * j] = (float)std::cos(b); t4[2 * j + 1] = (float)std::sin(b);
            }
            FastArgs a = base_args(n, n_lines);
            fill(a);
            a.tw = (const cplx*)tb.tw.data(); a.t4 = (const cplx*)t4.data();
            a.tiles_per_image = n_lines;
            emulate<Line1Fwd<P, PIX_PLANE>>(a, n_lines);
        }) ? 0 : -2;
    }
    return with_plan(n, [&](auto p) {
        using P = decltype(p);
        constexpr int G = RowG<P>::value;
        Tables<P> tb;
        FastArgs a = base_args(n, n_lines);
        fill(a);
        a.tw = (const cplx*)tb.tw.data(); a.t4 = (const cplx*)tb.t4.data();
        using K = RowFwd<P, G, PIX_PLANE>;
        a.tiles_per_image = K::tiles_per_image(n, n_lines);
        emulate<K>(a, a.tiles_per_image);
    }) ? 0 : -2;
}

// forward RGB8 row pass through the prefetching kernel, `per` tiles per CTA
int emul_fast_row_fwd_pf(const void* src, int w, int h, int batch, float* plane, int per) {
    return with_plan(w, [&](auto p) {
        using P = decltype(p);
        if constexpr ((3 * P::N) % 16 == 0) {
            constexpr int G = RowG<P>::value;
            Tables<P> tb;
            FastArgs a = base_args(w, h);
            a.src = src; a.plane = plane;
            a.tw = (const cplx*)tb.tw.data(); a.t4 = (const cplx*)tb.t4.data();
            using K = RowFwdPF<P, G>;
            a.tiles_per_image = K::tiles_per_image(w, h);
            emulate_pf<K>(a, a.tiles_per_image * batch, per);
        }
    }) ? 0 : -2;
}

// inverse pass over SEGMENTED coefficient lines -> contiguous f32 lines `out` (scale applied)
int emul_inv_segmented(int line1, const float* src, int n, int n_lines, int seg_len, int chunks, int ranks, float* out, float scale) {
    int seg_shift = 0, chunk_shift = 0;
    while ((1 << seg_shift) < seg_len) ++seg_shift;
    while ((1 << chunk_shift) < chunks) ++chunk_shift;
    if ((1 << seg_shift) != seg_len || (1 << chunk_shift) != chunks || seg_len * chunks * ranks != n) return -3;
    auto fill = [&](FastArgs& a) {
        a.plane = const_cast<float*>(src); a.dst = out; a.scale0 = scale;
        a.seg_shift = seg_shift; a.chunk_shift = chunk_shift; a.seg_ranks = ranks; a.seg_lines = n_lines;
    };
    if (line1) {
        return with_line1_plan(n, [&](auto p) {
            using P = decltype(p);
            Tables<P> tb;
            std::vector<float> t4(2 * (size_t)n);
            for (int j = 0; j < n; ++j) {
                const double b = -M_PI * (double)j / (2.0 * (double)n);
                t4[2 * j] = (float)std::cos(b); t4[2 * j + 1] = (float)std::sin(b);
            }
            FastArgs a = base_args(n, n_lines);
            fill(a);
            a.tw = (const cplx*)tb.tw.data(); a.t4 = (const cplx*)t4.data();
            a.tiles_per_image = n_lines;
            emulate<Line1Inv<P, PIX_PLANE, PIX_PLANE>>(a, n_lines);
        }) ? 0 : -2;
    }
    return with_plan(n, [&](auto p) {
        using P = decltype(p);
        constexpr int G = RowG<P>::value;
        Tables<P> tb;
        FastArgs a = base_args(n, n_lines);
        fill(a);
        a.tw = (const cplx*)tb.tw.data(); a.t4 = (const cplx*)tb.t4.data();
        using K = RowInv<P, G, PIX_PLANE, PIX_PLANE>;
        a.tiles_per_image = K::tiles_per_image(n, n_lines);
        emulate<K>(a, a.tiles_per_image);
    }) ? 0 : -2;
}

// exactness of the division-free u8 -> [0,1] conversion: returns the number of mismatching inputs
int emul_fast_u8_unit_mismatches(void) {
    int bad = 0;
    for (unsigned v = 0; v < 256; ++v) bad += (u8_unit(v) != u8_to_unit(v));
    return bad;
}

// unit_to_u8_fast against the reference-faithful unit_to_u8(clamp01(v)) for a caller-supplied list
int emul_fast_unit_to_u8_mismatches(const float* v, int n) {
    int bad = 0;
    for (int i = 0; i < n; ++i) bad += (unit_to_u8_fast(v[i]) != unit_to_u8(v[i]));
    return bad;
}

// byte_to_float / unpack4_unit (PRMT + FADD form) against u8_to_unit for every byte value in every byte lane
int emul_fast_byte_unpack_mismatches(void) {
    int bad = 0;
    for (unsigned v = 0; v < 256; ++v) {
        const unsigned w0 = v | ((255u - v) << 8) | (((v * 7u) & 255u) << 16) | (((v * 13u + 5u) & 255u) << 24);
        const unsigned w1 = (w0 << 8) | (w0 >> 24), w2 = (w0 << 16) | (w0 >> 16);
        const unsigned ws[3] = {w0, w1, w2};
        float c[12];
        unpack4_unit(w0, w1, w2, c);
        for (int i = 0; i < 12; ++i) bad += (c[i] != u8_to_unit((ws[i / 4] >> (8 * (i % 4))) & 255u));
    }
    return bad;
}

// pack_u8x4 (two round-toward-zero adds + PRMT gather) against unit_to_u8 over ALL 2^32 float bit patterns,
// split over `parts` callers (part p covers the bit patterns congruent to ... a contiguous 1/parts range)
long long emul_fast_pack_u8_mismatches(unsigned part, unsigned parts) {
    long long bad = 0;
    const unsigned long long total = 1ull << 32, lo = total * part / parts, hi = total * (part + 1) / parts;
    for (unsigned long long b = lo; b < hi; b += 4) {
        float o[4];
        unsigned want = 0;
        for (int i = 0; i < 4; ++i) {
            o[i] = ssw_host_u2f((unsigned)(b + i));
            want |= unit_to_u8(o[i]) << (8 * i);
        }
        bad += (pack_u8x4(o) != want);
    }
    return bad;
}

}  // extern "C"
