// Column passes of the full-frame DCT as persistent, warp-specialised TMA pipelines (sm_100a).
//
// Same arithmetic as ColPass (dct_fast.cuh) -- replaces the column half of /root/reference/src/dct2d.rs:171-206 --
// but restructured around the asynchronous copy engine:
//   * one CTA per SM slot, looping over its column tiles (tile = 2*G adjacent columns x N rows);
//   * a PRODUCER warp: one elected thread issues cp.async.bulk.tensor (TMA) loads of the next tile into one of two
//     tile buffers, signalled by mbarriers (complete_tx), and TMA stores of the finished tile back to the plane,
//     retired with cp.async.bulk.wait_group.read before the buffer is loaded again;
//   * COMPUTE warps: TEAMS teams of P::T threads run the FFT of TEAMS line pairs per round out of a separate FFT
//     buffer, synchronising with named barriers (bar.sync id, count) instead of CTA-wide barriers;
//   * the Makhoul even/odd reordering is done by the TMA descriptor: the plane is addressed as a 4-D tensor
//     (column, row parity, row pair, image), so a tile arrives as [even rows ascending | odd rows ascending] and the
//     first FFT stage reads sample n of the permuted sequence at row  n < N/2 ? n : 3N/2-1-n  ("semi-Makhoul");
//     the coefficient side of the tile uses the natural 3-D view (column, row, image).
// Buffers of one CTA: BUF[2] (tile as it lives in the plane: N rows x 2G floats, once as input and once -- in the
// slots of the line pairs already consumed -- as output), FFT (TEAMS line pairs), optional copies of the twiddle tables.
//
// Like every fast kernel the compute part is written as PHASES, `ColPipe::phase<PH>(...)`, executed for all compute
// threads between barriers; tests/emul runs the same phase functions on the CPU with memcpy standing in for the TMA.
#pragma once
#include <type_traits>
#include "dct_fast.cuh"
#if defined(__CUDACC__)
#include "select_kernels.cuh"
#endif

namespace ssw {
namespace fast {

struct PipeArgs {
    int w, h, batch;
    int tiles_per_image, total_tiles;
    float scale0, scalen;      // forward: factors for k == 0 / k > 0; inverse: scale0 = output scale
    const cplx* tw;            // stage twiddles (global copy)
    const cplx* t4;            // exp(-i*pi*k/(2N)) (global copy)
    int pdl_late;
    long long* trace;          // diagnostics (ssw_ctx_set_trace): 64 time stamps per CTA, see pipe_trace_slot(); nullptr = off
#if defined(__CUDACC__)
    // forward pass only: the selection bin of the ordering that follows (select_kernels.cuh), computed on the fly.  The
    // tiles that produce the low-frequency block -- coefficient rows < hist_rows, columns < hist_cols -- add its
    // histogram (top 12 key bits) to ts.hist[image]; the last of them (ticket) finds the bin of the block's k-th
    // largest key, stores it in ts.sel_bin[image] and leaves histogram and ticket zeroed.  This replaces the
    // topk_block_bin kernel on the latency chain of the step.  ts.hist == nullptr: no histogram (inverse passes, derived frames).
    TopkScratch ts;
    unsigned hist_k;
    int hist_rows, hist_cols;
    OrderConsts oc;
    int tile_rot;   // tile = (first + j*step + total - tile_rot) % total: keeps the histogram tiles off the CTAs that get one tile more
    // collect != 0 (needs ts.hist): every tile also appends its candidates of the ordering -- coefficients at or above the
    // selection bin -- to ts.cand straight from shared memory, which takes the topk_collect kernel (one more read of the
    // plane) off the step.  The bin is published as ts.sel_bin[image] = bin | 0x80000000 by the last histogram tile; tiles
    // that finish earlier wait for it (all CTAs of a persistent grid are resident and the histogram tiles come first in
    // every CTA's sequence, so the wait cannot deadlock).  topk_rank clears the word again.
    int collect;
#endif
    // Tile schedule.  Tiles [0, full_tiles) are dealt round-robin to the CTAs as whole tiles (2G columns); the remaining
    // total_tiles - full_tiles tiles are cut into half_tiles = 2 * (total_tiles - full_tiles) half tiles (G columns), one
    // for each of the first half_tiles CTAs, processed in a single round.  One 4K frame is 480 tiles on 148 CTAs: 3 rounds
    // of whole tiles + 72 half tiles instead of a fourth round that keeps 36 SMs busy and 112 idle.
    // forward pass: tile_max[image * tiles_per_image + tile] = bits of the largest |coefficient| of the tile (atomicMax of
    // non-negative floats; zero between calls: the consumer, topk_collect_tiles, clears what it reads).  The candidate scan of
    // the ordering then reads only the tiles that can hold a candidate -- for natural frames the 16 tiles of the low-frequency
    // block instead of the whole plane.  nullptr: not wanted.
    unsigned* tile_max;
    int full_tiles, half_tiles;
    // forward pass of one frame with the histogram: the first hist_relief CTAs own the histogram tiles (~5-9 us of extra work
    // each, the last of them also the search) -- each gives its last whole tile away as two half tiles (see the kernel)
    int hist_relief;
    // inverse pass only: *col_limit = the largest column index whose coefficients were modified after the forward transform
    // (written by topk_rank); tiles beyond it are not processed.  The pass then runs OUT OF PLACE (tensor maps of two planes):
    // it reads the coefficient plane and overwrites, in the plane that still holds the row-transformed frame of the forward
    // pass, exactly the columns that changed -- the other columns of that plane ARE the inverse column transform of the
    // unchanged coefficients, up to the round-off of the two column passes.  nullptr: all tiles.
    const unsigned* col_limit;
    const unsigned* col_limit_img;   // [batch] the same per image: a scheduled tile beyond its own image's limit is not stored
    int tab_bulk;   // twiddle tables are 16-byte aligned: the producer stages them with bulk copies (else: a loop of all threads)
};

// largest divisor of n that is <= cap (rows per TMA box)
constexpr int box_rows(int n, int cap) {
    int best = 1;
    for (int d = 1; d <= cap; ++d) if (n % d == 0) best = d;
    return best;
}

// MINB_: CTAs the kernel is meant to keep resident per SM (1 or 2) -- bounds registers and shared memory per CTA
template <class P_, int G_, int TEAMS_, bool INVERSE_, int MINB_ = 1>
struct ColPipe {
    using P = P_;
    static_assert(G_ % TEAMS_ == 0 && (G_ == 2 || G_ == 4), "tile = 2 or 4 line pairs, whole rounds");
    static_assert(P_::N % 2 == 0, "even line length (row parity split)");
    static constexpr int G = G_, TEAMS = TEAMS_, ROUNDS = G_ / TEAMS_, N = P_::N, T = P_::T;
    static constexpr bool INVERSE = INVERSE_;
    static constexpr int NC = TEAMS_ * P_::T;          // compute threads
    static constexpr int THREADS = NC + 32;            // + the producer warp
    static constexpr int ROWB = 8 * G_;                // bytes per tile row (2G floats)
    static constexpr int BUF_BYTES = N * ROWB;
    // pitch between the FFT buffers of the teams (float2 units): offsets of 8 words (4 teams) / 16 words (2 teams) mod 32
    // keep the stride-R0 first-stage stores of TEAMS line pairs on distinct banks
    static constexpr int WANT = (TEAMS_ >= 4) ? 4 : 8;
    static constexpr int PITCH = P_::LINE + ((WANT - P_::LINE % 16) + 16) % 16;
    static constexpr int FFT_BYTES = TEAMS_ * PITCH * (int)sizeof(cplx);
    static constexpr int TW_BYTES = ((P_::TW_TOTAL * 8 + 15) / 16) * 16;
    static constexpr int T4_BYTES = (((N / 2 + 1) * 8 + 15) / 16) * 16;
    static constexpr int BAR_BYTES = 64;
    static constexpr int SMAX_OFF = 48;   // word behind the five mbarriers: running maximum of the tile (TRACK)
    static constexpr int BASE_BYTES = 2 * BUF_BYTES + FFT_BYTES + BAR_BYTES;
    static constexpr int MINB = MINB_;
    static constexpr int LIMIT = (228 * 1024) / MINB_ - 1024;   // 1 KB per resident CTA is reserved by the system
    // twiddle tables in shared memory when they fit (they would otherwise compete for the few KB of L1 that are left)
    static constexpr bool TW_SMEM = BASE_BYTES + TW_BYTES <= LIMIT;
    static constexpr bool T4_SMEM = BASE_BYTES + (TW_SMEM ? TW_BYTES : 0) + T4_BYTES <= LIMIT;
    static constexpr int SMEM = BASE_BYTES + (TW_SMEM ? TW_BYTES : 0) + (T4_SMEM ? T4_BYTES : 0);
    static constexpr bool FITS = BASE_BYTES <= LIMIT;
    static constexpr int OFF_FFT = 2 * BUF_BYTES, OFF_TW = OFF_FFT + FFT_BYTES, OFF_T4 = OFF_TW + (TW_SMEM ? TW_BYTES : 0),
                         OFF_BAR = OFF_T4 + (T4_SMEM ? T4_BYTES : 0);
    // TMA boxes: the sample side is the 4-D parity view (N/2 row pairs per parity), the coefficient side the 3-D view
    static constexpr int RB_HALF = box_rows(N / 2, 256), RB_FULL = box_rows(N, 256);
    static constexpr int NBOX_HALF = (N / 2) / RB_HALF, NBOX_FULL = N / RB_FULL;
    static_assert((RB_HALF * ROWB) % 128 == 0 && (RB_FULL * ROWB) % 128 == 0, "TMA boxes must start on 128-byte boundaries");
    // half tiles (G columns = G/2 line pairs in one round): the layout of the buffer becomes [row][G floats]
    static constexpr int GH = G_ / 2;
    // Only the 2-team, one-CTA-per-SM shape carries the half-tile and candidate-collecting code: it has registers to spare
    // (13 warps: up to 128 per thread), while the 4-team / two-CTA shapes run at the 72-register cap of 25-26 warps per SM
    // and spill as soon as the kernel grows (measured: 1080-point fwd_cols 242 -> 317 us per 64-frame launch).
    static constexpr bool HALF_OK = (G_ == 4 && (TEAMS_ == 2 || TEAMS_ == 4) && MINB_ == 1) && ((RB_HALF * ROWB / 2) % 128 == 0) && ((RB_FULL * ROWB / 2) % 128 == 0);
    static constexpr bool COLLECT_OK = (TEAMS_ == 2 && MINB_ == 1);
    // per-tile coefficient maxima for the tile-wise candidate scan (PipeArgs::tile_max): the shape of the batched launches only
    // (4 teams, two CTAs per SM) -- there the scan of the whole plane is 10 % of the step; on single frames the ~3 % the
    // bookkeeping costs the post pass outweigh the shorter scan (measured)
    static constexpr bool TRACK = (G_ == 4 && TEAMS_ == 4 && MINB_ == 2 && !INVERSE_);
    // 2 teams read and write HALF rows of the tile buffer per round (line pairs {0,1} or {2,3} of the four in a 32-byte row):
    // rows r and r+4 would meet on the same banks (2-way conflicts, measured 21 % of the wavefronts).  The tensor maps of
    // whole tiles therefore use the 32-byte swizzle (16-byte half of a row ^= bit 2 of the row index): lanes that walk
    // down the rows alternate between the bank groups.  sw(row, q): line pair q of buffer row `row` -> its slot.
    static constexpr bool SWZ = (G_ == 4 && TEAMS_ == 2);
    template <int GG> static SSW_HD int sw(int row, int q) { return (SWZ && GG == 4) ? (q ^ ((row >> 1) & 2)) : q; }
    using Thread = ThreadState<P_>;
    static int tiles_per_image(int w, int h) { (void)h; return (w + 2 * G - 1) / (2 * G); }

    // row of the tile buffer that holds sample n of the Makhoul-permuted sequence
    static SSW_HD int semi(int n) { return n < N / 2 ? n : 3 * N / 2 - 1 - n; }

    // ---- compute phases of one round `rd` (line pairs rd*TEAMS .. rd*TEAMS+TEAMS-1 of the tile) ------------------
    //   phase 0              : forward: first FFT stage, operands straight from the tile buffer (semi-Makhoul rows),
    //                                   lanes interleave (butterfly, pair) so that a warp reads whole rows
    //                          inverse: DCT-III pre pass, coefficient rows (natural order) -> FFT buffers
    //   phase 1 .. 2*NST-1   : forward: store of stage 0 is part of phase 0; phases (2s-1, 2s) = (load, store) of stage s >= 1
    //                          inverse: phases (2s+1, 2s+2) = (load, store) of stage s >= 0      [one more phase, see NPH_INV]
    //   last phase           : forward: DCT-II post pass, FFT buffers -> coefficient rows of the tile buffer
    //                          inverse: FFT buffers -> sample rows (semi-Makhoul) of the tile buffer
    // Barriers: phases that hand data from one team mapping to the (pair-interleaved) CTA mapping are separated by a
    // barrier over all compute threads, the stages in between only by the team's own barrier.
    static constexpr int NPH_FWD = 2 * P_::NST;       // 0: S0 load+store | 1..2(NST-1): stages 1.. | last: post
    static constexpr int NPH_INV = 2 * P_::NST + 2;   // 0: pre | 1..2NST: stages 0.. | last: output

    // forward phase 0: stage 0 from the tile buffer (GG: line pairs per buffer row -- G, or G/2 for a half tile)
    template <int GG>
    static SSW_HD void fwd_stage0(const cplx* buf, cplx* fft, int rd, int c) {
        using I = StageInfo<P, 0>;
        constexpr int R = I::R, NB = I::NB;
        const int qq = c % TEAMS, jj = c / TEAMS;   // c < NC: jj < T
        if (GG < TEAMS && qq >= GG) return;         // half tile on a 4-team CTA: two line pairs
        const int q = rd * TEAMS + qq;
        cplx* s = fft + qq * PITCH;
#pragma unroll
        for (int it = 0; it < (NB + T - 1) / T; ++it) {
            const int j = jj + it * T;
            if (j < NB) {
                cplx x[R];
                static_for<R>([&](auto rc) {
                    constexpr int r = decltype(rc)::value;
                    const int row = semi(j + r * NB);
                    x[r] = buf[row * GG + sw<GG>(row, q)];
                });
                Dft<R>::run(x);
                const int j0 = j * R, b0 = P::idx(j0);
                static_for<R>([&](auto rc) {
                    constexpr int r = decltype(rc)::value;
                    s[idx_off<P, r, (R == 16)>(b0, j0)] = x[r];
                });
            }
        }
    }

    // forward last phase: post pass -> coefficient rows (natural order) in the slots of this round's pairs
    template <int GG>
    // TRACK: the largest |coefficient| written by this call is folded into the shared-memory word *smax (bits of a non-negative float)
    static SSW_HD void fwd_post(cplx* buf, const cplx* fft, const cplx* t4, int rd, int c, float scale0, float scalen, unsigned* smax) {
        float mx = 0.f;
        constexpr int TT = GG < TEAMS ? GG : TEAMS;   // line pairs in flight per round
#pragma unroll 2
        for (int e = c; e < (N / 2 + 1) * TT; e += NC) {
            const int k = e / TT, qq = e - k * TT;
            const int kr = k ? N - k : 0;
            const cplx* s = fft + qq * PITCH;
            float xa, xb, ya, yb;
            dct2_post(s[P::idx(k)], s[P::idx(kr)], t4[k], xa, xb, ya, yb);
            const float sk = k ? scalen : scale0;
            const int q = rd * TT + qq;
            const cplx ck = cmul_lanes(mk(xa, xb), sk, sk);
            buf[k * GG + sw<GG>(k, q)] = ck;
            if constexpr (TRACK) mx = fmaxf(mx, fmaxf(fabsf(ck.x), fabsf(ck.y)));
            if (k && kr != k) {
                const cplx cr = cmul_lanes(mk(ya, yb), scalen, scalen);
                buf[kr * GG + sw<GG>(kr, q)] = cr;
                if constexpr (TRACK) mx = fmaxf(mx, fmaxf(fabsf(cr.x), fabsf(cr.y)));
            }
        }
        if constexpr (TRACK) {
#if defined(__CUDA_ARCH__)
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));   // (whole warps run the post pass)
            if ((c & 31) == 0) atomicMax(smax, __float_as_uint(mx));
#else
            unsigned b; std::memcpy(&b, &mx, 4); if (b > *smax) *smax = b;
#endif
        } else { (void)mx; (void)smax; }
    }

    // inverse phase 0: pre pass, coefficient rows of this round's pairs -> FFT buffers
    template <int GG>
    static SSW_HD void inv_pre(const cplx* buf, cplx* fft, const cplx* t4, int rd, int c) {
        constexpr int TT = GG < TEAMS ? GG : TEAMS;   // line pairs in flight per round
#pragma unroll 2
        for (int e = c; e < (N / 2 + 1) * TT; e += NC) {
            const int k = e / TT, qq = e - k * TT;
            const int kr = k ? N - k : 0;
            const int q = rd * TT + qq;
            cplx* s = fft + qq * PITCH;
            const cplx pv = buf[k * GG + sw<GG>(k, q)];
            cplx qv = mk(0.f, 0.f);
            if (k) qv = buf[kr * GG + sw<GG>(kr, q)];
            cplx zk, zr;
            dct3_pre(pv.x, pv.y, qv.x, qv.y, t4[k], zk, zr);
            s[P::idx(k)] = zk;
            if (k && kr != k) s[P::idx(kr)] = zr;
        }
    }

    // inverse last phase: FFT buffers -> sample rows (semi-Makhoul order) in the slots of this round's pairs
    template <int GG>
    static SSW_HD void inv_out(cplx* buf, const cplx* fft, int rd, int c, float scale) {
        constexpr int TT = GG < TEAMS ? GG : TEAMS;   // line pairs in flight per round
#pragma unroll 4
        for (int e = c; e < N * TT; e += NC) {
            const int n = e / TT, qq = e - n * TT;
            const int q = rd * TT + qq;
            const cplx f = fft[qq * PITCH + P::idx(semi(n))];   // row n of the buffer holds FFT position semi(n) (an involution)
            buf[n * GG + sw<GG>(n, q)] = cmul_lanes(f, scale, -scale);
        }
    }

    // all phases of one round, as the emulation and the kernel run them; `sync_all(id)` / `sync_team()` are supplied
    // by the caller (named barriers on the device, nothing on the CPU where phases run one after the other)
    template <int PH, int GG = G_>
    static SSW_HD void phase(const PipeArgs& a, cplx* buf, cplx* fft, const cplx* tw, const cplx* t4, int rd, int c, Thread& th) {
        const int g = c / T, t = c - g * T;       // team mapping of the middle stages
        cplx* s = fft + g * PITCH;
        if constexpr (!INVERSE) {
            if constexpr (PH == 0) fwd_stage0<GG>(buf, fft, rd, c);
            else if constexpr (PH == NPH_FWD - 1)
                fwd_post<GG>(buf, fft, t4, rd, c, a.scale0, a.scalen, TRACK ? (unsigned*)((unsigned char*)fft + (OFF_BAR - OFF_FFT) + SMAX_OFF) : nullptr);
            else { if (GG >= TEAMS || g < GG) fft_phase<P, PH + 2, TW_SMEM>(s, tw, t, th.v); }   // PH 1 -> fft_phase 3 (load of stage 1), ...
        } else {
            if constexpr (PH == 0) inv_pre<GG>(buf, fft, t4, rd, c);
            else if constexpr (PH == NPH_INV - 1) inv_out<GG>(buf, fft, rd, c, a.scale0);
            else { if (GG >= TEAMS || g < GG) fft_phase<P, PH, TW_SMEM>(s, tw, t, th.v); }       // PH 1 -> fft_phase 1 (load of stage 0), ...
        }
    }
    static constexpr int NPHASES = INVERSE_ ? NPH_INV : NPH_FWD;
    // true: the barrier AFTER phase PH must cover all compute threads (the next phase uses another thread mapping)
    template <int PH> struct AllAfter {
        static constexpr bool value = INVERSE_ ? (PH == 0 || PH >= NPH_INV - 2) : (PH == 0 || PH >= NPH_FWD - 2);
    };
};

// =====================================================================================================================
// Row passes as persistent, warp-specialised bulk-copy pipelines (RGB8 in / RGB8 out: the passes of the embed / extract
// step).  Same arithmetic as RowFwd / RowInv (dct_fast.cuh) -- the row half of /root/reference/src/dct2d.rs:129-170 with
// the colour conversion of /root/reference/src/yiq.rs:177-197 fused -- restructured like the column pipelines:
//   * one CTA per SM slot looping over its tiles; tile = TEAMS row pairs = 2*TEAMS adjacent rows, which are CONTIGUOUS
//     in memory on both sides, so every transfer is one linear bulk copy (cp.async.bulk, no tensor map);
//   * a PRODUCER warp (one elected thread) moves the tiles, signalled by mbarriers:
//       forward : A <- RGB8 rows of tile j+1 as soon as the conversion phase of tile j has consumed A; the post pass
//                 stores its coefficient rows straight to the plane (coalesced 128-byte warp stores -- staging them
//                 for a bulk store costs the same number of store instructions and 8N bytes of shared memory per
//                 team, i.e. a third resident CTA per SM);
//       inverse : A <- coefficient rows (consumed by the pre pass); B <- original RGB8 rows, converted IN PLACE to the
//                 output RGB8 rows by the last phase and stored from there; the next tile's originals follow the store;
//   * COMPUTE teams of P::T threads own one row pair each; they never touch global memory except for twiddles, wait only
//     on mbarriers (tile landed / buffer drained) and on their own team's named barrier between FFT stages, so the
//     teams of a CTA drift apart and the conversion, FFT and post phases of different row pairs overlap on the SM.
// Shared memory per CTA: forward A + FFT buffers (14*N bytes per team: 54 KB for one 3840-point row pair -> three CTAs
// per SM), inverse A + B + FFT buffers (22*N bytes per team -> two CTAs per SM).
// =====================================================================================================================
struct RowPipeArgs {
    int w, h, batch;
    int tiles_per_image, total_tiles;
    const unsigned char* pix;   // forward: source RGB8 frames; inverse: the original RGB8 frames (chroma)
    float* plane;               // forward: destination coefficient planes; inverse: source
    unsigned char* out;         // inverse: destination RGB8 frames
    float scale0, scalen;       // forward: factors for k == 0 / k > 0; inverse: scale0 = output scale
    const cplx* tw;
    const cplx* t4;
    int pdl_late;
    float neg_zero;             // -0.0f at run time, see FastArgs
    long long* trace;           // diagnostics (ssw_ctx_set_trace), nullptr = off
    // inverse pass after a PARTIAL inverse column pass (PipeArgs::col_limit): the columns of image i beyond
    // ((col_cut_img[i] >> col_tile_shift) + 1) << col_tile_shift were not transformed back and still hold the forward row
    // transform, which is the inverse column transform of their coefficients up to that pass pair's gain h/2 -- applied here,
    // to the coefficients as they are read (col_gain = h/2).  nullptr: every column went through the inverse column pass.
    const unsigned* col_cut_img;
    int col_tile_shift;
    float col_gain;
};

// INPLACE_ (inverse only): the coefficient rows land IN the FFT buffers (a row pair is 8N bytes, exactly the N complex values
// it becomes) and the pre pass runs in place -- every thread reads its operands into registers, the team synchronises, then
// the pre-twiddled values overwrite them.  No buffer A: 14N instead of 22N bytes per team, i.e. a third resident CTA per SM
// for the 3840-point rows (54 KB), at the price of the prefetch of A (the next tile's coefficient rows are requested when
// this tile's output phase has read the FFT buffers; the other resident CTAs cover the wait).
template <class P_, int TEAMS_, bool INVERSE_, int MINB_ = 2, bool INPLACE_ = false>
struct RowPipe {
    using P = P_;
    static constexpr int N = P_::N, T = P_::T, TEAMS = TEAMS_, ROWS = 2 * TEAMS_, MINB = MINB_;
    static constexpr bool INVERSE = INVERSE_, INPLACE = INPLACE_;
    static_assert(!INPLACE_ || INVERSE_, "the in-place pre pass belongs to the inverse pipeline");
    static constexpr int NC = TEAMS_ * P_::T, THREADS = NC + 32;
    static constexpr int PIX_ROW = 3 * N, COEF_ROW = 4 * N;              // bytes per row
    static constexpr int PIX_BYTES = ROWS * PIX_ROW, COEF_BYTES = ROWS * COEF_ROW;
    static_assert(PIX_BYTES % 16 == 0 && COEF_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
    static constexpr int A_BYTES = INVERSE_ ? COEF_BYTES : PIX_BYTES;    // input side
    static constexpr int B_BYTES = INVERSE_ ? PIX_BYTES : 0;             // inverse: originals in, output bytes out (in place)
    static constexpr int al128(int b) { return (b + 127) / 128 * 128; }
    static constexpr int OFF_B = INPLACE_ ? 0 : al128(A_BYTES), OFF_FFT = OFF_B + al128(B_BYTES);
    static constexpr int OFF_A = INPLACE_ ? OFF_FFT : 0;                  // in place: team g's rows at OFF_FFT + g * PITCH * 8
    static constexpr int FFT_BYTES = TEAMS_ * P_::PITCH * (int)sizeof(cplx);
    static constexpr int OFF_BAR = OFF_FFT + al128(FFT_BYTES);
    static constexpr int SMEM = OFF_BAR + 64;
    static constexpr bool FITS = SMEM <= (228 * 1024) / MINB_ - 1024;
    static constexpr int NPH = 2 + 2 * P_::NST + (INPLACE_ ? 1 : 0);     // in place: the pre pass is two phases (read | write)
    static constexpr int PRE_IT = (N / 2 + 1 + T - 1) / T;               // pre-pass elements per thread
    static constexpr bool INPLACE_OK = 2 * PRE_IT <= MaxRegs<P_>::value && (P_::PITCH * (int)sizeof(cplx)) % 16 == 0 && P_::PITCH >= N;
    using Thread = ThreadState<P_>;
    static bool supports(int w, int h) { return w == N && h > 0 && h % ROWS == 0; }
    static int tiles_per_image(int w, int h) { (void)w; return h / ROWS; }

    // c: compute thread id (team g = c / T owns rows 2g, 2g+1 of the tile); gout: the tile's coefficient rows in the plane (forward)
    // kcut (inverse): first column whose coefficients still need the gain of the skipped column passes (RowPipeArgs::col_cut_img)
    template <int PH>
    static SSW_HD void phase(const RowPipeArgs& a, unsigned char* bufA, cplx* fft, unsigned char* bufB, float* gout, int c, Thread& th, int kcut = N + 1) {
        const int g = c / T, t = c - g * T;
        cplx* s = fft + g * P::PITCH;
        if constexpr (INPLACE && PH == 0) {
            // in-place pre pass, first half: the operands of this thread's elements -> registers (the coefficient rows of the
            // team lie where its FFT buffer is: row a = floats [0, N), row b = floats [N, 2N))
            const float* ia = (const float*)s;
            const float* ib = ia + N;
#pragma unroll
            for (int it = 0; it < PRE_IT; ++it) {
                const int k = t + it * T;
                if (k <= N / 2) {
                    const int kr = k ? N - k : 0;
                    const float gk = k >= kcut ? a.col_gain : 1.f, gr = kr >= kcut ? a.col_gain : 1.f;
                    th.v[2 * it] = mk(SSW_FMUL(ia[k], gk), SSW_FMUL(ib[k], gk));
                    th.v[2 * it + 1] = k ? mk(SSW_FMUL(ia[kr], gr), SSW_FMUL(ib[kr], gr)) : mk(0.f, 0.f);
                }
            }
        } else if constexpr (INPLACE && PH == 1) {
            // second half (every operand of the team is in registers): pre-twiddled values -> the same buffer
#pragma unroll
            for (int it = 0; it < PRE_IT; ++it) {
                const int k = t + it * T;
                if (k <= N / 2) {
                    const int kr = k ? N - k : 0;
                    const cplx pv = th.v[2 * it], qv = th.v[2 * it + 1];
                    cplx zk, zr;
                    dct3_pre(pv.x, pv.y, qv.x, qv.y, SSW_LDG(&a.t4[k]), zk, zr);
                    s[P::idx(k)] = zk;
                    if (k && kr != k) s[P::idx(kr)] = zr;
                }
            }
        } else if constexpr (PH > (INPLACE ? 1 : 0) && PH < NPH - 1) {
            fft_phase<P, PH - (INPLACE ? 1 : 0)>(s, a.tw, t, th.v);
        } else if constexpr (!INVERSE && PH == 0) {
            // staged RGB8 bytes -> luma -> Makhoul-ordered FFT input (RowFwd phase 0 without the global loads)
            const unsigned* sa = (const unsigned*)(bufA + (2 * g) * PIX_ROW);
            const unsigned* sb = (const unsigned*)(bufA + (2 * g + 1) * PIX_ROW);
#pragma unroll
            for (int it = 0; it < (N / 4 + T - 1) / T; ++it) {
                const int u = t + it * T;
                if (u < N / 4) {
                    const unsigned wa[3] = {sa[3 * u], sa[3 * u + 1], sa[3 * u + 2]}, wb[3] = {sb[3 * u], sb[3 * u + 1], sb[3 * u + 2]};
                    cplx y2[4];
                    luma4x2_words(wa, wb, a.neg_zero, y2);
                    put4x2<P>(s, u, y2);
                }
            }
        } else if constexpr (!INVERSE) {
            // DCT-II post pass -> coefficient rows of the plane (RowFwd last phase)
            float* oa = gout + (size_t)(2 * g) * N;
            float* ob = oa + N;
#pragma unroll 4
            for (int k = t; k <= N / 2; k += T) {
                const int kr = k ? N - k : 0;
                float xa, xb, ya, yb;
                dct2_post(s[P::idx(k)], s[P::idx(kr)], SSW_LDG(&a.t4[k]), xa, xb, ya, yb);
                const float sk = k ? a.scalen : a.scale0;
                oa[k] = xa * sk;
                ob[k] = xb * sk;
                if (k && kr != k) {
                    oa[kr] = ya * a.scalen;
                    ob[kr] = yb * a.scalen;
                }
            }
        } else if constexpr (PH == 0) {
            // DCT-III pre pass from the staged coefficient rows (RowInv phase 0)
            const float* ia = (const float*)(bufA + (2 * g) * COEF_ROW);
            const float* ib = ia + N;
#pragma unroll 4
            for (int k = t; k <= N / 2; k += T) {
                const int kr = k ? N - k : 0;
                const float gk = k >= kcut ? a.col_gain : 1.f, gr = kr >= kcut ? a.col_gain : 1.f;
                const float pa = SSW_FMUL(ia[k], gk), pb = SSW_FMUL(ib[k], gk);
                const float qa = k ? SSW_FMUL(ia[kr], gr) : 0.f, qb = k ? SSW_FMUL(ib[kr], gr) : 0.f;
                cplx zk, zr;
                dct3_pre(pa, pb, qa, qb, SSW_LDG(&a.t4[k]), zk, zr);
                s[P::idx(k)] = zk;
                if (k && kr != k) s[P::idx(kr)] = zr;
            }
        } else {
            // FFT output -> new luma; + chroma of the original bytes in B -> output bytes, in place (RowInv last phase)
            unsigned* wa = (unsigned*)(bufB + (2 * g) * PIX_ROW);
            unsigned* wb = (unsigned*)(bufB + (2 * g + 1) * PIX_ROW);
#pragma unroll
            for (int it = 0; it < (N / 4 + T - 1) / T; ++it) {
                const int u = t + it * T;
                if (u < N / 4) {
                    cplx f[4];
                    get4<P>(s, u, f);
                    cplx y2[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) y2[i] = cmul_lanes_x(f[i], a.scale0, -a.scale0, a.neg_zero);   // feeds the colour adds: no contraction
                    const unsigned ia[3] = {wa[3 * u], wa[3 * u + 1], wa[3 * u + 2]}, ib[3] = {wb[3 * u], wb[3 * u + 1], wb[3 * u + 2]};
                    unsigned oa[3], ob[3];
                    rgb8_out4x2_words(ia, ib, a.neg_zero, y2, oa, ob);
#pragma unroll
                    for (int j = 0; j < 3; ++j) { wa[3 * u + j] = oa[j]; wb[3 * u + j] = ob[j]; }
                }
            }
        }
    }
};

// default shape of the row pipelines of a plan: ~256 compute threads per CTA, two CTAs per SM when they fit
template <class P> struct RowPipeCfg {
    static constexpr int TEAMS = P::T >= 192 ? 1 : (P::T >= 96 ? 2 : 4);
    static constexpr int MINB_I = RowPipe<P, TEAMS, true, 2>::FITS ? 2 : 1;
    static constexpr int MINB_F = (RowPipe<P, TEAMS, false, 3>::FITS && 3 * (TEAMS * P::T + 32) * 72 <= 65536) ? 3 : MINB_I;
    static constexpr bool OK = !P::PAD && RowPipe<P, TEAMS, false, MINB_F>::FITS && RowPipe<P, TEAMS, true, MINB_I>::FITS &&
                               (TEAMS * P::T + 32) <= 1024;
    using Fwd = RowPipe<P, TEAMS, false, MINB_F>;
    using Inv = RowPipe<P, TEAMS, true, MINB_I>;
    // in-place pre pass: worth it where it buys one more resident CTA (and the registers of that many threads exist)
    static constexpr int MINB_P = MINB_I + 1;
    using InvPTry = RowPipe<P, TEAMS, true, MINB_P, true>;
    static constexpr bool INPLACE_OK = OK && InvPTry::INPLACE_OK && InvPTry::FITS && !RowPipe<P, TEAMS, true, MINB_P>::FITS &&
                                       (TEAMS * P::T + 32) * MINB_P * 72 <= 65536;
    using InvP = std::conditional_t<INPLACE_OK, InvPTry, Inv>;   // == Inv where the in-place shape does not apply
};

#if defined(__CUDACC__)
// ---- PTX wrappers: mbarrier, TMA, named barriers ---------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol error traps (reported as a launch failure) instead of hanging the device
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"   // suspended (no issue slots) until the phase completes or the hint (ns) expires
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity), "r"(100000u) : "memory");
        if (!done && spin > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void tma_load_4d(unsigned dst, const void* map, unsigned bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const void* map, unsigned bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* map, unsigned src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* map, unsigned src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- pipeline time line (diagnostics; one uniform branch per event when off) ------------------------------------------
// 64 slots of 8 bytes per CTA: 0 %globaltimer at start, 1 clock64 at start, 2 clock64 after the dependency wait, 3 clock64 when
// the compute warps have finished, 4 %globaltimer then, 5 tiles of this CTA, 6 %smid;  tile j < 7 at 8 + 8j:
//   +0 load issued (producer)   +1 tile landed (compute)   +2 compute done   +3 store issued (producer)
//   +4 store has read the buffer (producer)   +5 row pipes: input buffer A released   +6 row pipes (inverse): originals landed
#if defined(SSW_TRACE)
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void trace_at(long long* trace, int slot) { if (trace) trace[blockIdx.x * 64 + slot] = clock64(); }
__device__ __forceinline__ void trace_tile(long long* trace, int j, int what) { if (trace && j < 7) trace[blockIdx.x * 64 + 8 + 8 * j + what] = clock64(); }
__device__ __forceinline__ void trace_begin(long long* trace) { if (trace) { trace[blockIdx.x * 64] = (long long)global_ns(); trace[blockIdx.x * 64 + 1] = clock64(); } }
__device__ __forceinline__ void trace_info(long long* trace, int nt) {
    if (trace) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); trace[blockIdx.x * 64 + 5] = nt; trace[blockIdx.x * 64 + 6] = smid; }
}
__device__ __forceinline__ void trace_end(long long* trace) { if (trace) { trace[blockIdx.x * 64 + 3] = clock64(); trace[blockIdx.x * 64 + 4] = (long long)global_ns(); } }
#else   // the stamps cost registers and issue slots in the tile loops (measured: fwd_cols 30.3 -> 33.9 us): diagnostic builds only (-DSSW_TRACE)
__device__ __forceinline__ void trace_at(long long*, int) {}
__device__ __forceinline__ void trace_tile(long long*, int, int) {}
__device__ __forceinline__ void trace_begin(long long*) {}
__device__ __forceinline__ void trace_info(long long*, int) {}
__device__ __forceinline__ void trace_end(long long*) {}
#endif

struct alignas(64) TmaMap { unsigned long long v[16]; };   // CUtensorMap (128 bytes, 64-byte aligned)

// ---- linear bulk copies (cp.async.bulk, no tensor map): global -> shared signals an mbarrier, shared -> global joins a bulk group
__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

// map_s: sample side  (4-D: column, parity, row pair, image);  map_c: coefficient side (3-D: column, row, image);
// map_s2 / map_c2: the same views with boxes of G columns (half tiles, PipeArgs::half_tiles)
template <class K>
__global__ void __launch_bounds__(K::THREADS, K::MINB)
col_pipe_kernel(const __grid_constant__ PipeArgs a, const __grid_constant__ TmaMap map_s, const __grid_constant__ TmaMap map_c,
                const __grid_constant__ TmaMap map_s2, const __grid_constant__ TmaMap map_c2) {
    extern __shared__ __align__(1024) unsigned char pipe_smem[];
    constexpr int NC = K::NC;
    const int tid = threadIdx.x;
    const unsigned sbase = smem_u32(pipe_smem);
    const unsigned bar_full0 = sbase + K::OFF_BAR, bar_ready0 = bar_full0 + 16;   // full[2], ready[2]
    cplx* fft = (cplx*)(pipe_smem + K::OFF_FFT);
    const cplx* tw = a.tw;
    const cplx* t4 = a.t4;

    const unsigned bar_tab = bar_full0 + 32;
    if (tid == 0) {
        mbar_init(bar_full0, 1); mbar_init(bar_full0 + 8, 1);
        mbar_init(bar_ready0, 1); mbar_init(bar_ready0 + 8, 1);
        mbar_init(bar_tab, 1);
        if constexpr (K::TRACK) *(unsigned*)(pipe_smem + K::OFF_BAR + K::SMAX_OFF) = 0u;
        fence_mbar_init();
    }
    if (tid == 0) trace_begin(a.trace);
    if (!a.pdl_late) pdl_trigger();
    // twiddle tables (read-only, written long before the previous kernel) go to shared memory: as bulk copies issued by the
    // producer thread that land beside the first tile (a loop of all threads costs ~3 us in front of every launch: nothing
    // of the previous grid is left to overlap it with once this CTA has found room on an SM)
    constexpr bool TAB = K::TW_SMEM || K::T4_SMEM;
    if constexpr (K::TW_SMEM) tw = (const cplx*)(pipe_smem + K::OFF_TW);
    if constexpr (K::T4_SMEM) t4 = (const cplx*)(pipe_smem + K::OFF_T4);
    if (TAB && !a.tab_bulk) {
        if constexpr (K::TW_SMEM) { cplx* d = (cplx*)(pipe_smem + K::OFF_TW); for (int i = tid; i < K::P::TW_TOTAL; i += K::THREADS) d[i] = __ldg(a.tw + i); }
        if constexpr (K::T4_SMEM) { cplx* d = (cplx*)(pipe_smem + K::OFF_T4); for (int i = tid; i < K::N / 2 + 1; i += K::THREADS) d[i] = __ldg(a.t4 + i); }
    }
    __syncthreads();
    if (TAB && a.tab_bulk && tid == NC) {
        mbar_expect_tx(bar_tab, (K::TW_SMEM ? K::TW_BYTES : 0) + (K::T4_SMEM ? K::T4_BYTES : 0));
        if constexpr (K::TW_SMEM) bulk_load(sbase + K::OFF_TW, a.tw, K::TW_BYTES, bar_tab);
        if constexpr (K::T4_SMEM) bulk_load(sbase + K::OFF_T4, a.t4, K::T4_BYTES, bar_tab);
    }
    pdl_wait();
    if (tid == 0) trace_at(a.trace, 2);

    // this CTA's sequence: its whole tiles (round-robin), then at most one half tile
    const int first = blockIdx.x, step = gridDim.x;
    int tiles_per_image = a.tiles_per_image, full_tiles = a.full_tiles, half_tiles = a.half_tiles;
    if constexpr (K::INVERSE) {
        // partial inverse (PipeArgs::col_limit): only the columns 0 .. *col_limit hold coefficients that differ from the forward
        // transform's; the tile schedule shrinks to them (same rule as the host's: whole rounds of whole tiles, a remainder
        // that keeps at most half of the CTAs busy goes out as half tiles)
        if (a.col_limit) {
            const int lim = min(a.tiles_per_image, (int)(__ldcg(a.col_limit) / (unsigned)(2 * K::G)) + 1);
            const int total = lim * a.batch, rounds = total / step, rem = total - rounds * step;
            tiles_per_image = lim; full_tiles = total; half_tiles = 0;
            if (K::HALF_OK && a.batch == 1 && rem > 0 && 2 * rem <= step && (a.w % (2 * K::G)) == 0) { full_tiles = rounds * step; half_tiles = 2 * rem; }
        }
    }
    int nt_full = first < full_tiles ? (full_tiles - first + step - 1) / step : 0;
    const int half_idx = first - (step - half_tiles);         // the LAST half_tiles CTAs take one (the histogram tiles live on the first)
    bool has_half = K::HALF_OK && half_tiles > 0 && half_idx >= 0;
    int half_tile = full_tiles + (half_idx >> 1), half_sub = half_idx & 1;
    if constexpr (!K::INVERSE && K::HALF_OK) {
        // relief for the CTAs that build the histogram (PipeArgs::hist_relief): CTA c < relief hands its last whole tile, as two
        // half tiles, to the CTAs relief + 2c and relief + 2c + 1, which have neither a histogram nor a half tile of their own
        const int relief = a.hist_relief;
        if (first < relief) nt_full -= 1;
        else if (first < 3 * relief) {
            const int h = first - relief, owner = h >> 1;
            has_half = true;
            half_tile = owner + ((full_tiles - owner + step - 1) / step - 1) * step;
            half_sub = h & 1;
        }
    }
    const int nt = nt_full + (has_half ? 1 : 0);
    if (tid == 0) trace_info(a.trace, nt);
    // j-th element of the sequence -> image, first column, half tile?
    auto tile_at = [&](int j, int& img, int& c0) __attribute__((always_inline)) -> bool {
        if (j < nt_full) {
            int t = first + j * step - a.tile_rot;
            if (t < 0) t += full_tiles;
            img = t / tiles_per_image;
            c0 = (t - img * tiles_per_image) * 2 * K::G;
            return false;
        }
        const int t = half_tile;
        img = t / tiles_per_image;
        c0 = (t - img * tiles_per_image) * 2 * K::G + half_sub * K::G;
        return true;
    };

    if (tid >= NC) {
        // ===================== producer warp =====================
        if (tid == NC) {
            auto issue_load = [&](int j) __attribute__((always_inline)) {
                int img, c0;
                const bool half = tile_at(j, img, c0) && K::HALF_OK;
                const int b = j & 1;
                const unsigned dst = sbase + b * K::BUF_BYTES, bar = bar_full0 + 8 * b;
                const unsigned rowb = half ? K::ROWB / 2 : K::ROWB;
                trace_tile(a.trace, j, 0);
                mbar_expect_tx(bar, half ? K::BUF_BYTES / 2 : K::BUF_BYTES);
                if constexpr (!K::INVERSE) {
                    const TmaMap* m = half ? &map_s2 : &map_s;
                    for (int par = 0; par < 2; ++par)
                        for (int bx = 0; bx < K::NBOX_HALF; ++bx)
                            tma_load_4d(dst + (par * (K::N / 2) + bx * K::RB_HALF) * rowb, m, bar, c0, par, bx * K::RB_HALF, img);
                } else {
                    const TmaMap* m = half ? &map_c2 : &map_c;
                    for (int bx = 0; bx < K::NBOX_FULL; ++bx)
                        tma_load_3d(dst + bx * K::RB_FULL * rowb, m, bar, c0, bx * K::RB_FULL, img);
                }
            };
            auto issue_store = [&](int j) __attribute__((always_inline)) {
                int img, c0;
                const bool half = tile_at(j, img, c0) && K::HALF_OK;
                const int b = j & 1;
                const unsigned src = sbase + b * K::BUF_BYTES;
                const unsigned rowb = half ? K::ROWB / 2 : K::ROWB;
                if constexpr (K::INVERSE) {
                    // partial inverse of a batch: the schedule follows the widest image; every image keeps exactly its own columns
                    if (a.col_limit_img && (unsigned)(c0 & ~(2 * K::G - 1)) > __ldcg(a.col_limit_img + img)) { tma_commit(); return; }
                }
                if constexpr (!K::INVERSE) {
                    const TmaMap* m = half ? &map_c2 : &map_c;
                    for (int bx = 0; bx < K::NBOX_FULL; ++bx)
                        tma_store_3d(m, src + bx * K::RB_FULL * rowb, c0, bx * K::RB_FULL, img);
                } else {
                    const TmaMap* m = half ? &map_s2 : &map_s;
                    for (int par = 0; par < 2; ++par)
                        for (int bx = 0; bx < K::NBOX_HALF; ++bx)
                            tma_store_4d(m, src + (par * (K::N / 2) + bx * K::RB_HALF) * rowb, c0, par, bx * K::RB_HALF, img);
                }
                tma_commit();
            };
            if (nt > 0) issue_load(0);
            if (nt > 1) issue_load(1);
            for (int j = 0; j < nt; ++j) {
                const int b = j & 1;
                mbar_wait(bar_ready0 + 8 * b, (j >> 1) & 1);   // the compute warps have finished tile j (results in BUF[b])
                trace_tile(a.trace, j, 3);
                issue_store(j);
                if (j + 2 < nt) {
                    tma_wait_read0();                          // the store has read BUF[b]: it may be overwritten
                    trace_tile(a.trace, j, 4);
                    issue_load(j + 2);
                }
            }
            tma_wait_all0();                                   // all stores complete before the CTA exits
        }
        return;
    }

    // ===================== compute warps =====================
    typename K::Thread th;
    const int team = tid / K::T;
    if (TAB && a.tab_bulk) mbar_wait(bar_tab, 0);              // twiddle tables have landed
    for (int j = 0; j < nt; ++j) {
        const int b = j & 1;
        cplx* buf = (cplx*)(pipe_smem + b * K::BUF_BYTES);
        int img, c0;
        const bool half = tile_at(j, img, c0) && K::HALF_OK;   // (uniform over the CTA)
        mbar_wait(bar_full0 + 8 * b, (j >> 1) & 1);            // tile j has landed in BUF[b]
        if (tid == 0) trace_tile(a.trace, j, 1);
        if (!half) {
#pragma unroll 1
            for (int rd = 0; rd < K::ROUNDS; ++rd) {
                static_for<K::NPHASES>([&](auto ph) __attribute__((always_inline)) {
                    constexpr int p = decltype(ph)::value;
                    if constexpr (p == K::NPHASES - 1) { if (a.pdl_late && j + 1 == nt && rd + 1 == K::ROUNDS) pdl_trigger(); }
                    K::template phase<p>(a, buf, fft, tw, t4, rd, tid, th);
                    if constexpr (p + 1 < K::NPHASES) {
                        if constexpr (K::template AllAfter<p>::value) named_sync(1, NC);
                        else named_sync(2 + team, K::T);
                    }
                });
                if (rd + 1 < K::ROUNDS) named_sync(1, NC);     // FFT buffers are free for the next round
            }
        } else if constexpr (K::HALF_OK) {
            static_for<K::NPHASES>([&](auto ph) __attribute__((always_inline)) {              // one round over the G/2 line pairs of a half tile
                constexpr int p = decltype(ph)::value;
                if constexpr (p == K::NPHASES - 1) { if (a.pdl_late && j + 1 == nt) pdl_trigger(); }
                K::template phase<p, K::GH>(a, buf, fft, tw, t4, 0, tid, th);
                if constexpr (p + 1 < K::NPHASES) {
                    if constexpr (K::template AllAfter<p>::value) named_sync(1, NC);
                    else named_sync(2 + team, K::T);
                }
            });
        }
        if constexpr (!K::INVERSE) {
            if (a.ts.hist) {
                const int cols = (K::HALF_OK && half) ? K::G : 2 * K::G;   // floats per buffer row
                unsigned* sh = (unsigned*)fft;                 // 4096 bins + scratch, free once the post passes are done
                static_assert(K::FFT_BYTES >= (kHistBins + 40) * 4, "the FFT buffers double as the histogram");
                if (c0 < a.hist_cols) {                        // (uniform over the CTA)
                    named_sync(1, NC);                         // the post pass of every team has written its coefficient rows
                    for (int i = tid; i < kHistBins; i += NC) sh[i] = 0u;
                    named_sync(1, NC);
                    const float* cf = (const float*)buf;       // row k of the tile: `cols` adjacent columns starting at c0
                    for (int e = tid; e < a.hist_rows * cols; e += NC) {
                        const int r = e / cols, cc = e - r * cols;
                        const unsigned pidx = (unsigned)r * (unsigned)a.w + (unsigned)(c0 + cc);
                        const int ep = (K::SWZ && !half) ? (e ^ (r & 4)) : e;   // swizzled whole tiles: the 16-byte halves of rows 4..7 (mod 8) are swapped
                        if (pidx && c0 + cc < a.hist_cols) atomicAdd(sh + (order_key(cf[ep], pidx, a.oc) >> (32 - kHistBits)), 1u);
                    }
                    named_sync(1, NC);
                    unsigned* gh = a.ts.hist + (size_t)img * kHistBins;
                    for (int i = tid; i < kHistBins; i += NC)
                        if (sh[i]) atomicAdd(gh + i, sh[i]);
                    __threadfence();
                    named_sync(1, NC);
                    const unsigned n_hist_tiles = (unsigned)((a.hist_cols + 2 * K::G - 1) / (2 * K::G));
                    if (tid == 0) sh[kHistBins + 36] = (atomicAdd(a.ts.ticket + img, 1u) == n_hist_tiles - 1u) ? 1u : 0u;
                    named_sync(1, NC);
                    if (sh[kHistBins + 36]) {                  // last histogram tile of this image: the block is complete
                        __threadfence();
                        for (int i = tid; i < kHistBins; i += NC) { sh[i] = __ldcg(gh + i); gh[i] = 0u; }
                        named_sync(1, NC);
                        const unsigned bsel = find_kth_bin_team(sh, a.hist_k, tid, NC, 1, sh + kHistBins);
                        if (tid == 0) {
                            a.ts.ticket[img] = 0u;
                            __threadfence();
                            *(volatile unsigned*)(a.ts.sel_bin + img) = a.collect ? (bsel | 0x80000000u) : bsel;
                        }
                    }
                    named_sync(1, NC);                         // the FFT buffers go back to the next tile
                }
                if (K::COLLECT_OK && a.collect) {
                    named_sync(1, NC);                         // every team's coefficient rows are in BUF[b]; the FFT buffers are free
                    if (tid == 0) {
                        unsigned v = 0;
                        for (unsigned spin = 0; ; ++spin) {
                            v = *(volatile const unsigned*)(a.ts.sel_bin + img);
                            if (v & 0x80000000u) break;
                            if (spin > (1u << 22)) __trap();   // protocol error: report a launch failure instead of hanging
                            __nanosleep(40);
                        }
                        sh[0] = v & 0x7FFFFFFFu;
                    }
                    named_sync(1, NC);
                    const unsigned bin_sel = sh[0];
                    unsigned* count = a.ts.cand_count + img;
                    unsigned long long* cand = a.ts.cand + (size_t)img * kTopkCap;
                    // Energy ordering: key = bits(c*c) | 2^31, so "bin(key) >= bin_sel" is "c*c >= thr" for the float whose bits are
                    // (bin_sel - 2^11) << 20 -- two instructions per coefficient; the few that pass (and NaNs) take the exact path
                    const float thr = bin_sel > (1u << (kHistBits - 1)) ? __uint_as_float((bin_sel - (1u << (kHistBits - 1))) << (32 - kHistBits)) : 0.f;
                    const float4* c4 = (const float4*)buf;
                    const int n4 = K::N * cols / 4;
                    const int shift = half ? (K::G == 4 ? 2 : 1) : (K::G == 4 ? 3 : 2);   // log2(cols), G = 2 or 4
#pragma unroll 4
                    for (int e = tid; e < n4; e += NC) {
                        const float4 v = c4[e];
                        const bool lo = (v.x * v.x < thr) & (v.y * v.y < thr) & (v.z * v.z < thr) & (v.w * v.w < thr);
                        if (!lo) {
                            const int f = e * 4, r = f >> shift, cp = f - (r << shift);
                            const int cc = (K::SWZ && !half) ? (cp ^ (r & 4)) : cp;   // (un-swizzle: see the histogram above)
                            const unsigned p0 = (unsigned)r * (unsigned)a.w + (unsigned)(c0 + cc);
                            const float ev[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (p0 + u && c0 + cc + u < a.w) topk_push(order_key(ev[u], p0 + u, a.oc), p0 + u, bin_sel, count, cand);
                        }
                    }
                }
            }
        }
        fence_proxy_async();                                   // generic-proxy writes of BUF[b] -> visible to the TMA store
        named_sync(1, NC);                                     // (also: FFT buffers free for the next tile, candidates read)
        if (tid == 0) { trace_tile(a.trace, j, 2); mbar_arrive(bar_ready0 + 8 * b); }
        if constexpr (K::TRACK) {
            if (tid == 0) {                                    // every post pass of the tile has folded its maximum into the word
                unsigned* smax = (unsigned*)(pipe_smem + K::OFF_BAR + K::SMAX_OFF);
                if (a.tile_max) atomicMax(a.tile_max + (size_t)img * a.tiles_per_image + c0 / (2 * K::G), *smax);
                *smax = 0u;                                    // (the next tile's post pass is several barriers away)
            }
        }
    }
    if (tid == 0) trace_end(a.trace);
}

// mbarriers (8 bytes each, at OFF_BAR): fullA, freeA; inverse also fullB (originals landed) and readyB (output bytes written)
template <class K>
__global__ void __launch_bounds__(K::THREADS, K::MINB) row_pipe_kernel(const __grid_constant__ RowPipeArgs a) {
    extern __shared__ __align__(1024) unsigned char pipe_smem[];
    constexpr int NC = K::NC;
    const int tid = threadIdx.x;
    const unsigned sbase = smem_u32(pipe_smem);
    const unsigned bar_fullA = sbase + K::OFF_BAR, bar_freeA = bar_fullA + 8, bar_fullB = bar_fullA + 16, bar_readyB = bar_fullA + 24;
    if (tid == 0) {
        mbar_init(bar_fullA, 1);      // producer's expect_tx arrival + the bytes of the copy
        mbar_init(bar_freeA, NC);     // every compute thread has consumed A
        mbar_init(bar_fullB, 1);      // inverse: expect_tx of the originals
        mbar_init(bar_readyB, NC);    // inverse: every compute thread has written (and fenced) its part of B
        fence_mbar_init();
    }
    if (tid == 0) trace_begin(a.trace);
    if (!a.pdl_late) pdl_trigger();
    __syncthreads();
    pdl_wait();
    if (tid == 0) trace_at(a.trace, 2);

    const int first = blockIdx.x, step = gridDim.x;
    const int nt = first < a.total_tiles ? (a.total_tiles - first + step - 1) / step : 0;
    if (tid == 0) trace_info(a.trace, nt);
    const size_t frame_px = (size_t)a.w * a.h;
    auto px_of = [&](int j) __attribute__((always_inline)) {   // first pixel of tile j of this CTA, counted over the whole batch
        const int tile = first + j * step, img = tile / a.tiles_per_image;
        return img * frame_px + (size_t)(tile - img * a.tiles_per_image) * K::ROWS * K::N;
    };

    if (tid >= NC) {
        // ===================== producer warp =====================
        if (tid == NC) {
            auto load_a = [&](int j) __attribute__((always_inline)) {
                trace_tile(a.trace, j, 0);
                mbar_expect_tx(bar_fullA, K::A_BYTES);
                if constexpr (K::INPLACE) {   // one row pair per team, straight into its FFT buffer
                    for (int g = 0; g < K::TEAMS; ++g)
                        bulk_load(sbase + K::OFF_FFT + g * K::P::PITCH * (int)sizeof(cplx), a.plane + px_of(j) + (size_t)g * 2 * K::N, 2 * K::COEF_ROW, bar_fullA);
                } else if constexpr (K::INVERSE) bulk_load(sbase + K::OFF_A, a.plane + px_of(j), K::A_BYTES, bar_fullA);
                else bulk_load(sbase + K::OFF_A, a.pix + 3 * px_of(j), K::A_BYTES, bar_fullA);
            };
            auto load_b = [&](int j) __attribute__((always_inline)) {   // inverse only: the original pixels of the tile
                mbar_expect_tx(bar_fullB, K::B_BYTES);
                bulk_load(sbase + K::OFF_B, a.pix + 3 * px_of(j), K::B_BYTES, bar_fullB);
            };
            if (nt > 0) {
                load_a(0);
                if constexpr (K::INVERSE) load_b(0);
            }
            for (int j = 0; j < nt; ++j) {
                if (j + 1 < nt) {
                    mbar_wait(bar_freeA, j & 1);               // tile j has left A
                    load_a(j + 1);
                }
                if constexpr (K::INVERSE) {
                    mbar_wait(bar_readyB, j & 1);              // output bytes of tile j are in B (writers fenced)
                    trace_tile(a.trace, j, 3);
                    bulk_store(a.out + 3 * px_of(j), sbase + K::OFF_B, K::B_BYTES);
                    tma_commit();
                    if (j + 1 < nt) {
                        tma_wait_read0();                      // the store has read B
                        trace_tile(a.trace, j, 4);
                        load_b(j + 1);
                    }
                }
            }
            if constexpr (K::INVERSE) tma_wait_all0();         // all stores complete before the CTA exits
        }
        return;
    }

    // ===================== compute teams =====================
    typename K::Thread th;
    const int team = tid / K::T;
    cplx* fft = (cplx*)(pipe_smem + K::OFF_FFT);
    for (int j = 0; j < nt; ++j) {
        float* gout = K::INVERSE ? nullptr : a.plane + px_of(j);
        int kcut = K::N + 1;
        if constexpr (K::INVERSE) {
            if (a.col_cut_img) {
                const int img = (first + j * step) / a.tiles_per_image;
                kcut = (int)(((__ldcg(a.col_cut_img + img) >> a.col_tile_shift) + 1u) << a.col_tile_shift);
            }
        }
        mbar_wait(bar_fullA, j & 1);                           // tile j has landed in A
        if (tid == 0) trace_tile(a.trace, j, 1);
        static_for<K::NPH>([&](auto ph) __attribute__((always_inline)) {
            constexpr int p = decltype(ph)::value;
            if constexpr (p == K::NPH - 1) {
                if (a.pdl_late && j + 1 == nt) pdl_trigger();
                if constexpr (K::INVERSE) { mbar_wait(bar_fullB, j & 1); if (tid == 0) trace_tile(a.trace, j, 6); }   // the originals of tile j have landed in B
            }
            K::template phase<p>(a, pipe_smem + K::OFF_A, fft, pipe_smem + K::OFF_B, gout, tid, th, kcut);
            if constexpr (p == 0 && !K::INPLACE) { mbar_arrive(bar_freeA); if (tid == 0) trace_tile(a.trace, j, 5); }   // (this thread's) reads of A are done
            if constexpr (p + 1 < K::NPH) named_sync(1 + team, K::T);
        });
        if constexpr (K::INVERSE) {
            fence_proxy_async();                               // generic-proxy writes of B -> visible to the bulk store
            mbar_arrive(bar_readyB);
            // in place, A is the FFT buffer: busy until the output phase has read it (the fence above also orders this
            // thread's FFT-stage writes before the bulk copy that overwrites them)
            if constexpr (K::INPLACE) { mbar_arrive(bar_freeA); if (tid == 0) trace_tile(a.trace, j, 5); }
        }
        named_sync(1 + team, K::T);                            // the team has read its FFT buffer: the next tile may overwrite it
        if (tid == 0) trace_tile(a.trace, j, 2);
    }
    if (tid == 0) trace_end(a.trace);
}
#endif

}  // namespace fast
}  // namespace ssw
