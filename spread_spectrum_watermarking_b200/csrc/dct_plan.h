// Host-side planning for the shared-memory FFT-DCT line kernels (pure C++, no CUDA types).
//
// Replaces `rustdct::DctPlanner::plan_dct2/plan_dct3` as used by the reference's 2-D driver
// (/root/reference/src/dct2d.rs:119-123): a plan is the mixed-radix factorisation of the line
// length N, the twiddle tables (computed in double, stored as f32) and the thread shape.
//
// Algorithm (see DESIGN.md "DCT line kernels"): two real lines A,B are packed as one complex
// sequence z = perm(A) + i*perm(B) (Makhoul even/odd reordering), one N-point complex Stockham
// FFT runs in shared memory, and a twiddle post-pass separates the two spectra into the two DCTs.
#pragma once
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

namespace ssw {

constexpr int kMaxStages = 16;
constexpr int kGenericOutMax = 8;     // outputs per thread held in registers by the generic-prime stage
constexpr int kMaxThreadsPerLinePair = 1024;

struct DctStage {
    int radix;      // butterfly size
    int ns;         // product of the radices of the previous stages (Stockham "Ns")
    int generic;    // 1: radix is a prime > 5 handled by the O(p^2) stage
    int tw_offset;  // offset (in float2) of this stage's table inside stage_tw: layout [r-1][k], k < ns
};

struct DctPlanHost {
    int n = 0;
    int npad = 0;  // padded shared-memory length (float2 units) of one line pair
    int tp = 0;    // threads cooperating on one line pair
    int nstages = 0;
    DctStage stages[kMaxStages];
    std::vector<float> stage_tw;  // float2 pairs: exp(-2*pi*i*k*r/(ns*radix))
    std::vector<float> wn;        // float2 pairs: exp(-2*pi*i*j/n), j < n      (generic stage)
    std::vector<float> t4;        // float2 pairs: exp(-i*pi*k/(2n)), k < n     (DCT post/pre twiddle)
    std::string error;
};

inline int pad_index(int a) { return a + (a >> 5); }

// registers budget: a thread holds at most 16 complex values across the in-place barrier
inline int stage_max_iter(int radix) { return radix >= 9 ? 1 : 16 / radix; }

inline bool is_register_radix(int r) {
    switch (r) {
        case 2: case 3: case 4: case 5: case 6: case 8: case 9: case 10: case 12: case 15: case 16:
            return true;
        default:
            return false;
    }
}

namespace detail {
// choose register radices for the {2,3,5}-smooth part: minimise stage count, then prefer big radices
inline void best_split(int n, std::vector<int>& cur, std::vector<int>& best) {
    static const int cand[] = {16, 15, 12, 10, 9, 8, 6, 5, 4, 3, 2};
    if (n == 1) {
        if (best.empty() || cur.size() < best.size()) best = cur;
        return;
    }
    if (!best.empty() && cur.size() + 1 > best.size()) return;
    for (int c : cand) {
        if (n % c) continue;
        if (!cur.empty() && c > cur.back()) continue;  // non-increasing: canonical order
        cur.push_back(c);
        best_split(n / c, cur, best);
        cur.pop_back();
    }
}
}  // namespace detail

inline DctPlanHost make_dct_plan(int n) {
    DctPlanHost p;
    p.n = n;
    if (n < 1) { p.error = "line length must be >= 1"; return p; }
    // factor out 2,3,5; everything else is a generic prime stage
    int smooth = 1, rest = n;
    for (int f : {2, 3, 5}) while (rest % f == 0) { rest /= f; smooth *= f; }
    std::vector<int> generic;
    for (int f = 7; (int64_t)f * f <= rest; f += 2) while (rest % f == 0) { generic.push_back(f); rest /= f; }
    if (rest > 1) generic.push_back(rest);
    std::vector<int> cur, radices;
    detail::best_split(smooth, cur, radices);
    if (smooth == 1) radices.clear();
    // order: odd register radices first (their stride-R first-stage stores are bank-conflict free),
    // then the generic primes, then the even radices largest last.
    std::vector<int> order;
    for (int r : radices) if (r & 1) order.push_back(r);
    for (int g : generic) order.push_back(g);
    for (auto it = radices.rbegin(); it != radices.rend(); ++it) if (!(*it & 1)) order.push_back(*it);
    if ((int)order.size() > kMaxStages) { p.error = "too many FFT stages"; return p; }

    int tp = 32;
    int ns = 1;
    for (int r : order) {
        DctStage s;
        s.radix = r;
        s.ns = ns;
        s.generic = is_register_radix(r) ? 0 : 1;
        s.tw_offset = (int)(p.stage_tw.size() / 2);
        if (!s.generic) {
            for (int q = 1; q < r; ++q)
                for (int k = 0; k < ns; ++k) {
                    double ang = -2.0 * M_PI * (double)k * (double)q / ((double)ns * (double)r);
                    p.stage_tw.push_back((float)std::cos(ang));
                    p.stage_tw.push_back((float)std::sin(ang));
                }
            int nb = n / r;
            int need = (nb + stage_max_iter(r) - 1) / stage_max_iter(r);
            if (need > tp) tp = need;
        } else {
            int need = (n + kGenericOutMax - 1) / kGenericOutMax;
            if (need > tp) tp = need;
        }
        p.stages[p.nstages++] = s;
        ns *= r;
    }
    tp = (tp + 31) / 32 * 32;
    if (tp > kMaxThreadsPerLinePair) {
        p.error = "line length " + std::to_string(n) + " needs more than 1024 threads per line pair";
        return p;
    }
    p.tp = tp;
    p.npad = pad_index(n - 1) + 1;
    if (!(p.npad & 1)) p.npad += 1;  // odd stride between line pairs spreads banks in the column pass
    p.wn.resize(2 * (size_t)n);
    p.t4.resize(2 * (size_t)n);
    for (int j = 0; j < n; ++j) {
        double a = -2.0 * M_PI * (double)j / (double)n;
        p.wn[2 * j] = (float)std::cos(a);
        p.wn[2 * j + 1] = (float)std::sin(a);
        double b = -M_PI * (double)j / (2.0 * (double)n);
        p.t4[2 * j] = (float)std::cos(b);
        p.t4[2 * j + 1] = (float)std::sin(b);
    }
    return p;
}

}  // namespace ssw
