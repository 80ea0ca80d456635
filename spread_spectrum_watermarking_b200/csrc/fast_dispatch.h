// Length -> compile-time plan dispatch shared by libssw (ssw_api.cu) and tests/emul.
#pragma once
#include "dct_fast.cuh"

namespace ssw {
namespace fast {

// default team counts per CTA: rows G pairs (one team each), columns G pairs = 2G adjacent columns
template <class P> struct RowG { static constexpr int value = (P::T >= 192) ? 1 : 2; };
constexpr int kColG = 4;      // column pairs per CTA (8 adjacent columns = one 32-byte sector per row)
constexpr int kColTeams = 2;  // teams transforming them (G / TEAMS rounds)
constexpr int kMaxSmem = 227 * 1024;
// column tiles of the long power-of-two lines: fewer pairs so that the tile still fits shared memory
template <class P> struct ColG {
    static constexpr int value = (4 * P::PITCH * 8 <= kMaxSmem) ? 4 : ((2 * P::PITCH * 8 <= kMaxSmem) ? 2 : 0);
};
template <class P> struct ColTeams { static constexpr int value = (P::T >= 512) ? 1 : 2; };

template <class F>
inline bool with_plan(int n, F&& f) {
    switch (n) {
        case 3840: f(Plan3840{}); return true;
        case 2160: f(Plan2160{}); return true;
        case 1920: f(Plan1920{}); return true;
        case 1080: f(Plan1080{}); return true;
        case 640: f(Plan640{}); return true;
        case 1280: f(Plan1280{}); return true;
        case 720: f(Plan720{}); return true;
        case 2560: f(Plan2560{}); return true;
        case 1440: f(Plan1440{}); return true;
        case 7680: f(Plan7680{}); return true;
        case 4320: f(Plan4320{}); return true;
        case 1024: f(Plan1024{}); return true;
        case 2048: f(Plan2048{}); return true;
        case 4096: f(Plan4096{}); return true;
        case 8192: f(Plan8192{}); return true;
        case 16384: f(Plan16384{}); return true;
        default: return false;
    }
}

inline bool has_plan(int n) { return with_plan(n, [](auto) {}); }

// single-line kernels (one real line through an n/2-point complex FFT): line length -> plan of the FFT
template <class F>
inline bool with_line1_plan(int n, F&& f) {
    switch (n) {
        case 1024: f(PlanL512{}); return true;
        case 4096: f(PlanL2048{}); return true;
        case 32768: f(PlanL16384{}); return true;
        default: return false;
    }
}
inline bool has_line1_plan(int n) { return with_line1_plan(n, [](auto) {}); }

}  // namespace fast
}  // namespace ssw
