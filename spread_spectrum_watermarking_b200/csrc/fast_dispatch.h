// Length -> compile-time plan dispatch shared by libssw (ssw_api.cu) and tests/emul.
#pragma once
#include "dct_fast.cuh"

namespace ssw {
namespace fast {

// default team counts per CTA: rows G pairs (one team each), columns G pairs = 2G adjacent columns
template <class P> struct RowG { static constexpr int value = (P::T >= 192) ? 1 : 2; };
constexpr int kColG = 4;      // column pairs per CTA (8 adjacent columns = one 32-byte sector per row)
constexpr int kColTeams = 2;  // teams transforming them (G / TEAMS rounds)

template <class F>
inline bool with_plan(int n, F&& f) {
    switch (n) {
        case 3840: f(Plan3840{}); return true;
        case 2160: f(Plan2160{}); return true;
        case 1920: f(Plan1920{}); return true;
        case 1080: f(Plan1080{}); return true;
        case 640: f(Plan640{}); return true;
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
