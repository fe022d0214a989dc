// ref_draw_order.h — TEST INFRASTRUCTURE (oracle).  See gen_pair_lines.py.
// Included AFTER the reference's lcg_random.cuh and BEFORE the reference files that
// draw random numbers; then `#define lcg_randomf(r) refemu::ordered_draw(...)`.
#pragma once
namespace refemu {
struct PairLine { const char* file; int line; };
static const PairLine kPairLines[] = {
#include "_ref/ref_pair_lines.inc"
};
inline bool is_pair_line(const char* path, int line) {
    const char* base = strrchr(path, '/');
    base = base ? base + 1 : path;
    for (const PairLine& p : kPairLines)
        if (p.line == line && strcmp(p.file, base) == 0) return true;
    return false;
}
// g++ calls the textual-right draw first.  On a paired line: pull both draws, hand
// the SECOND to this (right) call and keep the FIRST for the upcoming (left) call.
inline float ordered_draw(LCGRand& rng, const char* file, int line) {
    static thread_local bool holding = false;
    static thread_local float held = 0.f;
    static thread_local int held_line = -1;
    if (holding && held_line == line) { holding = false; return held; }
    if (is_pair_line(file, line)) {
        float first = lcg_randomf(rng);
        float second = lcg_randomf(rng);
        held = first; held_line = line; holding = true;
        return second;
    }
    return lcg_randomf(rng);
}
}  // namespace refemu
