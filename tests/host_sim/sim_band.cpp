// tests/host_sim/sim_band.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the per-lane band arithmetic of isocon_b200/csrc/{myers_band,band_group}.cuh for the
// host (one "lane" at a time) so the window / strip / early-exit logic can be checked against
// the oracle without a GPU.  The product never executes this: the shipped library has no CPU path.
#include <stdint.h>
#include <vector>
#include <algorithm>
#include "../../isocon_b200/csrc/band_group.cuh"
#include "../../isocon_b200/csrc/diag_band.cuh"

using namespace isocon;

template <int W>
static int run(const uint32_t* peq, int m, const uint32_t* tgt, int n, int k, int dhi) {
    int cols; return ed_group<W>(peq, m, tgt, 1, n, k, true, dhi, &cols);
}

template <int W>
static int dispatch(int w, const uint32_t* peq, int m, const uint32_t* tgt, int n, int k, int dhi) {
    if constexpr (W > 48) { return -99; }
    else {
        if (w <= W) return run<W>(peq, m, tgt, n, k, dhi);
        return dispatch<W + 1>(w, peq, m, tgt, n, k, dhi);
    }
}

extern "C" int sim_ed(const uint8_t* q, int m, const uint8_t* t, int n, int k, int widen_lo, int widen_hi,
                      int force_w) {
    const int delta = n - m;
    if (std::abs(delta) > k) return -1;
    int lo, hi;
    lane_strip(delta, k, lo, hi);
    lo -= widen_lo; hi += widen_hi;
    int w = band_words(lo, hi);
    if (force_w > w) w = force_w;
    const int nb = (m + 31) / 32;
    std::vector<uint32_t> peq((size_t)(nb + w + 2) * 4, 0u);
    auto code = [](uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; };
    for (int i = 0; i < m; ++i) peq[(size_t)(i >> 5) * 4 + code(q[i])] |= 1u << (i & 31);
    std::vector<uint32_t> tg((n + 15) / 16 + 4, 0u);
    for (int i = 0; i < n; ++i) tg[i >> 4] |= (uint32_t)code(t[i]) << (2 * (i & 15));
    return dispatch<1>(w, peq.data(), m, tg.data(), n, k, hi);
}

// ---- diagonal band (diag_band.cuh): same harness, the shifted match-mask table built here

template <int W>
static int dispatch_diag(int w, const uint32_t* tab, int padbits, int m, const uint32_t* tgt, int n, int k, int dhi) {
    if constexpr (W > 48) { return -99; }
    else {
        if (w <= W) { int cols, done; return ed_group_diag<W>(tab, padbits, m, tgt, 1, n, k, true, dhi, &cols, &done); }
        return dispatch_diag<W + 1>(w, tab, padbits, m, tgt, n, k, dhi);
    }
}

extern "C" int sim_ed_diag(const uint8_t* q, int m, const uint8_t* t, int n, int k, int widen_lo, int widen_hi,
                           int force_w, int extra_pad_words) {
    const int delta = n - m;
    if (std::abs(delta) > k) return -1;
    int lo, hi;
    lane_strip(delta, k, lo, hi);
    lo -= widen_lo; hi += widen_hi;
    int w = diag_words(lo, hi);
    if (force_w > w) w = force_w;
    const int padbits = 32 * ((hi + 31) / 32 + extra_pad_words);
    auto code = [](uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; };
    const int X = ((padbits + m) >> 5) + w + 2;          // window words the table must serve
    std::vector<uint32_t> base((size_t)(X + 1) * 4, 0u);   // unshifted masks, one guard word
    for (int i = 0; i < m; ++i) base[(size_t)((padbits + i) >> 5) * 4 + code(q[i])] |= 1u << ((padbits + i) & 31);
    std::vector<uint32_t> tab((size_t)X * 32 * 4, 0u);
    for (int x = 0; x < X; ++x)
        for (int s = 0; s < 32; ++s)
            for (int c = 0; c < 4; ++c)
                tab[((size_t)x * 32 + s) * 4 + c] = funnel_r(base[(size_t)x * 4 + c], base[(size_t)(x + 1) * 4 + c], s);
    std::vector<uint32_t> tg((n + 15) / 16 + 4, 0u);
    for (int i = 0; i < n; ++i) tg[i >> 4] |= (uint32_t)code(t[i]) << (2 * (i & 15));
    return dispatch_diag<1>(w, tab.data(), padbits, m, tg.data(), n, k, hi);
}
