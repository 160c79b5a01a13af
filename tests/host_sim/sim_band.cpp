// tests/host_sim/sim_band.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the per-lane band arithmetic of isocon_b200/csrc/{myers_band,band_group}.cuh for the
// host (one "lane" at a time) so the window / strip / early-exit logic can be checked against
// the oracle without a GPU.  The product never executes this: the shipped library has no CPU path.
#include <stdint.h>
#include <vector>
#include <algorithm>
#include <pthread.h>
#include <thread>
#define ISO_SIM_WARP 1
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

// ---- the kernels' group walk with the shrinking window (ed_group_diag_run): one lane; uwlo / uwhi widen the
// alive interval like the other lanes of a warp would.  out[0] = sum of words x columns, out[1] = columns walked.
extern "C" int sim_ed_diag_run(const uint8_t* q, int m, const uint8_t* t, int n, int k, int widen_lo, int widen_hi,
                               int force_w, int extra_pad_words, int narrow, int uwlo, int uwhi, int* out) {
    const int delta = n - m;
    if (std::abs(delta) > k) return -1;
    int lo, hi;
    lane_strip(delta, k, lo, hi);
    lo -= widen_lo; hi += widen_hi;
    int w = diag_words(lo, hi);
    if (force_w > w) w = force_w;
    if (w > DIAG_WMAX) return -99;
    w = diag_avail(w);
    const int padbits = 32 * ((hi + 31) / 32 + extra_pad_words);
    auto code = [](uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; };
    const int X = ((padbits + m) >> 5) + w + 2;
    std::vector<uint32_t> base((size_t)(X + 1) * 4, 0u);
    for (int i = 0; i < m; ++i) base[(size_t)((padbits + i) >> 5) * 4 + code(q[i])] |= 1u << ((padbits + i) & 31);
    std::vector<uint32_t> tab((size_t)X * 32 * 4, 0u);
    for (int x = 0; x < X; ++x)
        for (int s = 0; s < 32; ++s)
            for (int c = 0; c < 4; ++c)
                tab[((size_t)x * 32 + s) * 4 + c] = funnel_r(base[(size_t)x * 4 + c], base[(size_t)(x + 1) * 4 + c], s);
    std::vector<uint32_t> tg((n + 15) / 16 + 4, 0u);
    for (int i = 0; i < n; ++i) tg[i >> 4] |= (uint32_t)code(t[i]) << (2 * (i & 15));
    g_sim_widen_lo = uwlo; g_sim_widen_hi = uwhi;
    int cols = 0; unsigned wcols = 0, ucells = 0;
    const int rows = std::min(m, std::abs(delta) + 2 * ((k - std::abs(delta)) >> 1) + 1);
    const int r = ed_group_diag_run(w, tab.data(), padbits, m, tg.data(), 1, n, k, true, hi, narrow, rows, &cols, &wcols, &ucells);
    g_sim_widen_lo = g_sim_widen_hi = 0;
    if (out) { out[0] = (int)wcols; out[1] = cols; out[2] = (int)ucells; out[3] = w; }
    return r;
}

// ---- a whole warp on the host: 32 threads in lock step, the warp reductions through a barrier.  A lane that
// takes another path through a warp collective than its neighbours dead-locks here (the test has a timeout)
// exactly where the GPU would misbehave.
namespace {
thread_local int tl_lane = -1;           // -1: single-lane call, reductions are the identity
int g_slot[32];
pthread_barrier_t g_bar;
}
namespace isocon {
int sim_warp_reduce(int v, int op) {
    if (tl_lane < 0) return v;
    g_slot[tl_lane] = v;
    pthread_barrier_wait(&g_bar);
    int r = g_slot[0];
    for (int l = 1; l < 32; ++l) r = op ? std::max(r, g_slot[l]) : std::min(r, g_slot[l]);
    pthread_barrier_wait(&g_bar);
    return r;
}
}

// q: the query; targets concatenated in tcat with toff[33]; k[l] < 0 = lane without a pair.  The window placement is
// the row kernel's (kmax, per-lane dhi).  out_r[32] results, out[0] = words x columns, out[1] = columns, out[2] = W0.
extern "C" int sim_warp_diag_run(const uint8_t* q, int m, const uint8_t* tcat, const int* toff, const int* k,
                                 int narrow, int* out_r, int* out) {
    auto code = [](uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; };
    int kmax = -1, nmaxlen = 0;
    bool need[32]; int n[32], dl[32], dhi_l[32];
    for (int l = 0; l < 32; ++l) {
        n[l] = toff[l + 1] - toff[l]; dl[l] = std::abs(n[l] - m);
        need[l] = k[l] >= 0 && dl[l] <= k[l];
        if (need[l]) kmax = std::max(kmax, k[l]);
        nmaxlen = std::max(nmaxlen, n[l]);
    }
    if (kmax < 0) return -1;
    const int Wd = (kmax + 32) >> 5;
    if (Wd > DIAG_WMAX) return -99;
    int dhi_max = 0;
    for (int l = 0; l < 32; ++l) { dhi_l[l] = need[l] ? std::max(0, n[l] - m) + ((kmax - dl[l]) >> 1) : 0; dhi_max = std::max(dhi_max, dhi_l[l]); }
    for (int l = 0; l < 32; ++l) if (!need[l]) dhi_l[l] = dhi_max;
    const int padbits = 32 * ((dhi_max + 31) / 32);
    const int X = ((padbits + m) >> 5) + DIAG_WMAX + 3;
    std::vector<uint32_t> base((size_t)(X + 1) * 4, 0u);
    for (int i = 0; i < m; ++i) base[(size_t)((padbits + i) >> 5) * 4 + code(q[i])] |= 1u << ((padbits + i) & 31);
    std::vector<uint32_t> tab((size_t)X * 32 * 4, 0u);
    for (int x = 0; x < X; ++x)
        for (int s = 0; s < 32; ++s)
            for (int c = 0; c < 4; ++c)
                tab[((size_t)x * 32 + s) * 4 + c] = funnel_r(base[(size_t)x * 4 + c], base[(size_t)(x + 1) * 4 + c], s);
    const int words = (nmaxlen + 15) / 16 + 4;               // interleaved like the device layout: word w of lane l at il[32 w + l]
    std::vector<uint32_t> il((size_t)words * 32, 0u);
    for (int l = 0; l < 32; ++l)
        for (int i = 0; i < n[l]; ++i) il[(size_t)(i >> 4) * 32 + l] |= (uint32_t)code(tcat[toff[l] + i]) << (2 * (i & 15));
    pthread_barrier_init(&g_bar, nullptr, 32);
    int colsv[32]; unsigned wc[32], uc[32];
    std::vector<std::thread> th;
    for (int l = 0; l < 32; ++l)
        th.emplace_back([&, l]() {
            tl_lane = l;
            const int rows = need[l] ? std::min(m, dl[l] + 2 * ((k[l] - dl[l]) >> 1) + 1) : 0;
            out_r[l] = ed_group_diag_run(Wd, tab.data(), padbits, m, il.data() + l, 32, n[l], k[l], need[l], dhi_l[l], narrow,
                                         rows, &colsv[l], &wc[l], &uc[l]);
            tl_lane = -1;
        });
    for (auto& t : th) t.join();
    pthread_barrier_destroy(&g_bar);
    for (int l = 1; l < 32; ++l) if (wc[l] != wc[0] || colsv[l] != colsv[0]) return -98;   // must be warp-uniform
    out[0] = (int)wc[0]; out[1] = colsv[0]; out[2] = diag_avail(Wd);
    return 0;
}
