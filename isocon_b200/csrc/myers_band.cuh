// myers_band.cuh -- per-lane arithmetic of the banded Myers/Hyyro bit-vector edit distance.
//
// This is the arithmetic that replaces edlib.align(x, y, mode="NW", task="distance", k=K)
// (/root/reference/modules/nearest_neighbor_graph.py:104-107) on the device.  One CUDA
// thread (lane) owns one (query, target) pair and keeps a sliding window of W 32-bit words
// of the vertical-delta vectors Pv/Mv in REGISTERS; the words of one pair never leave the
// thread, so the multi-word carry of the Hyyro recurrence is the hardware carry flag
// (add.cc / addc.cc) and the inter-word bit of Ph/Mh is a funnel shift -- no shuffles on
// the narrow-band path.  The 32 lanes of a warp share the query (its match masks Peq sit in
// shared memory, laid out [word][symbol]) and walk their 32 targets in lock-step, so every
// Peq fetch is a 4-address broadcast LDS.
//
// Band geometry (Ukkonen strip at 32-row granularity):
//   allowed diagonals d = j - i in [dlo, dhi] (warp-uniform superset of every lane's strip);
//   window of column j = words first(j) .. first(j)+W-1, first(j) = max(0, (j-dhi-1) >> 5);
//   W >= ((dhi-dlo+1 + 30) >> 5) + 1 covers rows j-dhi .. j-dlo in every column.
//   Top word always sees hin = +1 (exact for word 0, an upper bound afterwards); a word that
//   enters at the bottom starts as Pv = ~0, Mv = 0 (upper bound D[i][j-1] = D[bot][j-1] + ...).
//   Cells whose true value is <= k keep their exact value; everything else is >= the truth.
//
// The code in this header is scalar per lane and compiles for the host as well, so the band
// logic can be unit-tested without a GPU (tests/host_sim; test infrastructure only -- the
// product never runs it on the CPU).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ISO_HD __host__ __device__ __forceinline__
#else
#define ISO_HD inline
#endif

namespace isocon {

ISO_HD int iso_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

ISO_HD uint32_t iso_funnel_l1(uint32_t below, uint32_t x) {  // (x << 1) | (below >> 31)
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(below, x, 1);
#else
    return (x << 1) | (below >> 31);
#endif
}

// 0xffffffff << clamp(sh, 0, 32): the bits of a 32-bit word at or above position sh (sh may be negative or > 32).
ISO_HD uint32_t iso_mask_from(int sh) {
    sh = sh < 0 ? 0 : sh;
#if defined(__CUDA_ARCH__)
    return __funnelshift_lc(0u, 0xffffffffu, sh);   // shift clamped at 32 by the hardware
#else
    return sh >= 32 ? 0u : (0xffffffffu << sh);
#endif
}

static constexpr int ED_PENDING = -2;

template <int W>
struct Band {
    uint32_t Pv[W], Mv[W];
    uint32_t accP, accM;  // bit 31 of the bottom word's Ph / Mh for the columns since the last flush
    int score;            // D[bottom row of the window] at the column of the last flush

    ISO_HD void init() {
#pragma unroll
        for (int w = 0; w < W; ++w) { Pv[w] = 0xffffffffu; Mv[w] = 0u; }
        accP = accM = 0u;
        score = 32 * W;
    }

    // One column.  eq points at Peq[first][c]; consecutive window words are 4 entries apart.
    ISO_HD void column(const uint32_t* __restrict__ eq) {
        uint32_t ph_below = 0x80000000u;  // hin = +1 for the top word
        uint32_t mh_below = 0u;
#if !defined(__CUDA_ARCH__)
        uint32_t carry = 0u;
#endif
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const uint32_t Eq = eq[4 * w];
            const uint32_t pv = Pv[w], mv = Mv[w];
            const uint32_t t = Eq & pv;
            uint32_t s;
#if defined(__CUDA_ARCH__)
            if (w == 0) asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(s) : "r"(t), "r"(pv));
            else        asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(s) : "r"(t), "r"(pv));
#else
            const uint64_t wide = (uint64_t)t + (uint64_t)pv + (uint64_t)carry;
            s = (uint32_t)wide; carry = (uint32_t)(wide >> 32);
#endif
            const uint32_t Xh = (s ^ pv) | Eq;
            const uint32_t Ph = mv | ~(Xh | pv);
            const uint32_t Mh = pv & Xh;
            const uint32_t Xv = Eq | mv;
            const uint32_t Phs = iso_funnel_l1(ph_below, Ph);
            const uint32_t Mhs = iso_funnel_l1(mh_below, Mh);
            Pv[w] = Mhs | ~(Xv | Phs);
            Mv[w] = Phs & Xv;
            ph_below = Ph; mh_below = Mh;
        }
        accP = iso_funnel_l1(ph_below, accP);
        accM = iso_funnel_l1(mh_below, accM);
    }

    ISO_HD void flush() {  // at most 32 columns may be pending
        score += iso_popc(accP) - iso_popc(accM);
        accP = accM = 0u;
    }

    // The window slides down by one word (call BEFORE the first column of the new position).
    ISO_HD void shift() {
#pragma unroll
        for (int w = 0; w + 1 < W; ++w) { Pv[w] = Pv[w + 1]; Mv[w] = Mv[w + 1]; }
        Pv[W - 1] = 0xffffffffu; Mv[W - 1] = 0u;
        score += 32;  // pending accP/accM stay valid: they are deltas of the old bottom row
    }

    // D at window-relative row o (rows 1..o of the window lie at or above it); 0 <= o <= 32W.
    // Requires flush().
    ISO_HD int value_at(int o) const {
        int d = score;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int rel = o - 32 * w;  // rows of this word at or above the cell
            const uint32_t above = rel <= 0 ? 0xffffffffu : (rel >= 32 ? 0u : (0xffffffffu << rel));
            d -= iso_popc(Pv[w] & above);
            d += iso_popc(Mv[w] & above);
        }
        return d;
    }
};

// Number of window words needed for a strip of diagonals [dlo, dhi].
ISO_HD int band_words(int dlo, int dhi) { return ((dhi - dlo + 1 + 30) >> 5) + 1; }

// One lane's strip for threshold k and delta = n - m (requires |delta| <= k).
ISO_HD void lane_strip(int delta, int k, int& lo, int& hi) {
    const int ad = delta < 0 ? -delta : delta;
    const int p = (k - ad) >> 1;
    lo = (delta < 0 ? delta : 0) - p;
    hi = (delta > 0 ? delta : 0) + p;
}

}  // namespace isocon
