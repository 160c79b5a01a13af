// band_group.cuh -- one warp walks 32 (query, target) pairs that share the query, in lock-step.
//
// Lane l owns target l of an interleaved group (word w of lane l sits at tgt[w * 32]),
// so every 32-bit target load of the warp is one coalesced 128-byte transaction.  All lanes
// use the same strip of diagonals [dlo, dhi] (the union of the lanes' Ukkonen strips), which
// keeps the window position, the Peq row and the control flow warp-uniform; each lane keeps
// its OWN threshold k for the early exit and for the final "<= k else -1" test, so results
// are exactly edlib's (distance if <= k, else -1).
//
// Early exit (exact; SURVEY.md Appendix C.2 rule 9 left it to the builder): values never
// decrease along a diagonal and  min_i D[i][j] + |delta - (j - i)|  is attained exactly on the
// final diagonal, so a lane is finished as soon as the cell on its final diagonal
// (row j - delta) exceeds k.  Checked every 32 columns.
//
// Columns are processed as: a short generic head that aligns j to the window-shift phase,
// fully unrolled 32-column chunks (two 16-symbol registers cut out of the target stream
// with funnel shifts), and a generic tail that also catches the lanes whose target ends.
#pragma once
#include "myers_band.cuh"

namespace isocon {

#define ISO_FULL 0xffffffffu
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ bool warp_all(bool p) { return __all_sync(ISO_FULL, p); }
__device__ __forceinline__ int warp_max(int v) { return __reduce_max_sync(ISO_FULL, v); }
__device__ __forceinline__ int warp_min(int v) { return __reduce_min_sync(ISO_FULL, v); }
__device__ __forceinline__ uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh) { return __funnelshift_r(lo, hi, sh); }
#else
#if defined(ISO_SIM_WARP)
// tests/host_sim runs 32 host threads in lock step: the harness provides the reduction (op 0 = min, 1 = max)
int sim_warp_reduce(int v, int op);
inline bool warp_all(bool p) { return sim_warp_reduce(p ? 1 : 0, 0) != 0; }
inline int warp_max(int v) { return sim_warp_reduce(v, 1); }
inline int warp_min(int v) { return sim_warp_reduce(v, 0); }
#else
inline bool warp_all(bool p) { return p; }
inline int warp_max(int v) { return v; }
inline int warp_min(int v) { return v; }
#endif
inline uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh) {
    sh &= 31;
    return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
}
#endif

// peq      : match masks of the query, [word][4]; zero for every word at or beyond
//            ceil(m/32) (at least W + 1 such words must be readable)
// m        : query length (warp-uniform)
// tgt, ts  : this lane's 2-bit target stream, word w at tgt[w * ts], 16 bases per word, at
//            least 3 readable words beyond the last one (ts = 32 for an interleaved group,
//            1 for a row-major read)
// n, k     : this lane's target length and threshold;  active: lane has a pair
// dhi      : warp-uniform top diagonal of the strip [dlo, dhi], a superset of every active
//            lane's own strip; W >= band_words(dlo, dhi)
// cols     : out, number of columns the warp walked (work counter)
// returns  : edit distance if <= k, else -1   (inactive lanes: -1)
template <int W>
ISO_HD int ed_group(const uint32_t* __restrict__ peq, int m, const uint32_t* __restrict__ tgt, int ts,
                    int n, int k, bool active, int dhi, int* cols) {
    const int delta = n - m;
    Band<W> B;
    B.init();
    int res = active ? ED_PENDING : -1;
    if (active && n == 0) res = (m <= k) ? m : -1;
    if (active && m == 0) res = (n <= k) ? n : -1;
    const int nmax = warp_max(res == ED_PENDING ? n : 0);
    const int nmin = warp_min(res == ED_PENDING ? n : 0x7fffffff);
    *cols = 0;
    if (nmax == 0) return res;

    int first = 0;   // top word of the window
    int j = 1;       // next column (1-based)
    int tw_idx = -1; // target word held in tw (generic steps)
    uint32_t tw = 0;

#define ISO_DIAG_CHECK(jc)                                                       \
    do {                                                                         \
        const int r_ = (jc) - delta; /* row of the final diagonal in column jc */ \
        if (res == ED_PENDING && r_ >= 1 && (jc) < n) {                          \
            if (B.value_at(r_ - 32 * first) > k) res = -1;                       \
        }                                                                        \
    } while (0)

#define ISO_GENERIC_STEP(jc)                                                     \
    do {                                                                         \
        const int a_ = (jc) - dhi - 1;                                           \
        if (a_ >= 32 && (a_ & 31) == 0) { B.shift(); ++first; }                  \
        const int wi_ = ((jc) - 1) >> 4;                                         \
        if (wi_ != tw_idx) { tw = tgt[wi_ * ts]; tw_idx = wi_; }            \
        const uint32_t c_ = (tw >> (2 * (((jc) - 1) & 15))) & 3u;                \
        B.column(peq + 4 * first + c_);                                          \
        if ((jc) == n && res == ED_PENDING) {                                    \
            B.flush();                                                           \
            const int d_ = B.value_at(m - 32 * first);                           \
            res = d_ <= k ? d_ : -1;                                             \
        }                                                                        \
    } while (0)

    // head: columns 1 .. (dhi mod 32); afterwards j - dhi - 1 is a multiple of 32
    {
        const int head = dhi & 31;
        const int head_end = head < nmax ? head : nmax;
        for (; j <= head_end; ++j) ISO_GENERIC_STEP(j);
        B.flush();
    }

    // body: unrolled chunks of 32 columns while every pending lane still has 32 columns left
    while (j + 31 <= nmin) {
        if (j - dhi - 1 >= 32) { B.shift(); ++first; }
        const int b0 = j - 1, wi = b0 >> 4, sh = 2 * (b0 & 15);
        const uint32_t w0 = tgt[wi * ts], w1 = tgt[(wi + 1) * ts], w2 = tgt[(wi + 2) * ts];
        const uint32_t lo = funnel_r(w0, w1, sh), hi = funnel_r(w1, w2, sh);
        const uint32_t* prow = peq + 4 * first;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int h = 0; h < 2; ++h) {
            const uint32_t cur = h ? hi : lo;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < 16; ++i) B.column(prow + ((cur >> (2 * i)) & 3u));
        }
        B.flush();
        j += 32;
        if (j - 1 == n && res == ED_PENDING) {  // the chunk ended exactly on this lane's last column
            const int d = B.value_at(m - 32 * first);
            res = d <= k ? d : -1;
        }
        ISO_DIAG_CHECK(j - 1);
        if (warp_all(res != ED_PENDING)) { *cols = j; return res; }
    }

    // tail: one column at a time; lanes finish when their target ends
    for (; j <= nmax; ++j) {
        ISO_GENERIC_STEP(j);
        if ((j & 31) == 0) {
            B.flush();
            ISO_DIAG_CHECK(j);
            if (warp_all(res != ED_PENDING)) { *cols = j; return res; }
        }
    }
#undef ISO_DIAG_CHECK
#undef ISO_GENERIC_STEP
    *cols = nmax;
    return res;
}

}  // namespace isocon
