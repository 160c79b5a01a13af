// diag_band.cuh -- diagonal-band Myers/Hyyro bit-vector edit distance, lane-per-pair.
//
// Same contract as band_group.cuh (edlib.align(x, y, mode="NW", task="distance", k=K) of
// /root/reference/modules/nearest_neighbor_graph.py:104-107: the distance if <= k, else -1),
// different band geometry.  band_group.cuh keeps a window of W words aligned to 32-row
// blocks and slides it every 32 columns, so a strip of `width` diagonals costs
// ceil((width + 31) / 32) words per column.  Here the window slides ONE row per column and
// therefore always covers exactly the diagonals [dhi - 32W + 1, dhi]: ceil(width / 32) words
// per column -- one word less on average (c2: 5 instead of 6; late correction rounds with
// k < 32: 1 instead of 2).
//
// Geometry.  Column j (1-based) holds rows top_j .. top_j + 32W - 1, top_j = j - dhi; bit b of
// the window is row top_j + b.  Rows <= 0 are VIRTUAL: the matrix is extended upwards with
// D[i][j] = j - i (i <= 0), which satisfies the edit-distance recurrence when those rows match
// nothing, reproduces the NW boundary D[0][j] = j exactly and needs no special case while the
// window still hangs over the top of the matrix.  Rows > m match nothing either and never feed
// a real row.
//
// Recurrence per column (Hyyro 2003, diagonal tiling), vertical deltas stored PRE-SHIFTED: after
// column j, bit b of VP/VN is the vertical delta D[r][j] - D[r-1][j] of row r = top_j + b + 1,
// i.e. it is already aligned with the window of column j + 1:
//     D0 = (((Eq & VP) + VP) ^ VP) | Eq | VN      diagonal-zero vector, carry chain across words
//     HP = VN | ~(D0 | VP)        HN = D0 & VP    horizontal deltas of column j
//     X  = D0 >> 1                                 (across words; the new bottom bit gets 0)
//     VN = X & HP                 VP = HN | ~(X | HP)
// 7 LOP3 + 1 IADD3.X + 1 SHF per word.  The top cell of the window gets no carry-in (the cell
// above the band is an upper bound), the cell entering at the bottom is D[r-1][j-1] + 1 (an
// upper bound), so every cell is >= the truth and exact whenever the truth is <= k (Ukkonen).
//
// The match vector of column j must be aligned to the sliding window: bits [o, o + 32W) of the
// query's match mask, o = padbits + j - dhi - 1.  A per-column funnel shift would cost one more
// SHF per word, so the block stages ALL 32 bit-shifts of the query's masks in shared memory
// once per query:   tab[4 * o + c] = bits [o, o + 32) of mask c   for EVERY bit offset o,
// where mask bit padbits + i is "query[i] == c".  A column's fetch is then W LDS at
// tab + 4 * o + c + 128 * w with compile-time offsets in the unrolled 32-column body; the lanes
// of a warp differ in c and (by a few diagonals) in o, i.e. they read a handful of neighbouring
// 16-byte entries: one shared-memory wavefront.
//
// Score: the bottom diagonal's value follows D[r][j] - D[r-1][j-1] = 1 - D0[bottom bit]; the
// bottom bits of 32 columns are collected with one funnel shift per column and counted with
// one POPC per 32 columns.  Any other cell of the column follows from the vertical deltas.
//
// Like band_group.cuh this header is scalar per lane and compiles for the host (tests/host_sim).
#pragma once
#include "band_group.cuh"

namespace isocon {

template <int W>
struct DiagBand {
    uint32_t VP[W], VN[W];
    uint32_t acc;   // D0 bottom bits of the columns since the last flush (acc starts at 0)
    int score;      // D on the bottom diagonal at the column of the last flush

    ISO_HD void init(int dhi) {
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int lo = dhi - 32 * w;   // bits below lo are rows <= 0 at column 0: delta -1
            const uint32_t vn = lo <= 0 ? 0u : (lo >= 32 ? 0xffffffffu : ((1u << lo) - 1u));
            VN[w] = vn; VP[w] = ~vn;
        }
        acc = 0u;
        score = 32 * W - 1 - dhi;          // D[-(dhi - 32W + 1)][0]; the bottom diagonal is <= 0
    }

    // eq points at tab[(x0 * 32 + s) * 4 + c]; consecutive window words are 128 entries apart.
    ISO_HD void column(const uint32_t* __restrict__ eq) {
        uint32_t D0[W];
#if !defined(__CUDA_ARCH__)
        uint32_t carry = 0u;
#endif
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const uint32_t Eq = eq[128 * w];
            const uint32_t vp = VP[w];
            const uint32_t t = Eq & vp;
            uint32_t s;
#if defined(__CUDA_ARCH__)
            if (W == 1)          s = t + vp;
            else if (w == 0)     asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(s) : "r"(t), "r"(vp));
            else if (w == W - 1) asm volatile("addc.u32 %0, %1, %2;" : "=r"(s) : "r"(t), "r"(vp));
            else                 asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(s) : "r"(t), "r"(vp));
#else
            const uint64_t wide = (uint64_t)t + (uint64_t)vp + (uint64_t)carry;
            s = (uint32_t)wide; carry = (uint32_t)(wide >> 32);
#endif
            D0[w] = (s ^ vp) | Eq | VN[w];
        }
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const uint32_t d0 = D0[w], vp = VP[w], vn = VN[w];
            const uint32_t HP = vn | ~(d0 | vp);
            const uint32_t HN = d0 & vp;
            const uint32_t X = (w + 1 < W) ? funnel_r(d0, D0[w + 1 < W ? w + 1 : w], 1) : (d0 >> 1);
            VN[w] = X & HP;
            VP[w] = HN | ~(X | HP);
        }
        acc = iso_funnel_l1(D0[W - 1], acc);
    }

    ISO_HD void flush(int ncols) {  // ncols <= 32 columns since the last flush
        score += ncols - iso_popc(acc);
        acc = 0u;
    }

    // D[top_j + pos][j] of the last processed column j (0 <= pos <= 32W - 1); `pend` columns
    // (<= 32) are still collected in acc.
    ISO_HD int value_at(int pos, int pend) const {
        int d = score + pend - iso_popc(acc);
#pragma unroll
        for (int w = 0; w < W; ++w) {
            uint32_t mask = iso_mask_from(pos - 32 * w);
            if (w == W - 1) mask &= 0x7fffffffu;   // bit 32W-1 belongs to the row below the window
            d -= iso_popc(VP[w] & mask);
            d += iso_popc(VN[w] & mask);
        }
        return d;
    }
};

// Window words needed for a strip of diagonals [dlo, dhi].
ISO_HD int diag_words(int dlo, int dhi) { return (dhi - dlo + 32) >> 5; }

// ---------------------------------------------------------------------------------------------
// Shrinking window.  A cell of column j is ALIVE when  D + |its diagonal - final diagonal| <= k:
// only alive cells can lie on an alignment of cost <= k.  Going outwards from the final diagonal
// that sum never falls (neighbours in a column differ by at most 1), every cell on an optimal path
// to an alive cell is alive, and a dead cell only has dead successors -- so the alive cells of a
// column form ONE interval around the final diagonal, it only shrinks as the score on the final
// diagonal grows, and cells outside it may be dropped for good: computed values stay exact on every
// alive cell (they are >= the truth elsewhere), hence the result is still edlib's.
// The static Ukkonen strip is the alive interval of column 0; at 10 % divergence the interval has
// shrunk to half of it halfway to the early exit.  Every `narrow` chunks of 32 columns each lane
// bounds its interval at 16-bit granularity (4 POPC per word), the warp takes the union over its
// pending lanes, and when that fits fewer words all lanes shift their windows by the SAME number of
// bits (so their table offsets stay neighbours: one shared-memory wavefront) and the walk continues in
// the instance for the narrower width.  The state between two widths is a DiagCarry.

static constexpr int DIAG_WMAX = 14;

// Columns per unrolled block of the 32-column chunk.  The shrinking window has the warps of an SM in different
// width instances at the same time, so the code of ALL widths in use must fit the instruction caches together (B200:
// L0 ~6 KB per sub-partition, L1.5 32 KB): with 16 columns per block (11-14 KB per body) the kernel became
// fetch-bound as soon as two widths were live.  The blocks below are the best measured for c2 (<= 5 widths live,
// ALU-bound: shorter blocks cost 2 % in loop overhead); workloads with wide windows (c3 / c4: up to 14 widths
// live, fetch-bound) would gain another 6-7 % from 8, 4, 2, 1, 1, ... (profiles/r01h_ab_unroll_policies.txt).  Two
// sets of instances chosen per group by its initial width were tried and lost: both sets are then live at once.
#ifndef DIAG_UNROLL
#define DIAG_UNROLL 0   // 0 = by width (below); 16, 8, 4, 2, 1 = the same for every width (tuning builds)
#endif
template <int W> struct DiagUnroll {
    static constexpr int value = DIAG_UNROLL ? DIAG_UNROLL : (W <= 2 ? 8 : (W <= 4 ? 4 : (W <= 6 ? 2 : 1)));
};

// Widths a walk may use.  Every width up to DIAG_WMAX has an instance; DIAG_COARSE > 0 rounds the widths above it up
// to even ones, so that a walk that starts wide (c3 / c4 pilot passes: 13 words) passes through half as many
// instances on its way down and fewer bodies compete for the instruction caches.
#ifndef DIAG_COARSE
#define DIAG_COARSE 0
#endif
ISO_HD int diag_avail(int w) {
    if (DIAG_COARSE > 0 && w > DIAG_COARSE) return (w + 1) & ~1;
    return w;
}

template <int CAP>
struct DiagCarry {
    uint32_t VP[CAP], VN[CAP];
    int score;            // D at the bottom cell of the window, column j - 1
    int j;                // next column (1-based)
    int pos;              // window bit of this lane's final diagonal
    int res, done_col;
    const uint32_t* prow; // masks of the window top in column j
};

// Union hooks: on the host (one lane) the test harness widens the interval like other lanes would.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ int narrow_union_lo(int v) { return __reduce_min_sync(ISO_FULL, v); }
__device__ __forceinline__ int narrow_union_hi(int v) { return __reduce_max_sync(ISO_FULL, v); }
#else
static int g_sim_widen_lo = 0, g_sim_widen_hi = 0;
inline int narrow_union_lo(int v) { v = warp_min(v); return v - g_sim_widen_lo < 0 ? 0 : v - g_sim_widen_lo; }
inline int narrow_union_hi(int v) { return warp_max(v) + g_sim_widen_hi; }
#endif

// Conservative bounds [blo, bhi] (window bits, column of the last flush) of the alive interval: every cell
// outside is dead.  Only the cells at bits 0, 16, 32, ... (windows above 5 words: 0, 32, 64, ...) and the bottom
// cell are evaluated.
template <int W>
ISO_HD void diag_alive(const DiagBand<W>& B, int pos, int k, int& blo, int& bhi) {
    blo = 0; bhi = 32 * W - 1;
    const int kp = k + pos, km = k - pos;
    // below the final diagonal (b > pos) a cell is dead when v + (b - pos) > k, above it when v + (pos - b) > k
    if (32 * W - 1 > pos && B.score + (32 * W - 1) > kp) bhi = 32 * W - 2;      // bottom cell (bit 32W - 1): v = score
    int v = B.score;
    constexpr bool HALVES = W <= 5;     // wide windows: word boundaries only (half the POPCs, same relative precision)
#pragma unroll
    for (int w = W - 1; w >= 0; --w) {
        if (HALVES) {
            const uint32_t hm = (w == W - 1) ? 0x7fff0000u : 0xffff0000u;
            v += iso_popc(B.VN[w] & hm) - iso_popc(B.VP[w] & hm);       // v = D at bit 32w + 16
            const int b = 32 * w + 16;
            if (b > pos && v + b > kp) bhi = b - 1;                      // b falls: the smallest dead b wins
            if (b < pos && v - b > km) blo = blo > b + 1 ? blo : b + 1;  // the largest dead b wins
            v += iso_popc(B.VN[w] & 0xffffu) - iso_popc(B.VP[w] & 0xffffu);   // v = D at bit 32w
        } else {
            const uint32_t fm = (w == W - 1) ? 0x7fffffffu : 0xffffffffu;
            v += iso_popc(B.VN[w] & fm) - iso_popc(B.VP[w] & fm);       // v = D at bit 32w
        }
        const int b = 32 * w;
        if (b > pos && v + b > kp) bhi = b - 1;
        if (b < pos && v - b > km) blo = blo > b + 1 ? blo : b + 1;
    }
}

// Walks the group at window width W from column C.j on.  Returns 0 when every lane has its result, else the
// narrower width to continue with (C then holds the shifted state).
//   nmin, nmax : smallest / largest target length among the lanes pending at the start of the group
//   narrow     : try to shrink the window every `narrow` chunks of 32 columns (0 = never)
template <int W, int CAP>
ISO_HD int diag_segment(DiagCarry<CAP>& C, const uint32_t* __restrict__ tgt, int ts, int n, int k, int nmin, int nmax,
                        int narrow) {
    DiagBand<W> B;
#pragma unroll
    for (int w = 0; w < W; ++w) { B.VP[w] = C.VP[w]; B.VN[w] = C.VN[w]; }
    B.acc = 0u; B.score = C.score;
    int j = C.j, pos = C.pos, res = C.res;
    const uint32_t* prow = C.prow;
    int until = narrow;

    // body: unrolled chunks of 32 columns while every pending lane still has 32 columns left
    while (j + 31 <= nmin) {
        const int b0 = j - 1, wi = b0 >> 4, sh = 2 * (b0 & 15);
        const uint32_t w0 = tgt[wi * ts], w1 = tgt[(wi + 1) * ts], w2 = tgt[(wi + 2) * ts];
        const uint32_t lo = funnel_r(w0, w1, sh), hi = funnel_r(w1, w2, sh);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int h = 0; h < 2; ++h) {
            uint32_t cur = h ? hi : lo;
            const uint32_t* p = prow + (h << 6);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            constexpr int U = DiagUnroll<W>::value;
            for (int u = 0; u < 16 / U; ++u) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int i = 0; i < U; ++i) B.column(p + 4 * i + ((cur >> (2 * i)) & 3u));
                if (U < 16) { cur >>= (2 * U) & 31; p += 4 * U; }
            }
        }
        B.flush(32);
        j += 32; prow += 128;
        if (res == ED_PENDING) {
            const int d = B.value_at(pos, 0);    // the cell on the final diagonal: never decreases
            if (j - 1 == n) res = d <= k ? d : -1;
            else if (d > k) res = -1;
            if (res != ED_PENDING) C.done_col = j - 1;
        }
        if (warp_all(res != ED_PENDING)) { C.j = j; C.res = res; return 0; }
        if (W > 1 && narrow > 0 && --until == 0) {
            until = narrow;
            int blo = 0x3fffffff, bhi = -1;
            if (res == ED_PENDING) diag_alive<W>(B, pos, k, blo, bhi);
            const int BLO = narrow_union_lo(blo), BHI0 = narrow_union_hi(bhi);
            const int BHI = BHI0 > 32 * W - 1 ? 32 * W - 1 : BHI0;
            const int U = BHI - BLO + 1;
            const int Wn = diag_avail((U + 31) >> 5);
            if (Wn < W) {
                int s = BLO - ((32 * Wn - U) >> 1);
                s = s < 0 ? 0 : s;
                s = s > 32 * (W - Wn) ? 32 * (W - Wn) : s;
                // the new window = old bits [s, s + 32 Wn): warp-uniform shift
                C.score = B.value_at(s + 32 * Wn - 1, 0);
                const int q = s >> 5, r = s & 31;
#pragma unroll
                for (int qq = 0; qq < W; ++qq) {
                    if (q == qq) {
#pragma unroll
                        for (int w = 0; w + qq < W; ++w) {
                            const uint32_t pl = B.VP[w + qq], nl = B.VN[w + qq];
                            const uint32_t ph = (w + qq + 1 < W) ? B.VP[w + qq + 1 < W ? w + qq + 1 : w + qq] : 0u;
                            const uint32_t nh = (w + qq + 1 < W) ? B.VN[w + qq + 1 < W ? w + qq + 1 : w + qq] : 0u;
                            C.VP[w] = funnel_r(pl, ph, r);
                            C.VN[w] = funnel_r(nl, nh, r);
                        }
                    }
                }
                C.j = j; C.pos = pos - s; C.res = res; C.prow = prow + 4 * s;
                return Wn;
            }
        }
    }

    // tail: one column at a time; lanes finish when their target ends
    int pend = 0;                                     // columns since the last flush (warp-uniform)
    int tw_idx = -1;
    uint32_t tw = 0;
    for (; j <= nmax; ++j) {
        const int wi = (j - 1) >> 4;
        if (wi != tw_idx) { tw = tgt[wi * ts]; tw_idx = wi; }
        const uint32_t c = (tw >> (2 * ((j - 1) & 15))) & 3u;
        B.column(prow + c);
        prow += 4; ++pend;
        if (j == n && res == ED_PENDING) {
            const int d = B.value_at(pos, pend);
            res = d <= k ? d : -1;
            C.done_col = j;
        }
        if (pend == 32) {                // pend is warp-uniform
            B.flush(32); pend = 0;
            if (res == ED_PENDING && B.value_at(pos, 0) > k) { res = -1; C.done_col = j; }
            if (warp_all(res != ED_PENDING)) { C.j = j + 1; C.res = res; return 0; }
        }
    }
    C.j = nmax + 1; C.res = res;
    return 0;
}

// tab      : shifted match masks of the query; entry 4 * o + c holds bits [o, o + 32) of mask c (the layout
//            of the header comment, which is linear in the bit offset o); must be readable (zero) up to bit
//            padbits + m + 32 * (W + 1)
// padbits  : multiple of 32, >= dhi
// m        : query length (warp-uniform)
// tgt, ts  : this lane's 2-bit target stream (see band_group.cuh)
// n, k     : this lane's target length and threshold;  active: lane has a pair
// dhi      : THIS LANE's top diagonal: its window covers the diagonals [dhi - 32W + 1, dhi], which must
//            contain the lane's own strip.  Lanes may differ: the window position only enters through the
//            lane's table offset (lanes a few diagonals apart read neighbouring 16-byte entries, still one
//            shared-memory wavefront).  Inactive lanes pass any dhi in [0, padbits] not below the active ones'.
// Sets up the carry of a group at width W (columns 0 done).  Returns false when no lane has anything to walk.
template <int CAP>
ISO_HD bool diag_begin(DiagCarry<CAP>& C, int W, const uint32_t* __restrict__ tab, int padbits, int m, int n, int k,
                       bool active, int dhi, int* nmin, int* nmax) {
#pragma unroll
    for (int w = 0; w < CAP; ++w) {
        const int lo = dhi - 32 * w;   // bits below lo are rows <= 0 at column 0: delta -1
        const uint32_t vn = lo <= 0 ? 0u : (lo >= 32 ? 0xffffffffu : ((1u << lo) - 1u));
        C.VN[w] = vn; C.VP[w] = ~vn;
    }
    C.score = 32 * W - 1 - dhi;        // D[-(dhi - 32W + 1)][0]; the bottom diagonal is <= 0
    C.j = 1;
    C.pos = dhi - (n - m);
    C.done_col = 0;
    C.prow = tab + 4 * (padbits - dhi);
    int res = active ? ED_PENDING : -1;
    if (active && n == 0) res = (m <= k) ? m : -1;
    if (active && m == 0) res = (n <= k) ? n : -1;
    C.res = res;
    *nmax = warp_max(res == ED_PENDING ? n : 0);
    *nmin = warp_min(res == ED_PENDING ? n : 0x7fffffff);
    return *nmax != 0;
}

// One width, no shrinking: the arithmetic of a single instance (host tests of any W; the kernels use
// ed_group_diag_run below).
// cols, done_col : out, columns the warp walked / column at which this lane's own result was known
// returns  : edit distance if <= k, else -1   (inactive lanes: -1)
template <int W>
ISO_HD int ed_group_diag(const uint32_t* __restrict__ tab, int padbits, int m,
                         const uint32_t* __restrict__ tgt, int ts, int n, int k, bool active, int dhi, int* cols,
                         int* done_col) {
    DiagCarry<W> C;
    int nmin, nmax;
    *cols = 0; *done_col = 0;
    if (!diag_begin<W>(C, W, tab, padbits, m, n, k, active, dhi, &nmin, &nmax)) return C.res;
    diag_segment<W, W>(C, tgt, ts, n, k, nmin, nmax, 0);
    *cols = C.j - 1; *done_col = C.done_col;
    return C.res;
}

// The group walk of the kernels: starts at width diag_avail(Wd) and continues in narrower instances whenever
// the union of the lanes' alive intervals allows (narrow > 0).
// cols     : out, columns the warp walked
// wcols    : out, sum over the widths used of (window words x columns walked at that width)
// ucells   : out, this lane's share of the necessary work: for every column until ITS result was known, the rows
//            of its own strip (`rows`, the caller's Ukkonen strip for k) that the window in force still held
ISO_HD int ed_group_diag_run(int Wd, const uint32_t* __restrict__ tab, int padbits, int m,
                             const uint32_t* __restrict__ tgt, int ts, int n, int k, bool active, int dhi, int narrow,
                             int rows, int* cols, unsigned* wcols, unsigned* ucells) {
    DiagCarry<DIAG_WMAX> C;
    int nmin, nmax;
    *cols = 0; *wcols = 0u; *ucells = 0u;
    int W = diag_avail(Wd);
    if (!diag_begin<DIAG_WMAX>(C, W, tab, padbits, m, n, k, active, dhi, &nmin, &nmax)) return C.res;
    while (W > 0) {
        const int j0 = C.j;
        const bool was_pending = C.res == ED_PENDING;
        int Wn;
        switch (W) {
            case 1: Wn = diag_segment<1, DIAG_WMAX>(C, tgt, ts, n, k, nmin, nmax, narrow); break;
            case 2: Wn = diag_segment<2, DIAG_WMAX>(C, tgt, ts, n, k, nmin, nmax, narrow); break;
            case 3: Wn = diag_segment<3, DIAG_WMAX>(C, tgt, ts, n, k, nmin, nmax, narrow); break;
            case 4: Wn = diag_segment<4, DIAG_WMAX>(C, tgt, ts, n, k, nmin, nmax, narrow); break;
            case 5: Wn = diag_segment<5, DIAG_WMAX>(C, tgt, ts, n, k, nmin, nmax, narrow); break;
            case 6: Wn = diag_segment<6, DIAG_WMAX>(C, tgt, ts, n, k, nmin, nmax, narrow); break;
            case 7: Wn = diag_segment<7, DIAG_WMAX>(C, tgt, ts, n, k, nmin, nmax, narrow); break;
            case 8: Wn = diag_segment<8, DIAG_WMAX>(C, tgt, ts, n, k, nmin, nmax, narrow); break;
            case 9: Wn = diag_segment<9, DIAG_WMAX>(C, tgt, ts, n, k, nmin, nmax, narrow); break;
            case 10: Wn = diag_segment<10, DIAG_WMAX>(C, tgt, ts, n, k, nmin, nmax, narrow); break;
            case 11: Wn = diag_segment<11, DIAG_WMAX>(C, tgt, ts, n, k, nmin, nmax, narrow); break;
            case 12: Wn = diag_segment<12, DIAG_WMAX>(C, tgt, ts, n, k, nmin, nmax, narrow); break;
            case 13: Wn = diag_segment<13, DIAG_WMAX>(C, tgt, ts, n, k, nmin, nmax, narrow); break;
            default: Wn = diag_segment<14, DIAG_WMAX>(C, tgt, ts, n, k, nmin, nmax, narrow); break;
        }
        *wcols += (unsigned)(W * (C.j - j0));
        if (was_pending) {
            const int end = C.res != ED_PENDING ? C.done_col : C.j - 1;
            *ucells += (unsigned)((end - (j0 - 1)) * (rows < 32 * W ? rows : 32 * W));
        }
        W = Wn;
    }
    *cols = C.j - 1;
    return C.res;
}

}  // namespace isocon
