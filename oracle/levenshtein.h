/*
 * oracle/levenshtein.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the arithmetic behind IsoCon's hot path:
 *     edlib.align(x, y, mode="NW", task="distance", k=K)["editDistance"]
 * as called from /root/reference/modules/nearest_neighbor_graph.py:104-107
 * (call sites :156, :172, :387, :403).
 *
 * edlib itself (PyPI `edlib`, pinned >=1.1.2 in requirements.txt:1, tested ==1.2.1 per
 * docs/version_history.txt:6) is a third-party C++ library that is neither vendored in
 * /root/reference nor installed here.  What the call computes is the unit-cost global
 * Levenshtein distance over raw characters, reported as -1 when it exceeds k.  Any exact
 * algorithm is bit-identical by definition, so this header carries THREE independent
 * implementations that are cross-validated against each other in tests/:
 *
 *   ed_plain      full O(mn) two-row dynamic programme (the definition)
 *   ed_banded_dp  Ukkonen diagonal-strip DP with threshold k
 *   ed_myers64    Myers (1999) / Hyyro (2003) block bit-vector algorithm, 64-bit words,
 *                 Ukkonen band, exact early exit -- the published algorithm family of
 *                 edlib; used as the CPU baseline arithmetic.
 *
 * PARITY STATUS: the reference ships no golden vectors for this path (SURVEY.md §4, §8c:
 * "parity unpinned" at the edlib boundary).  The oracle is pinned instead against
 * outputs of the reference's own unmodified driver run in the authoring container
 * (oracle/make_golden.py -> tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use anything in oracle/.
 */
#ifndef ISOCON_ORACLE_LEVENSHTEIN_H
#define ISOCON_ORACLE_LEVENSHTEIN_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace isocon_oracle {

/* The definition: D[i][j] = min(D[i-1][j]+1, D[i][j-1]+1, D[i-1][j-1]+(x_i!=y_j)). */
static inline int ed_plain(const uint8_t* x, int m, const uint8_t* y, int n) {
    std::vector<int> prev(m + 1), cur(m + 1);
    for (int i = 0; i <= m; ++i) prev[i] = i;
    for (int j = 1; j <= n; ++j) {
        cur[0] = j;
        const uint8_t c = y[j - 1];
        for (int i = 1; i <= m; ++i) {
            int v = prev[i - 1] + (x[i - 1] != c);
            v = std::min(v, prev[i] + 1);
            v = std::min(v, cur[i - 1] + 1);
            cur[i] = v;
        }
        std::swap(prev, cur);
    }
    return prev[m];
}

/* Ukkonen strip: only diagonals d = j - i in [min(0,D)-p, max(0,D)+p], p = (k-|D|)/2. */
static inline int ed_banded_dp(const uint8_t* x, int m, const uint8_t* y, int n, int k) {
    if (k < 0) return ed_plain(x, m, y, n);
    const int delta = n - m;
    if (std::abs(delta) > k) return -1;
    const int p = (k - std::abs(delta)) / 2;
    const int dmin = std::min(0, delta) - p, dmax = std::max(0, delta) + p;
    const int INF = 1 << 29;
    std::vector<int> prev(m + 2, INF), cur(m + 2, INF);
    for (int i = 0; i <= m && -i >= dmin; ++i) prev[i] = i; /* column 0: d = -i */
    for (int j = 1; j <= n; ++j) {
        const int ilo = std::max(0, j - dmax), ihi = std::min(m, j - dmin);
        if (ilo > 0) cur[ilo - 1] = INF;
        for (int i = ilo; i <= ihi; ++i) {
            int v;
            if (i == 0) v = j;
            else {
                v = prev[i - 1] + (x[i - 1] != y[j - 1]);
                v = std::min(v, prev[i] + 1);
                v = std::min(v, cur[i - 1] + 1);
            }
            cur[i] = std::min(v, INF);
        }
        if (ihi < m) cur[ihi + 1] = INF;
        /* The strip moves down by at most one row per column, so the only cells of the
           previous column read outside its own strip are ilo-1 and ihi+1, poisoned above. */
        std::swap(prev, cur);
    }
    const int d = prev[m];
    return d <= k ? d : -1;
}

/*
 * Block bit-vector algorithm, 64-bit words.  Column-wise over y (the "target"), the
 * bit-vectors run down x (the "query").  Pv/Mv are the +1/-1 vertical deltas, score[b]
 * the value at the bottom row of block b.  Band rules follow SURVEY.md Appendix C.2.
 * k < 0 means unbounded.
 */
struct Myers64 {
    std::vector<uint64_t> peq;   /* [sym_id][nb] */
    std::vector<uint64_t> Pv, Mv;
    std::vector<int> score;
    int sym_id[256];
    int m = 0, nb = 0, nsym = 0;
    const uint8_t* x = nullptr;

    void set_query(const uint8_t* q, int len) {
        x = q; m = len; nb = (m + 63) / 64;
        for (int c = 0; c < 256; ++c) sym_id[c] = -1;
        nsym = 0;
        for (int i = 0; i < m; ++i) if (sym_id[q[i]] < 0) sym_id[q[i]] = nsym++;
        peq.assign((size_t)(nsym + 1) * std::max(nb, 1), 0); /* last row: all-zero */
        for (int i = 0; i < m; ++i)
            peq[(size_t)sym_id[q[i]] * nb + (i >> 6)] |= (uint64_t)1 << (i & 63);
        Pv.resize(std::max(nb, 1)); Mv.resize(std::max(nb, 1)); score.resize(std::max(nb, 1));
    }

    int distance(const uint8_t* y, int n, int k) {
        if (k < 0) k = std::max(m, n);
        const int delta = n - m;
        if (std::abs(delta) > k) return -1;
        if (m == 0) return n; /* n <= k by the check above */
        const int p = (k - std::abs(delta)) / 2;
        const int dmin = std::min(0, delta) - p, dmax = std::max(0, delta) + p;
        int last = std::min(nb - 1, (std::min(m, 0 - dmin) - 1) / 64);
        if (std::min(m, -dmin) < 1) last = 0;
        for (int b = 0; b <= last; ++b) { Pv[b] = ~(uint64_t)0; Mv[b] = 0; score[b] = (b + 1) * 64; }
        for (int j = 1; j <= n; ++j) {
            const int first = std::max(0, (std::max(1, j - dmax) - 1) / 64);
            const int nl = std::min(nb - 1, (std::min(m, j - dmin) - 1) / 64);
            while (last < nl) { /* at most one new block per column, at the bottom */
                ++last;
                Pv[last] = ~(uint64_t)0; Mv[last] = 0; score[last] = score[last - 1] + 64;
            }
            const int id = sym_id[y[j - 1]];
            const uint64_t* eqrow = &peq[(size_t)(id < 0 ? nsym : id) * nb];
            int hin = 1; /* NW top boundary, or an upper bound once block 0 left the strip */
            for (int b = first; b <= last; ++b) {
                uint64_t Eq = eqrow[b];
                const uint64_t pv = Pv[b], mv = Mv[b];
                const uint64_t Xv = Eq | mv;
                if (hin < 0) Eq |= 1;
                const uint64_t Xh = (((Eq & pv) + pv) ^ pv) | Eq;
                uint64_t Ph = mv | ~(Xh | pv);
                uint64_t Mh = pv & Xh;
                int hout = 0;
                if (Ph >> 63) hout = 1; else if (Mh >> 63) hout = -1;
                Ph <<= 1; Mh <<= 1;
                if (hin < 0) Mh |= 1; else if (hin > 0) Ph |= 1;
                Pv[b] = Mh | ~(Xv | Ph);
                Mv[b] = Ph & Xv;
                score[b] += hout;
                hin = hout;
            }
            /* Exact early exit: values never decrease along a diagonal, and the best
               completion from column j is exactly the cell on the final diagonal
               (row r = j - delta).  Checked every 16 columns. */
            if ((j & 15) == 0) {
                const int r = j - delta;
                if (r >= 1) {
                    const int b = (r - 1) / 64;
                    if (b >= first && b <= last) {
                        const int bit = (r - 1) & 63;
                        const uint64_t above = bit == 63 ? 0 : (~(uint64_t)0 << (bit + 1));
                        const int d = score[b] - __builtin_popcountll(Pv[b] & above)
                                               + __builtin_popcountll(Mv[b] & above);
                        if (d > k) return -1;
                    }
                }
            }
        }
        const int bit = (m - 1) & 63;
        const uint64_t pad = bit == 63 ? 0 : (~(uint64_t)0 << (bit + 1));
        const int d = score[nb - 1] - __builtin_popcountll(Pv[nb - 1] & pad)
                                    + __builtin_popcountll(Mv[nb - 1] & pad);
        return d <= k ? d : -1;
    }
};

static inline int ed_myers64(const uint8_t* x, int m, const uint8_t* y, int n, int k) {
    Myers64 M;
    M.set_query(x, m);
    return M.distance(y, n, k);
}

} /* namespace isocon_oracle */
#endif
