/*
 * oracle/nn_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (C++) of the two scan loops of IsoCon's nearest-neighbour graph:
 *
 *   nn_oracle_1set  follows /root/reference/modules/nearest_neighbor_graph.py:110-198
 *                   (get_nearest_neighbors) and its Pool sharding :19-82
 *   nn_oracle_2set  follows /root/reference/modules/nearest_neighbor_graph.py:341-424
 *                   (get_nearest_neighbors_2set) and its Pool sharding :300-334
 *
 * Arithmetic: oracle/levenshtein.h (ed_myers64; `use_plain` switches to the O(mn) DP).
 * The loops are re-stated statement by statement: same scan order (down i-j, then up
 * i+j), same sticky length pruning, same running best, same symmetric seeding dict
 * (`lower_target_edit_distances`, per worker call), same chunking rule
 * chunk = max(int(N / (10*cores)), 20).  Edges are emitted in the dict insertion order
 * of the reference so the Python side can rebuild identical dicts.
 *
 * Also accumulates the implementation-independent work counters of SURVEY.md §8(d):
 * calls, calls returning -1, cells_full = sum m*n, cells_band = sum n*min(m, w).
 *
 * PARITY STATUS: pinned against the unmodified reference driver by oracle/make_golden.py
 * (run in the authoring container; fixtures under tests/golden/).
 */
#include "levenshtein.h"

#include <atomic>
#include <thread>
#include <unordered_map>

using namespace isocon_oracle;

namespace {

struct Stats {
    uint64_t calls = 0, neg = 0, cells_full = 0, cells_band = 0;
    void add(const Stats& o) { calls += o.calls; neg += o.neg; cells_full += o.cells_full; cells_band += o.cells_band; }
};

struct Reads {
    const uint8_t* cat; const int64_t* off; int n;
    const uint8_t* seq(int i) const { return cat + off[i]; }
    int len(int i) const { return (int)(off[i + 1] - off[i]); }
};

struct Edge { int q, t, d; };

static inline void count(Stats& st, int m, int n, int k, int ed) {
    st.calls++;
    if (ed < 0) st.neg++;
    st.cells_full += (uint64_t)m * (uint64_t)n;
    const int ad = std::abs(n - m);
    if (ad <= k) {
        const int64_t w = (int64_t)ad + 2 * (int64_t)((k - ad) / 2) + 1;
        st.cells_band += (uint64_t)n * (uint64_t)std::min<int64_t>(m, w);
    }
}

/* One worker call of get_nearest_neighbors (:110-198) on queries [q0, q1). */
static void scan_1set(const Reads& R, const uint8_t* converged, int64_t depth, int q0, int q1,
                      bool use_plain, std::vector<Edge>& out, int32_t* best_out, Stats& st) {
    const int N = R.n;
    std::unordered_map<int, int> lower; /* lower_target_edit_distances (:112) */
    Myers64 M;
    std::vector<Edge> cur;
    for (int i = q0; i < q1; ++i) {
        const int m = R.len(i);
        best_out[i] = -1;
        if (converged && converged[i]) continue;                 /* :121-123 */
        int best;
        auto it = lower.find(i);
        best = (it != lower.end()) ? it->second : m;             /* :125-129 */
        M.set_query(R.seq(i), m);
        cur.clear();
        bool stop_up = false, stop_down = false;
        int64_t j = 1;
        for (;;) {
            if (i - j < 0) stop_down = true;                     /* :136-139 */
            if (i + j >= N) stop_up = true;
            if (!stop_down && std::abs(m - R.len((int)(i - j))) > best) stop_down = true; /* :145 */
            if (!stop_up && std::abs(m - R.len((int)(i + j))) > best) stop_up = true;     /* :152 */
            for (int dir = 0; dir < 2; ++dir) {                  /* down first, then up */
                if (dir == 0 ? stop_down : stop_up) continue;
                const int t = (int)(dir == 0 ? i - j : i + j);
                const int n = R.len(t);
                const int k = best;
                int ed = use_plain ? ed_plain(R.seq(i), m, R.seq(t), n) : M.distance(R.seq(t), n, k);
                if (use_plain && ed > k) ed = -1;
                count(st, m, n, k, ed);
                if (0 < ed && ed < best) {                       /* :157-160 */
                    best = ed; cur.clear(); cur.push_back({i, t, ed});
                } else if (ed == best) {                         /* :161-162 */
                    cur.push_back({i, t, ed});
                }
                if (ed > 0) {                                    /* :164-169 */
                    auto lt = lower.find(t);
                    if (lt == lower.end()) lower[t] = ed;
                    else if (ed < lt->second) lt->second = ed;
                }
            }
            if (stop_down && stop_up) break;                     /* :187 */
            if (j >= depth) break;                               /* :190 */
            ++j;
        }
        if (!cur.empty()) best_out[i] = best;
        out.insert(out.end(), cur.begin(), cur.end());
    }
}

/* One worker call of get_nearest_neighbors_2set (:341-424) on list entries [q0, q1). */
static void scan_2set(const Reads& R, const uint8_t* is_target, int64_t depth, int q0, int q1,
                      bool use_plain, std::vector<Edge>& out, int32_t* best_out, Stats& st) {
    const int N = R.n;
    Myers64 M;
    std::vector<Edge> cur;
    for (int i = q0; i < q1; ++i) {
        best_out[i] = -1;
        if (is_target[i]) continue;                              /* :350-351 */
        const int m = R.len(i);
        int best = m;                                            /* :356 */
        M.set_query(R.seq(i), m);
        cur.clear();
        bool stop_up = false, stop_down = false;
        int64_t processed = 0, j = 1;
        for (;;) {
            if (i - j < 0) stop_down = true;
            if (i + j >= N) stop_up = true;
            if (!stop_down && std::abs(m - R.len((int)(i - j))) > best) stop_down = true; /* :373 */
            if (!stop_up && std::abs(m - R.len((int)(i + j))) > best) stop_up = true;     /* :380 */
            for (int dir = 0; dir < 2; ++dir) {
                if (dir == 0 ? stop_down : stop_up) continue;
                const int t = (int)(dir == 0 ? i - j : i + j);
                if (!is_target[t]) continue;                     /* :383, :397 */
                ++processed;
                const int n = R.len(t);
                const int k = best;
                int ed = use_plain ? ed_plain(R.seq(i), m, R.seq(t), n) : M.distance(R.seq(t), n, k);
                if (use_plain && ed > k) ed = -1;
                count(st, m, n, k, ed);
                if (0 <= ed && ed < best) {                      /* :388-391 */
                    best = ed; cur.clear(); cur.push_back({i, t, ed});
                } else if (ed == best) {                         /* :394-395 */
                    cur.push_back({i, t, ed});
                }
            }
            if (stop_down && stop_up) break;                     /* :413 */
            if (processed >= depth) break;                       /* :416 */
            ++j;
        }
        if (!cur.empty()) best_out[i] = best;
        out.insert(out.end(), cur.begin(), cur.end());
    }
}

typedef void (*scan_fn)(const Reads&, const uint8_t*, int64_t, int, int, bool,
                        std::vector<Edge>&, int32_t*, Stats&);

/* Pool sharding of :30-66 / :312-319: chunk = max(int(N/(10*cores)), 20); cores==1 is one call. */
static int64_t run(scan_fn fn, const Reads& R, const uint8_t* mask, int64_t depth, int q0, int q1,
                   int cores, int threads, bool use_plain,
                   int32_t* best_out, int32_t* eq, int32_t* et, int32_t* ed, int64_t cap, uint64_t* stats) {
    std::vector<std::pair<int, int>> chunks;
    if (cores <= 1) chunks.push_back({q0, q1});
    else {
        const int chunk = std::max((int)(R.n / (10 * cores)), 20);
        for (int s = 0; s < R.n; s += chunk) {
            const int a = std::max(s, q0), b = std::min(std::min(s + chunk, R.n), q1);
            if (a < b) chunks.push_back({a, b});
        }
    }
    std::vector<std::vector<Edge>> outs(chunks.size());
    std::vector<Stats> sts(chunks.size());
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        for (;;) {
            const size_t c = next.fetch_add(1);
            if (c >= chunks.size()) break;
            fn(R, mask, depth, chunks[c].first, chunks[c].second, use_plain, outs[c], best_out, sts[c]);
        }
    };
    threads = std::max(1, std::min<int>(threads, (int)chunks.size()));
    if (threads == 1) worker();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    Stats total;
    int64_t ne = 0;
    for (size_t c = 0; c < chunks.size(); ++c) {
        total.add(sts[c]);
        for (const Edge& e : outs[c]) {
            if (ne < cap) { eq[ne] = e.q; et[ne] = e.t; ed[ne] = e.d; }
            ++ne;
        }
    }
    if (stats) { stats[0] = total.calls; stats[1] = total.neg; stats[2] = total.cells_full; stats[3] = total.cells_band; }
    return ne;
}

} /* namespace */

extern "C" {

int oracle_ed_plain(const uint8_t* x, int m, const uint8_t* y, int n) { return ed_plain(x, m, y, n); }
int oracle_ed_banded_dp(const uint8_t* x, int m, const uint8_t* y, int n, int k) { return ed_banded_dp(x, m, y, n, k); }
int oracle_ed_myers64(const uint8_t* x, int m, const uint8_t* y, int n, int k) { return ed_myers64(x, m, y, n, k); }

/* Batch of explicit pairs over a concatenated read set (for kernel-level parity tests). */
void oracle_ed_pairs(const uint8_t* cat, const int64_t* off, const int32_t* a, const int32_t* b,
                     const int32_t* k, int64_t n_pairs, int32_t* out) {
    Myers64 M;
    int cur = -1;
    for (int64_t p = 0; p < n_pairs; ++p) {
        if (a[p] != cur) { cur = a[p]; M.set_query(cat + off[cur], (int)(off[cur + 1] - off[cur])); }
        out[p] = M.distance(cat + off[b[p]], (int)(off[b[p] + 1] - off[b[p]]), k ? k[p] : -1);
    }
}

/*
 * The sorted list is given as concatenated bytes + offsets (n+1).  `converged[i]` != 0
 * marks entries whose sequence is in has_converged.  Queries are list entries
 * [q_start, q_start+q_count).  Returns the number of edges (may exceed cap: call again).
 * best_out[i] = final best of query i, or -1 when its dict is empty.
 */
int64_t nn_oracle_1set(const uint8_t* cat, const int64_t* off, int n, const uint8_t* converged,
                       int64_t depth, int q_start, int q_count, int cores, int threads, int use_plain,
                       int32_t* best_out, int32_t* eq, int32_t* et, int32_t* ed, int64_t cap, uint64_t* stats) {
    Reads R{cat, off, n};
    return run(scan_1set, R, converged, depth, q_start, q_start + q_count, cores, threads, use_plain != 0,
               best_out, eq, et, ed, cap, stats);
}

int64_t nn_oracle_2set(const uint8_t* cat, const int64_t* off, int n, const uint8_t* is_target,
                       int64_t depth, int q_start, int q_count, int cores, int threads, int use_plain,
                       int32_t* best_out, int32_t* eq, int32_t* et, int32_t* ed, int64_t cap, uint64_t* stats) {
    Reads R{cat, off, n};
    return run(scan_2set, R, is_target, depth, q_start, q_start + q_count, cores, threads, use_plain != 0,
               best_out, eq, et, ed, cap, stats);
}

} /* extern "C" */
