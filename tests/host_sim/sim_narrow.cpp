// TEST INFRASTRUCTURE ONLY.  Work model of the shrinking window (CPU, plain DP values): for sampled queries of a workload and the
// class-binned groups of 32 targets, the word-columns a warp walks with fixed windows vs windows that shrink
//   default            whole-word drops, per lane top or bottom
//   BITSHIFT=1         per-lane windows re-centred at bit granularity (upper bound of what shrinking can give)
//   UNIFORM=<G>        warp-uniform shift from the union of the lanes' alive intervals at G-bit granularity (shipped: 16)
//   ADAPT=<f>          adaptive check schedule instead of a fixed interval
//   g++ -O2 -march=native -pthread -o sim_narrow tests/host_sim/sim_narrow.cpp
//   g++ -O3 -march=native -pthread -I oracle -o sim_best_all tests/host_sim/sim_best_all.cpp
//   ./sim_best_all reads.txt best.txt; UNIFORM=16 ./sim_narrow reads.txt best.txt <queries> <every n-th group> <check interval in columns>
// reads.txt: one read per line (e.g. "\n".join(workloads.config2().values())); best.txt: final best[] per length-sorted read.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>
#include <algorithm>
#include <thread>
#include <atomic>
#include <random>
#include <mutex>
static const int INF = 1 << 28;
struct Lane {
    bool need, pending; int n, k, dhi, pos, W; std::vector<int> v; const std::string* t; int done_col;
};
int main(int argc, char** argv) {
    std::ifstream f(argv[1]); std::ifstream fb(argv[2]);
    int nq = atoi(argv[3]); int gstep = atoi(argv[4]); int every = argc > 5 ? atoi(argv[5]) : 32;
    std::vector<std::string> R; std::string s;
    while (std::getline(f, s)) if (!s.empty()) R.push_back(s);
    std::stable_sort(R.begin(), R.end(), [](const std::string& a, const std::string& b) { return a.size() < b.size(); });
    const int N = R.size();
    std::vector<int> best(N); for (int i = 0; i < N; ++i) fb >> best[i];
    // class bins
    std::vector<int> order(N); for (int i = 0; i < N; ++i) order[i] = i;
    auto cls = [&](int i) { return (std::min(best[i], 400) + 32) >> 5; };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cls(a) < cls(b); });
    std::vector<std::vector<int>> groups;
    for (int i = 0; i < N;) { int c = cls(order[i]); int j = i; while (j < N && cls(order[j]) == c) ++j;
        for (int g = i; g < j; g += 32) groups.emplace_back(order.begin() + g, order.begin() + std::min(j, g + 32)); i = j; }
    std::mt19937 rng(7);
    std::vector<int> Q; for (int i = 0; i < nq; ++i) Q.push_back(rng() % N);
    std::atomic<int> next(0);
    std::atomic<long> nchecks_total(0), cost_cur(0), cost_new(0), cost_ideal(0), ngroups(0), npairs(0);
    long histW[16][16] = {{0}}; std::mutex mu;
    auto work = [&]() {
        for (;;) {
            int a = next++; if (a >= (int)(Q.size() * groups.size())) break;
            int qi = a / groups.size(), gi = a % groups.size();
            if (gi % gstep) continue;
            int q = Q[qi]; const std::string& x = R[q]; int m = x.size();
            std::vector<Lane> L(32);
            int kmax = -1; bool any = false;
            for (int l = 0; l < 32; ++l) {
                Lane& ln = L[l]; ln.need = false; ln.pending = false;
                if (l >= (int)groups[gi].size()) continue;
                int t = groups[gi][l]; if (t <= q) continue;   // symmetric: done from the smaller row
                ln.t = &R[t]; ln.n = R[t].size(); ln.k = std::min(std::max(best[q], best[t]), 400);
                if (std::abs(ln.n - m) > ln.k) continue;
                ln.need = true; ln.pending = true; kmax = std::max(kmax, ln.k); any = true;
            }
            if (!any) continue;
            int W0 = (kmax + 32) >> 5;
            int nmax = 0;
            for (auto& ln : L) if (ln.need) {
                int d = ln.n - m, dl = std::abs(d);
                ln.dhi = std::max(0, d) + ((kmax - dl) >> 1); ln.pos = ln.dhi - d; ln.W = W0;
                ln.v.assign(32 * W0, 0);
                for (int b = 0; b < 32 * W0; ++b) { int row = 0 - ln.dhi + b; ln.v[b] = row <= 0 ? -row : row; }  // column 0: D[i][0] = i, virtual rows j - i = -i
                nmax = std::max(nmax, ln.n); npairs++;
            }
            int W = W0; long cnew = 0, cideal = 0; int cols = 0; int next_check = 64; long nchecks = 0; int adaptive = getenv("ADAPT") ? 1 : 0; double safety = getenv("ADAPT") ? atof(getenv("ADAPT")) : 0;
            std::vector<int> nv;
            for (int j = 1; j <= nmax; ++j) {
                for (auto& ln : L) if (ln.need) {   // lanes keep computing in lock step even when done (cost only)
                    if (!ln.pending) continue;
                    int Wb = 32 * ln.W; nv.assign(Wb, 0);
                    char c = j <= ln.n ? (*ln.t)[j - 1] : 'N';
                    int top = j - ln.dhi;
                    for (int b = 0; b < Wb; ++b) {
                        int row = top + b;
                        // prev column window: row r was bit b+1
                        int diag = ln.v[b];                          // D[row-1][j-1]: prev bit of row-1 = (row-1) - (top-1) = b
                        int left = b + 1 < Wb ? ln.v[b + 1] : INF;   // D[row][j-1]
                        int up = b > 0 ? nv[b - 1] : INF;            // D[row-1][j]
                        int val;
                        if (row <= 0) val = j - row;
                        else { bool match = row <= m && x[row - 1] == c; val = std::min(std::min(diag + (match ? 0 : 1), left + 1), up + 1); }
                        nv[b] = val;
                    }
                    ln.v.swap(nv);
                }
                cnew += W; cols = j;
                for (auto& ln : L) if (ln.need && ln.pending) cideal += ln.W;  // placeholder
                bool boundary = adaptive ? (j == next_check) : (j % every) == 0;
                for (auto& ln : L) if (ln.need && ln.pending) {
                    if (j == ln.n) { ln.pending = false; ln.done_col = j; }
                    else if ((j % 32) == 0 && ln.v[ln.pos] > ln.k) { ln.pending = false; ln.done_col = j; }
                }
                bool anyp = false; for (auto& ln : L) if (ln.need && ln.pending) anyp = true;
                if (!anyp) break;
                if (boundary && getenv("BITSHIFT")) {
                    int Wn = 1;
                    for (auto& ln : L) if (ln.need && ln.pending) {
                        int Wb = 32 * W, blo = Wb, bhi = -1;
                        for (int b = 0; b < Wb; ++b) if (ln.v[b] + std::abs(b - ln.pos) <= ln.k) { blo = std::min(blo, b); bhi = b; }
                        if (bhi < 0) { blo = bhi = ln.pos; }
                        blo = std::min(blo, ln.pos); bhi = std::max(bhi, ln.pos);
                        Wn = std::max(Wn, (bhi - blo + 1 + 31) / 32);
                    }
                    if (Wn < W) {
                        for (auto& ln : L) if (ln.need && ln.pending) {
                            int Wb = 32 * W, blo = Wb, bhi = -1;
                            for (int b = 0; b < Wb; ++b) if (ln.v[b] + std::abs(b - ln.pos) <= ln.k) { blo = std::min(blo, b); bhi = b; }
                            if (bhi < 0) { blo = bhi = ln.pos; }
                            blo = std::min(blo, ln.pos); bhi = std::max(bhi, ln.pos);
                            int slack = 32 * Wn - (bhi - blo + 1);
                            int s = blo - slack / 2; s = std::max(0, std::min(s, 32 * (W - Wn)));
                            std::vector<int> nv2(ln.v.begin() + s, ln.v.begin() + s + 32 * Wn);
                            ln.v.swap(nv2); ln.dhi -= s; ln.pos -= s; ln.W = Wn;
                        }
                        W = Wn;
                    }
                } else if (boundary && getenv("UNIFORM")) {
                    // warp-uniform shift: union of the lanes' feasible intervals (window bits), granularity G
                    int G = atoi(getenv("UNIFORM"));
                    int BLO = 32 * W, BHI = -1;
                    for (auto& ln : L) if (ln.need && ln.pending) {
                        int Wb = 32 * W, blo = Wb, bhi = -1;
                        for (int b = 0; b < Wb; ++b) if (ln.v[b] + std::abs(b - ln.pos) <= ln.k) { blo = std::min(blo, b); bhi = b; }
                        if (bhi < 0) { blo = bhi = ln.pos; }
                        blo = std::min(blo, ln.pos); bhi = std::max(bhi, ln.pos);
                        blo = blo / G * G; bhi = (bhi / G + 1) * G - 1;
                        BLO = std::min(BLO, blo); BHI = std::max(BHI, bhi);
                    }
                    int Wn = (BHI - BLO + 1 + 31) / 32; nchecks_total++;
                    if (adaptive) { int Wt = std::min(Wn, W); int U = BHI - BLO + 1; int needbits = U - 32 * (Wt - 1); double rho = std::max(0.03, (double)(32 * W0 - U) / j); int dt = (int)(needbits / rho * safety); dt = std::max(32, (dt + 16) / 32 * 32); next_check = j + dt; }
                    if (Wn < W) {
                        int slack = 32 * Wn - (BHI - BLO + 1);
                        int s = BLO - slack / 2; s = std::max(0, std::min(s, 32 * (W - Wn)));
                        for (auto& ln : L) if (ln.need && ln.pending) {
                            std::vector<int> nv2(ln.v.begin() + s, ln.v.begin() + s + 32 * Wn);
                            ln.v.swap(nv2); ln.dhi -= s; ln.pos -= s; ln.W = Wn;
                        }
                        W = Wn;
                    }
                } else if (boundary) {
                    for (;;) {
                        if (W <= 1) break;
                        bool all = true;
                        for (auto& ln : L) if (ln.need && ln.pending) {
                            bool ct = ln.pos >= 32 && ln.v[32] + (ln.pos - 32) > ln.k;
                            int bb = 32 * (W - 1) - 1;
                            bool cb = ln.pos <= bb && ln.v[bb] + (bb - ln.pos) > ln.k;
                            if (!ct && !cb) { all = false; break; }
                        }
                        if (!all) break;
                        for (auto& ln : L) if (ln.need && ln.pending) {
                            bool ct = ln.pos >= 32 && ln.v[32] + (ln.pos - 32) > ln.k;
                            if (ct) { ln.v.erase(ln.v.begin(), ln.v.begin() + 32); ln.dhi -= 32; ln.pos -= 32; }
                            else ln.v.resize(32 * (W - 1));
                            ln.W = W - 1;
                        }
                        --W;
                    }
                }
            }
            cost_cur += (long)W0 * cols; cost_new += cnew; ngroups++;
            { std::lock_guard<std::mutex> g(mu); histW[W0][W]++; }
        }
    };
    { std::vector<std::thread> th; for (int i = 0; i < 8; ++i) th.emplace_back(work); for (auto& t : th) t.join(); }
    printf("checks/group %.1f ", (double)nchecks_total / ngroups); printf("groups %ld pairs %ld word-columns: current %ld, narrowing %ld (ratio %.3f)\n", ngroups.load(), npairs.load(), cost_cur.load(), cost_new.load(), (double)cost_new / cost_cur);
    for (int a = 1; a < 16; ++a) { bool any = false; for (int b = 0; b < 16; ++b) if (histW[a][b]) any = true; if (!any) continue; printf("W0=%d final W:", a); for (int b = 1; b <= a; ++b) printf(" %d:%ld", b, histW[a][b]); printf("\n"); }
}
