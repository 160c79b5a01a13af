// Final best[] of every read of a workload (length-sorted order) for tests/host_sim/sim_narrow.cpp; uses the oracle's Myers
// implementation (TEST INFRASTRUCTURE ONLY: it links the oracle).
#include "levenshtein.h"
#include <cstdio>
#include <fstream>
#include <string>
#include <thread>
#include <atomic>
using namespace isocon_oracle;
int main(int argc, char** argv) {
    std::ifstream f(argv[1]);
    std::vector<std::string> R; std::string s;
    while (std::getline(f, s)) if (!s.empty()) R.push_back(s);
    std::stable_sort(R.begin(), R.end(), [](const std::string& a, const std::string& b) { return a.size() < b.size(); });
    const int N = R.size();
    std::vector<int> best(N);
    std::atomic<int> next(0);
    auto w1 = [&]() {
        Myers64 M;
        for (;;) { int q = next++; if (q >= N) break; M.set_query((const uint8_t*)R[q].data(), R[q].size()); int b = R[q].size();
            // scan outward like the reference so that best falls quickly
            for (int j = 1; j < N; ++j) for (int sgn = -1; sgn <= 1; sgn += 2) { int t = q + sgn * j; if (t < 0 || t >= N) continue;
                int d = M.distance((const uint8_t*)R[t].data(), R[t].size(), b); if (d > 0 && d < b) b = d; }
            best[q] = b; }
    };
    { std::vector<std::thread> th; for (int i = 0; i < 8; ++i) th.emplace_back(w1); for (auto& t : th) t.join(); }
    FILE* o = fopen(argv[2], "w");
    for (int i = 0; i < N; ++i) fprintf(o, "%d\n", best[i]);
    fclose(o);
}
