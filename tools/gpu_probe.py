"""Quick device probe: INT32 peak, timing and work counters of a config at a given scale."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from isocon_b200 import _binding, workloads

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
    sym = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    ctx = _binding.NNContext(0)
    print("int32 peak lane-ops/s: %.4g" % ctx.int32_peak(), flush=True)
    t0 = time.time()
    if name == "c5":
        X, C = workloads.config5(scale=scale)
        L = sorted([(s, a) for a, s in X.items()] + [(s, a) for a, s in C.items()], key=lambda e: len(e[0]))
        ist = np.array([1 if a in C else 0 for _, a in L], dtype=np.uint8); isq = 1 - ist; mode = 2
    else:
        S = workloads.CONFIGS[name](scale=scale)
        L = sorted(((s, a) for a, s in S.items()), key=lambda e: len(e[0]))
        isq = np.ones(len(L), np.uint8); ist = None; mode = 1
    print("generated %d entries in %.1fs" % (len(L), time.time() - t0), flush=True)
    for rep in range(3):
        t0 = time.time()
        ctx.set_reads([s for s, _ in L])
        t1 = time.time()
        best, eq, et, ed = ctx.graph(mode, 2 ** 32, isq, ist, _binding.ALGO_TILE, bool(sym))
        t2 = time.time()
        st = ctx.stats()
        ms = ctx.last_ms(1)
        print("rep %d: set_reads %.3fs graph wall %.3fs kernel %.1f ms edges %d best[mean] %.1f stats %s" % (
            rep, t1 - t0, t2 - t1, ms, eq.size, float(best.mean()), st), flush=True)
        wc = st["word_columns"] * 32
        print("   lane word-columns/s %.4g  (x11 instr = %.4g lane-ops/s)" % (wc / (ms * 1e-3), 11 * wc / (ms * 1e-3)), flush=True)

main()
