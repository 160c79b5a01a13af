"""TEST INFRASTRUCTURE ONLY (uses the oracle): CPU experiment behind DESIGN.md §8a "unrelated pairs".

An exact pre-filter for pairs that cannot be within k: ed(x, y) <= k implies that x and y share at least
max(m, n) - q + 1 - k q q-grams (Jokinen & Ukkonen 1991).  With one hashed q-gram bitset B per read,
popc(B_x & B_y) + duplicates(x) is an upper bound of the shared count (duplicates(x) = q-gram positions of x beyond the
first of their hash bucket), so a pair below the bound can be skipped without alignment.

    python tests/host_sim/exp_qgram_filter.py [scale]

Round 1, c5 at scale 0.02 (2000 reads x 100 candidates of 10 families, k = 127): 90.0 % of the pairs rejected -- exactly
the cross-family ones -- for q in 8..16 and 32768-bit sets (89.8 % with 8192 bits), no pair within k rejected.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from isocon_b200 import workloads  # noqa: E402
from oracle import oracle as O  # noqa: E402

CODE = np.zeros(256, np.uint64)
for i, ch in enumerate(b"ACGT"):
    CODE[ch] = i


def bitset(s, q, m):
    a = CODE[np.frombuffer(s.encode(), np.uint8)]
    n = len(a) - q + 1
    v = np.zeros(n, np.uint64)
    for j in range(q):
        v = (v << np.uint64(2)) | a[j:j + n]
    h = (v * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(64 - int(np.log2(m)))
    b = np.zeros(m, bool)
    b[h.astype(np.int64)] = True
    return b, n - int(b.sum())


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
    X, C = workloads.config5(scale=scale)
    xs, cs = list(X.values()), list(C.values())
    lens = np.array([len(c) for c in cs])
    k = 127
    for q, m in ((12, 8192), (12, 32768), (8, 32768), (16, 32768)):
        BX = [bitset(s, q, m) for s in xs[:300]]
        MC = np.array([bitset(s, q, m)[0] for s in cs])
        rej = tot = wrong = 0
        for xi, (bx, dup) in enumerate(BX):
            shared = (MC & bx).sum(axis=1) + dup
            need = np.maximum(lens, len(xs[xi])) - q + 1 - k * q
            r = (shared < need) & (np.abs(lens - len(xs[xi])) <= k)
            rej += int(r.sum()); tot += int((np.abs(lens - len(xs[xi])) <= k).sum())
            for ci in np.flatnonzero(r)[:5]:        # a rejected pair must be farther than k
                wrong += O.ed_myers64(xs[xi].encode(), cs[ci].encode(), k) >= 0
        print("q %2d, %5d bits: %.4f of %d pairs rejected, %d wrongly" % (q, m, rej / tot, tot, wrong))


if __name__ == "__main__":
    main()
