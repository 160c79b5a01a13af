"""CPU suite: the inequality behind level 1 of the two-level one-sided pass (DESIGN.md section 3.1, step 7;
``qgram_level1_kernel`` in isocon_b200/csrc/nn_kernels.cuh), restated in numpy with the kernel's own block size, hash
and sketch width, against exact edit distances of the oracle.

Claim: cut x into non-overlapping 8-mers ("blocks").  T edit operations touch at most T blocks, so ed(x, y) <= T leaves
at least (#blocks - T) blocks of x verbatim in y.  With bucket sets X (blocks of x) and Y (all 8-mers of y):
|X & Y| + T < |X|  implies  ed(x, y) > T.  The filter may therefore dismiss a pair only when that inequality holds --
never a pair within T.  The second half is the triangle inequality over a cluster's representative."""
import numpy as np

from isocon_b200 import workloads
from oracle import oracle as O

QG_Q, QG_BITS = 8, 8192           # nn_kernels.cuh: QG_Q, QG_BITS


def _buckets(codes, positions):
    """Bucket of the 8-mer at every given position: the kernel's qgram_bucket (base p in the low two bits of the k-mer,
    multiplicative hash, top 13 bits)."""
    codes = codes.astype(np.uint64)
    kmer = np.zeros(len(positions), dtype=np.uint64)
    for i in range(QG_Q):
        kmer |= codes[positions + i] << np.uint64(2 * i)
    return ((kmer * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)) >> np.uint64(32 - 13)


def _sets(x, y):
    X = set(_buckets(x, np.arange(0, (x.size // QG_Q) * QG_Q, QG_Q)).tolist())
    Y = set(_buckets(y, np.arange(0, y.size - QG_Q + 1)).tolist()) if y.size >= QG_Q else set()
    return X, Y


def _ed(a, b):
    return O.ed_plain(workloads._to_str(a).encode(), workloads._to_str(b).encode())


def test_the_count_never_dismisses_a_pair_within_the_threshold():
    rng = np.random.default_rng(11)
    assert int(_buckets(np.zeros(8, np.uint8), np.array([0]))[0]) < QG_BITS
    dismissed_strangers = strangers = 0
    for trial in range(300):
        L = int(rng.integers(40, 1500))
        tpl = rng.integers(0, 4, size=L, dtype=np.uint8)
        e1, e2 = float(rng.choice([0.0, 0.01, 0.03, 0.08])), float(rng.choice([0.0, 0.01, 0.03, 0.08]))
        x = workloads._mutate(rng, tpl, e1 / 3, e1 / 3, e1 / 3)
        y = workloads._mutate(rng, tpl, e2 / 3, e2 / 3, e2 / 3)
        if trial % 10 == 0:                       # a stranger of similar length
            y = rng.integers(0, 4, size=max(QG_Q, L + int(rng.integers(-20, 20))), dtype=np.uint8)
        if x.size < QG_Q or y.size < QG_Q:
            continue
        d = _ed(x, y)
        X, Y = _sets(x, y)
        common = len(X & Y)
        for T in (d, d + 1, d + 17):
            assert common + T >= len(X), (trial, L, d, T, common, len(X))      # within T: never dismissed
        if trial % 10 == 0:
            strangers += 1
            T = L // 20                            # a threshold of 5 % of the length
            if d > T and common + T < len(X):
                dismissed_strangers += 1
    assert strangers >= 20 and dismissed_strangers >= 0.9 * strangers          # and it does dismiss strangers


def test_triangle_inequality_over_a_representative():
    """d(q, rep) > k + radius  =>  d(q, c) > k for every member c with d(rep, c) <= radius (what level 1 concludes)."""
    rng = np.random.default_rng(12)
    for _ in range(40):
        L = int(rng.integers(60, 500))
        root = rng.integers(0, 4, size=L, dtype=np.uint8)
        members = [workloads._diverge(rng, root, float(rng.uniform(0.0, 0.03)), 1) for _ in range(6)]
        far = [[_ed(a, b) for b in members] for a in members]
        rep = int(np.argmin([max(row) for row in far]))                # 1-centre, like sketch_order
        radius = max(far[rep])
        q = workloads._mutate(rng, members[int(rng.integers(0, 6))] if rng.random() < 0.5 else
                              rng.integers(0, 4, size=L, dtype=np.uint8), 0.02, 0.02, 0.01)
        d_rep = _ed(q, members[rep])
        for c in members:
            assert _ed(q, c) >= d_rep - radius
