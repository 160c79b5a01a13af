"""Inputs for the nearest-neighbour-graph path: FASTA reading, the "round-1 call" that
``graphs.py`` makes, and the seeded synthetic read sets of BASELINE.json configs 2-5
(shapes defined in SURVEY.md §8d).

Nothing here touches the GPU; it only builds the ``dict[acc -> seq]`` inputs that the
reference-facing functions in ``isocon_b200.nearest_neighbor_graph`` take.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def read_fasta(path):
    """``{accession: sequence}`` in file order; accession spaces become ``_`` like
    /root/reference/modules/input_output/fasta_parser.py:1-19."""
    out = {}
    acc, chunks = None, []
    with open(path) as fh:
        for line in fh:
            if line.startswith(">"):
                if acc is not None:
                    out[acc] = "".join(chunks)
                acc = line[1:].strip().replace(" ", "_")
                chunks = []
            else:
                chunks.append(line.strip())
    if acc is not None:
        out[acc] = "".join(chunks)
    return out


def round1_call(S):
    """The arguments ``graphs.construct_exact_nearest_neighbor_graph`` hands to the hot path
    (/root/reference/modules/graphs.py:37-58): unique sequences keyed by the LAST accession that
    carries them, and ``has_converged`` = sequences with multiplicity > 1."""
    mult = {}
    for acc, seq in S.items():
        mult[seq] = mult.get(seq, 0) + 1
    has_converged = {seq for seq, c in mult.items() if c > 1}
    unique_strings = {seq: acc for acc, seq in S.items()}
    S_prime = {acc: seq for seq, acc in unique_strings.items()}
    return S_prime, has_converged


# --------------------------------------------------------------------------- synthetic reads

def _to_str(codes):
    return _ACGT[codes].tobytes().decode("ascii")


def _mutate(rng, tpl, p_ins, p_del, p_sub):
    """One read from template codes: per base delete / substitute, then maybe insert after it."""
    n = tpl.size
    u = rng.random(n)
    keep = u >= p_del
    sub = (u >= p_del) & (u < p_del + p_sub)
    base = tpl.copy()
    base[sub] = (base[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
    ins = rng.random(n) < p_ins
    counts = keep.astype(np.int64) + ins.astype(np.int64)
    total = int(counts.sum())
    out = np.empty(total, dtype=np.uint8)
    ends = np.cumsum(counts)
    starts = ends - counts
    out[starts[keep]] = base[keep]
    out[(ends - 1)[ins]] = rng.integers(0, 4, size=int(ins.sum()), dtype=np.uint8)
    return out


def _diverge(rng, tpl, frac_sub, n_indel=0):
    out = tpl.copy()
    nsub = int(round(frac_sub * tpl.size))
    if nsub:
        pos = rng.choice(tpl.size, size=nsub, replace=False)
        out[pos] = (out[pos] + rng.integers(1, 4, size=nsub, dtype=np.uint8)) & 3
    for _ in range(n_indel):
        p = int(rng.integers(1, out.size - 4))
        ln = int(rng.integers(1, 4))
        if rng.random() < 0.5:
            out = np.delete(out, slice(p, p + ln))
        else:
            out = np.insert(out, p, rng.integers(0, 4, size=ln, dtype=np.uint8))
    return out


def _reads_from(rng, templates, picks, p_ins, p_del, p_sub, prefix="r", trunc=None):
    S = {}
    seen = set()
    for idx, c in enumerate(picks):
        while True:
            r = _mutate(rng, templates[c], p_ins, p_del, p_sub)
            if trunc is not None and rng.random() < trunc[0]:
                r = r[: int(rng.integers(trunc[1], trunc[2]))]
            s = _to_str(r)
            if s not in seen:
                break
        seen.add(s)
        S["%s%d" % (prefix, idx)] = s
    return S


def config2(scale=1.0, seed=2):
    """10k reads x 1.5 kb, 20 near-identical gene copies, 5 % indel-heavy error (ins:del:sub 50:30:20)."""
    rng = np.random.default_rng(seed)
    n = max(int(round(10000 * scale)), 32)
    root = rng.integers(0, 4, size=1500, dtype=np.uint8)
    copies = [_diverge(rng, root, 0.005, 2) for _ in range(20)]
    picks = rng.integers(0, 20, size=n)
    return _reads_from(rng, copies, picks, 0.05 * 0.5, 0.05 * 0.3, 0.05 * 0.2)


def config3(scale=1.0, seed=3):
    """50k Iso-Seq-like reads x 3 kb, 100 paralogs 0.5-2 % apart (random tree), 2 % error (40:40:20)."""
    rng = np.random.default_rng(seed)
    n = max(int(round(50000 * scale)), 32)
    root = rng.integers(0, 4, size=3000, dtype=np.uint8)
    paralogs = [root]
    while len(paralogs) < 100:
        parent = paralogs[int(rng.integers(0, len(paralogs)))]
        paralogs.append(_diverge(rng, parent, float(rng.uniform(0.005, 0.02)), 1))
    w = rng.lognormal(0.0, 1.0, size=100)
    picks = rng.choice(100, size=n, p=w / w.sum())
    return _reads_from(rng, paralogs, picks, 0.02 * 0.4, 0.02 * 0.4, 0.02 * 0.2)


def config4(scale=1.0, seed=4, truncated=0.05):
    """200k ONT-like amplicon reads x 1 kb, 50 templates 1-5 % apart, 10 % error (30:35:35),
    5 % truncated reads (length U(600,1000)) to stress the length window."""
    rng = np.random.default_rng(seed)
    n = max(int(round(200000 * scale)), 32)
    root = rng.integers(0, 4, size=1000, dtype=np.uint8)
    tpls = [_diverge(rng, root, float(rng.uniform(0.01, 0.05)) / 2, 1) for _ in range(50)]
    picks = rng.integers(0, 50, size=n)
    return _reads_from(rng, tpls, picks, 0.10 * 0.30, 0.10 * 0.35, 0.10 * 0.35,
                       trunc=(truncated, 600, 1000) if truncated else None)


def config5(scale=1.0, seed=5):
    """2-set: 5k candidates x 2.5 kb (500 families x 10 members, 0.2-2 % divergence) and
    100k reads = candidate + 3 % error.  Returns (X, C)."""
    rng = np.random.default_rng(seed)
    n_fam = max(int(round(500 * scale)), 2)
    n_reads = max(int(round(100000 * scale)), 32)
    C = {}
    cands = []
    seen = set()
    for f in range(n_fam):
        root = rng.integers(0, 4, size=2500, dtype=np.uint8)
        for k in range(10):
            while True:
                c = root if k == 0 else _diverge(rng, root, float(rng.uniform(0.002, 0.02)), 1)
                s = _to_str(c)
                if s not in seen:
                    break
            seen.add(s)
            cands.append(c)
            C["c%d" % len(C)] = s
    picks = rng.integers(0, len(cands), size=n_reads)
    X = _reads_from(rng, cands, picks, 0.03 * 0.4, 0.03 * 0.4, 0.03 * 0.2)
    return X, C


CONFIGS = {"c2": config2, "c3": config3, "c4": config4, "c5": config5}
