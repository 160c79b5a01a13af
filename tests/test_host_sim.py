"""CPU suite, part 2: the per-lane band arithmetic of the CUDA kernels (myers_band.cuh /
band_group.cuh), compiled for the host by tests/host_sim, against the oracle.  This checks the
window / strip / early-exit logic without a GPU; it is never part of the product path."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import util
from isocon_b200 import workloads
from oracle import oracle as O

SIM_DIR = os.path.join(util.ROOT, "tests", "host_sim")


@pytest.fixture(scope="module")
def sim():
    so = os.path.join(SIM_DIR, "libsim_band.so")
    src = os.path.join(SIM_DIR, "sim_band.cpp")
    hdrs = [os.path.join(util.ROOT, "isocon_b200", "csrc", h) for h in ("myers_band.cuh", "band_group.cuh", "diag_band.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-x", "c++", "-o", so, src])
    L = ctypes.CDLL(so)
    L.sim_ed.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int] + [ctypes.c_int] * 4
    L.sim_ed_diag.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int] + [ctypes.c_int] * 5
    L.sim_ed_diag_run.argtypes = ([ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int] + [ctypes.c_int] * 8
                                  + [ctypes.POINTER(ctypes.c_int)])
    L.sim_warp_diag_run.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int),
                                    ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                    ctypes.POINTER(ctypes.c_int)]
    return L


def test_band_arithmetic_matches_oracle(sim):
    rng = np.random.default_rng(5)
    checked = 0
    for _ in range(1500):
        L = int(rng.integers(1, 600))
        tpl = rng.integers(0, 4, size=L, dtype=np.uint8)
        err = float(rng.choice([0.0, 0.02, 0.1, 0.3]))
        a = workloads._mutate(rng, tpl, err / 3, err / 3, err / 3)
        b = workloads._mutate(rng, tpl, err / 3, err / 3, err / 3)
        if rng.random() < 0.1:
            b = rng.integers(0, 4, size=int(rng.integers(1, 400)), dtype=np.uint8)
        if a.size == 0 or b.size == 0:
            continue
        x, y = workloads._to_str(a).encode(), workloads._to_str(b).encode()
        d = O.ed_plain(x, y)
        for k in {0, 1, max(0, d - 1), d, d + 1, d + 13, d + 64, max(len(x), len(y))}:
            if abs(len(y) - len(x)) > k:
                continue
            want = d if d <= k else -1
            # own strip, and a strip widened like the union over a warp's lanes / a rounded-up W
            for wlo, whi, fw in ((0, 0, 0), (int(rng.integers(0, 40)), int(rng.integers(0, 40)), int(rng.integers(0, 12)))):
                assert sim.sim_ed(x, len(x), y, len(y), k, wlo, whi, fw) == want, (len(x), len(y), d, k, wlo, whi, fw)
                pad = int(rng.integers(0, 3))
                assert sim.sim_ed_diag(x, len(x), y, len(y), k, wlo, whi, fw, pad) == want, (
                    "diag", len(x), len(y), d, k, wlo, whi, fw, pad)
                checked += 1
    assert checked > 10000


def test_band_arithmetic_on_real_reads(sim):
    S = util.load_reads(200)
    seqs = list(S.values())
    rng = np.random.default_rng(9)
    for _ in range(120):
        i, j = rng.integers(0, len(seqs), size=2)
        x, y = seqs[i].encode(), seqs[j].encode()
        d = O.ed_myers64(x, y, -1)
        for k in (d - 1, d, d + 25):
            if k < 0 or abs(len(y) - len(x)) > k or k > 1300:   # the host harness instantiates W <= 48
                continue
            assert sim.sim_ed(x, len(x), y, len(y), k, 0, 0, 0) == (d if d <= k else -1)
            assert sim.sim_ed_diag(x, len(x), y, len(y), k, 0, 0, 0, 0) == (d if d <= k else -1)


def test_shrinking_window_is_exact(sim):
    """ed_group_diag_run (the row kernel's walk): dropping dead window cells never changes the result, for any
    check interval, any initial window and any widening of the alive interval by other lanes; and it does
    shrink the work on divergent pairs."""
    rng = np.random.default_rng(11)
    out = (ctypes.c_int * 4)()
    checked = narrowed = 0
    saved = []
    for it in range(700):
        L = int(rng.integers(40, 1400))
        tpl = rng.integers(0, 4, size=L, dtype=np.uint8)
        err = float(rng.choice([0.0, 0.01, 0.03, 0.06, 0.12]))
        a = workloads._mutate(rng, tpl, err * 0.5, err * 0.3, err * 0.2)
        b = workloads._mutate(rng, tpl, err * 0.5, err * 0.3, err * 0.2)
        if rng.random() < 0.05:
            b = rng.integers(0, 4, size=max(1, L + int(rng.integers(-30, 30))), dtype=np.uint8)
        if a.size == 0 or b.size == 0:
            continue
        x, y = workloads._to_str(a).encode(), workloads._to_str(b).encode()
        d = O.ed_myers64(x, y, -1)
        for k in {max(0, d - 40), max(0, d - 7), max(0, d - 1), d, d + 1, d + 9, d + 50, int(rng.integers(0, 400))}:
            if abs(len(y) - len(x)) > k or k > 400:
                continue
            want = d if d <= k else -1
            for narrow in (1, 2, 4, int(rng.integers(1, 9))):
                wlo, whi, fw = (0, 0, 0) if rng.random() < 0.5 else (int(rng.integers(0, 40)), int(rng.integers(0, 40)), int(rng.integers(0, 12)))
                uwlo, uwhi = (0, 0) if rng.random() < 0.5 else (int(rng.integers(0, 50)), int(rng.integers(0, 50)))
                got = sim.sim_ed_diag_run(x, len(x), y, len(y), k, wlo, whi, fw, int(rng.integers(0, 3)), narrow, uwlo, uwhi, out)
                if got == -99:
                    continue
                assert got == want, (len(x), len(y), d, k, narrow, wlo, whi, fw, uwlo, uwhi)
                checked += 1
                if out[0] < out[3] * out[1]:
                    narrowed += 1
                if narrow == 2 and (uwlo, uwhi, wlo, whi, fw) == (0, 0, 0, 0, 0) and out[1] >= 512 and out[3] >= 3:
                    saved.append(out[0] / float(out[3] * out[1]))
            # no shrinking at all: same answer, full width everywhere
            got = sim.sim_ed_diag_run(x, len(x), y, len(y), k, 0, 0, 0, 0, 0, 0, 0, out)
            if got != -99:
                assert got == want and out[0] == out[3] * out[1]
    assert checked > 5000 and narrowed > 1000
    assert saved and float(np.mean(saved)) < 0.85, (len(saved), float(np.mean(saved)))


@pytest.mark.timeout(600)
def test_shrinking_window_whole_warp(sim):
    """32 host threads in lock step (the warp reductions go through a barrier) walk one query against 32 targets
    the way nn_row_kernel does: per-lane thresholds and windows, union of the alive intervals, warp-uniform
    shifts.  Every lane must return edlib's answer for ITS threshold, with and without shrinking."""
    rng = np.random.default_rng(21)
    total = full = runs = 0
    for it in range(60):
        L = int(rng.integers(200, 1600))
        root = rng.integers(0, 4, size=L, dtype=np.uint8)
        err = float(rng.choice([0.01, 0.03, 0.05, 0.1]))
        mk = lambda: workloads._mutate(rng, workloads._mutate(rng, root, 0.002, 0.001, 0.001) if rng.random() < 0.5 else root,
                                       err * 0.5, err * 0.3, err * 0.2)
        q = workloads._to_str(mk()).encode()
        ts = [workloads._to_str(mk()).encode() for _ in range(32)]
        if it % 7 == 0:
            ts[3] = workloads._to_str(rng.integers(0, 4, size=L, dtype=np.uint8)).encode()   # unrelated read
            ts[9] = q                                                                          # distance 0
        ts.sort(key=len)
        ds = [O.ed_myers64(q, t, -1) for t in ts]
        base = int(np.percentile(ds, 20))
        ks = []
        for l in range(32):
            u = rng.random()
            ks.append(-1 if u < 0.06 else min(400, max(0, base + int(rng.integers(-25, 25)))))
        if it % 5 == 0:
            ks = [min(400, max(0, d + int(rng.integers(-2, 3)))) for d in ds]              # thresholds right at the answers
        toff = np.zeros(33, dtype=np.int32)
        toff[1:] = np.cumsum([len(t) for t in ts])
        tcat = b"".join(ts)
        karr = (ctypes.c_int * 32)(*ks)
        res = (ctypes.c_int * 32)()
        out = (ctypes.c_int * 3)()
        for narrow in (0, 1, 2, 4):
            rc = sim.sim_warp_diag_run(q, len(q), tcat, toff.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), karr, narrow, res, out)
            if rc in (-1, -99):
                continue
            assert rc == 0
            for l in range(32):
                k = ks[l]
                want = -1 if (k < 0 or abs(len(ts[l]) - len(q)) > k or ds[l] > k) else ds[l]
                assert res[l] == want, (it, narrow, l, k, ds[l], res[l])
            if narrow == 0:
                assert out[0] == out[1] * out[2]
            if narrow == 2:
                total += out[0]; full += out[1] * out[2]; runs += 1
    assert runs >= 40 and full > 0 and total < 0.9 * full, (runs, total, full)


@pytest.mark.timeout(900)
def test_shrinking_window_whole_warp_wide_windows(sim):
    """The same lock-step harness in the regime of the PILOT pass and of noisy reads: windows of 6-14 words (alive
    bounds at word granularity there), thresholds far above the distances, mixed with lanes that are pruned."""
    rng = np.random.default_rng(33)
    seen = set()
    total = full = 0
    for it in range(36):
        L = int(rng.integers(500, 1500))
        root = rng.integers(0, 4, size=L, dtype=np.uint8)
        err = float(rng.choice([0.04, 0.08, 0.14, 0.2]))
        mk = lambda: workloads._mutate(rng, root, err * 0.4, err * 0.35, err * 0.25)
        q = workloads._to_str(mk()).encode()
        ts = [workloads._to_str(mk()).encode() for _ in range(32)]
        if it % 6 == 0:
            ts[5] = workloads._to_str(rng.integers(0, 4, size=L, dtype=np.uint8)).encode()
        ts.sort(key=len)
        ds = [O.ed_myers64(q, t, -1) for t in ts]
        mode = it % 3
        if mode == 0:       # pilot: nothing known yet, every pair at the cap
            ks = [400] * 32
        elif mode == 1:     # thresholds a little above / below the answers, up to the cap
            ks = [min(400, max(0, d + int(rng.integers(-30, 60)))) for d in ds]
        else:               # one loose lane keeps the window wide for everybody
            ks = [min(400, max(0, d - int(rng.integers(0, 20)))) for d in ds]
            ks[int(rng.integers(0, 32))] = 400
        toff = np.zeros(33, dtype=np.int32)
        toff[1:] = np.cumsum([len(t) for t in ts])
        tcat = b"".join(ts)
        karr = (ctypes.c_int * 32)(*ks)
        res = (ctypes.c_int * 32)()
        out = (ctypes.c_int * 3)()
        for narrow in (0, 1, 3):
            rc = sim.sim_warp_diag_run(q, len(q), tcat, toff.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), karr, narrow, res, out)
            if rc in (-1, -99):
                continue
            assert rc == 0
            seen.add(out[2])
            for l in range(32):
                k = ks[l]
                want = -1 if (abs(len(ts[l]) - len(q)) > k or ds[l] > k) else ds[l]
                assert res[l] == want, (it, narrow, l, k, ds[l], res[l])
            if narrow == 3:
                total += out[0]; full += out[1] * out[2]
    assert max(seen) >= 12 and len([w for w in seen if w >= 6]) >= 4, sorted(seen)
    assert total < full, (total, full)
