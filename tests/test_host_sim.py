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
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-o", so, src])
    L = ctypes.CDLL(so)
    L.sim_ed.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int] + [ctypes.c_int] * 4
    L.sim_ed_diag.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int] + [ctypes.c_int] * 5
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
