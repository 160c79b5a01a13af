"""CPU suite: the host side of ``nearest_neighbor_graph._build_graph`` around the device call -- the helper thread that
makes the result dicts while the device works, the hand-over of the dicts' addresses to ``fill_graph``, error paths --
with a stand-in for the device (no kernel, no arithmetic of the path)."""
import threading

import numpy as np
import pytest

import util
from isocon_b200 import _hostops, sharding
from isocon_b200 import nearest_neighbor_graph as nn


class _FakeContext(object):
    """What _build_graph touches of NNContext: use_list and the on_run hook."""
    def __init__(self):
        self.on_run = None
        self.lists = 0

    def use_list(self, seqs, lens=None):
        self.lists += 1


def _edges(n, rng):
    eq = rng.integers(0, n, size=3 * n).astype(np.int32)
    et = rng.integers(0, n, size=3 * n).astype(np.int32)
    keep = eq != et
    eq, et = eq[keep], et[keep]
    return eq, et, ((eq.astype(np.int64) * 7 + et) % 90).astype(np.int32)


@pytest.mark.parametrize("fire_hook", [True, False])
def test_helper_thread_builds_the_dicts_beside_the_device_call(monkeypatch, fire_hook):
    n = 5000                                            # >= 2000: the threaded path
    rng = np.random.default_rng(1)
    accs = ["r%d" % i for i in range(n)]
    seqs = ["A" * (10 + i // 100) for i in range(n)]
    eq, et, ed = _edges(n, rng)
    ctx = _FakeContext()
    seen = {}

    def fake_device_graph(c, mode, depth, is_query, is_target, *a, **kw):
        assert c is ctx and callable(c.on_run)
        if fire_hook:                                   # what NNContext.graph_run does right before the library call
            hook, c.on_run = c.on_run, None
            hook()
        seen["threads"] = threading.active_count()
        return np.zeros(n, np.int32), eq, et, ed

    monkeypatch.setattr(nn, "_ctx", lambda: ctx)
    monkeypatch.setattr(sharding, "device_graph", fake_device_graph)
    got = nn._build_graph(seqs, accs, None, 1, np.ones(n, np.uint8), None, 2 ** 32, 0, n)
    want = _hostops.build_graph(accs, 0, n, None, eq, et, ed)
    util.assert_same_graph(got, want)
    assert ctx.on_run is None and ctx.lists == 1 and seen["threads"] >= 2
    # 2-set: targets are no keys
    ist = (rng.random(n) < 0.1).astype(np.uint8)
    m = ist[eq] == 0
    eq, et, ed = eq[m], et[m], ed[m]
    got = nn._build_graph(seqs, accs, None, 2, 1 - ist, ist, 2 ** 32, 0, n)
    util.assert_same_graph(got, _hostops.build_graph(accs, 0, n, ist, eq, et, ed))
    assert not any(accs[i] in got for i in np.flatnonzero(ist)[:50])


def test_device_error_reaches_the_caller_and_leaves_no_thread_behind(monkeypatch):
    n = 3000
    ctx = _FakeContext()

    def failing(c, *a, **kw):
        raise RuntimeError("device lost")

    monkeypatch.setattr(nn, "_ctx", lambda: ctx)
    monkeypatch.setattr(sharding, "device_graph", failing)
    before = threading.active_count()
    with pytest.raises(RuntimeError, match="device lost"):
        nn._build_graph(["A"] * n, ["r%d" % i for i in range(n)], None, 1, np.ones(n, np.uint8), None, 2 ** 32, 0, n)
    assert ctx.on_run is None and threading.active_count() == before


def test_helper_error_is_raised_on_the_calling_thread(monkeypatch):
    n = 3000
    ctx = _FakeContext()
    monkeypatch.setattr(nn, "_ctx", lambda: ctx)
    z = np.zeros(0, np.int32)
    monkeypatch.setattr(sharding, "device_graph", lambda c, *a, **kw: (None, z, z, z))
    with pytest.raises(ValueError):                     # key range outside the list: raised inside the helper
        nn._build_graph(["A"] * n, ["r%d" % i for i in range(n)], None, 1, np.ones(n, np.uint8), None, 2 ** 32, 0, n + 5)
