"""CPU suite, part 3: the C-ABI library loads and exports every symbol include/isocon_nn.h
declares; without a GPU every entry point fails loudly (no CPU fallback)."""
import os
import re

import pytest

import util
from isocon_b200 import _binding


def declared_functions():
    text = open(os.path.join(util.ROOT, "include", "isocon_nn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(isocon_nn_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(_binding.EXPORTS)


def test_library_exports_every_declared_symbol():
    L = _binding.load_library()
    for name in declared_functions():
        assert hasattr(L, name), name


def test_product_never_imports_the_oracle():
    pkg = os.path.join(util.ROOT, "isocon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace(
                    "oracle/make_golden.py", ""), f


def test_no_gpu_means_loud_failure():
    try:
        n = _binding.device_count()
    except _binding.IsoconNNError:
        n = 0
    if n > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_binding.IsoconNNError):
        _binding.NNContext(0)
    from isocon_b200 import nearest_neighbor_graph as nn
    with pytest.raises(_binding.IsoconNNError):
        nn.compute_nearest_neighbor_graph({"a": "ACGT", "b": "ACGA"}, set(), util.Params())


def test_edge_ordering_is_the_scan_order():
    """Host logic: device edges come unordered (and possibly twice); the reference's dict insertion order is per
    query by offset |t - q|, down before up (nearest_neighbor_graph.py:131-188)."""
    import numpy as np
    from isocon_b200 import nearest_neighbor_graph as nn
    rng = np.random.default_rng(3)
    n = 5000
    eq = rng.integers(0, n, size=20000).astype(np.int32)
    et = rng.integers(0, n, size=20000).astype(np.int32)
    keep = eq != et
    eq, et = eq[keep], et[keep]
    ed = ((eq.astype(np.int64) * 31 + et) % 200).astype(np.int32)          # a function of the pair
    eq = np.concatenate([eq, eq[:500]]); et = np.concatenate([et, et[:500]]); ed = np.concatenate([ed, ed[:500]])
    q, t, d = nn._order_edges(eq, et, ed)
    want = sorted(set(zip(eq.tolist(), et.tolist(), ed.tolist())), key=lambda e: (e[0], abs(e[1] - e[0]), e[1] > e[0]))
    assert list(zip(q.tolist(), t.tolist(), d.tolist())) == want
    z = np.zeros(0, np.int32)
    assert all(a.size == 0 for a in nn._order_edges(z, z, z))
