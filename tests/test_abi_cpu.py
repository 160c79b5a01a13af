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
