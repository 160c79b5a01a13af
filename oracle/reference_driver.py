"""oracle/reference_driver.py -- TEST INFRASTRUCTURE, authoring container only.

Imports the UNMODIFIED reference module ``/root/reference/modules/nearest_neighbor_graph.py``
with the ``edlib`` stand-in of ``oracle/edlib_shim.cpp`` first on ``sys.path`` (the real
edlib, parasail and pysam are not installed here).  Used by
``oracle/make_golden.py`` to produce the fixtures in ``tests/golden/`` and by the
container-only tests that pin the C++ oracle against the reference driver.

``/root/reference`` does not exist on the GPU box: nothing that runs there imports this.
"""
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ISOCON_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_MOD = None


def available():
    shim = os.path.join(_HERE, "shim")
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "modules")) or not os.path.isdir(shim):
        return False
    return any(f.startswith("edlib") and f.endswith(".so") for f in os.listdir(shim))


def load():
    """Return (reference nearest_neighbor_graph module, edlib shim module)."""
    global _MOD
    if _MOD is not None:
        return _MOD
    if not available():
        raise RuntimeError("reference tree or edlib shim missing (authoring container only)")
    shim_dir = os.path.join(_HERE, "shim")
    sys.path.insert(0, shim_dir)
    import edlib  # noqa: the stand-in
    assert getattr(edlib, "__shim__", None), "a real edlib shadowed the shim?"
    for name in ("parasail", "pysam"):  # imported by write_output -> SW_alignment_module; unused here
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)
    from modules import nearest_neighbor_graph as ref_nn
    assert ref_nn.__file__.startswith(REFERENCE_ROOT)
    _MOD = (ref_nn, edlib)
    return _MOD


class Params(object):
    """Attribute bag like modules/isocon_parameters.py:10-19 with the four fields the path reads."""

    def __init__(self, nr_cores=1, neighbor_search_depth=2 ** 32, verbose=False, develop_logfile=None):
        self.nr_cores = nr_cores
        self.neighbor_search_depth = neighbor_search_depth
        self.verbose = verbose
        self.develop_logfile = develop_logfile


@contextlib.contextmanager
def quiet():
    """The reference prints progress lines (:116-117, :273, :345-346); they are not API."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
