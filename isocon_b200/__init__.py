"""isocon_b200 -- B200-native nearest-neighbour graph for IsoCon.

Only the hot path of ``modules/nearest_neighbor_graph.py`` lives here; the
rest of IsoCon runs unchanged and reaches this package through ``install()``.
"""
import sys

__all__ = ["install", "nearest_neighbor_graph", "edlib_alignment_module"]


def install(package="modules", pair_distances=True):
    """Shadow ``<package>.nearest_neighbor_graph`` (and, with ``pair_distances``,
    ``<package>.edlib_alignment_module``) with the device implementations.

    Call before ``modules.graphs`` is imported (it does ``from modules import
    nearest_neighbor_graph``, graphs.py:17) -- or afterwards: the attributes on already
    imported consumers are patched too (``graphs``; ``isocon_get_candidates`` and
    ``isocon_statistical_test`` bind the two ``edlib_align_sequences*`` functions by name,
    isocon_get_candidates.py:15, isocon_statistical_test.py:25).  Returns the replacement
    nearest_neighbor_graph module.
    """
    import importlib
    from . import nearest_neighbor_graph as replacement
    shadows = {"nearest_neighbor_graph": replacement}
    if pair_distances:
        from . import edlib_alignment_module as pairs
        shadows["edlib_alignment_module"] = pairs
    try:
        pkg = importlib.import_module(package)
    except ImportError:
        pkg = None
    for name, mod in shadows.items():
        sys.modules[package + "." + name] = mod
        if pkg is not None:
            setattr(pkg, name, mod)
    graphs = sys.modules.get(package + ".graphs")
    if graphs is not None:
        graphs.nearest_neighbor_graph = replacement
    if pair_distances:
        for consumer in ("isocon_get_candidates", "isocon_statistical_test"):
            m = sys.modules.get(package + "." + consumer)
            if m is None:
                continue
            for fn in ("edlib_align_sequences", "edlib_align_sequences_keeping_accession"):
                if hasattr(m, fn):
                    setattr(m, fn, getattr(shadows["edlib_alignment_module"], fn))
    return replacement
