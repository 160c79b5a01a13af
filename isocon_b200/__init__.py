"""isocon_b200 -- B200-native nearest-neighbour graph for IsoCon.

Only the hot path of ``modules/nearest_neighbor_graph.py`` lives here; the
rest of IsoCon runs unchanged and reaches this package through ``install()``.
"""
import sys

__all__ = ["install", "nearest_neighbor_graph"]


def install(package="modules"):
    """Shadow ``<package>.nearest_neighbor_graph`` with the device implementation.

    Call before ``modules.graphs`` is imported (it does ``from modules import
    nearest_neighbor_graph``, graphs.py:17) -- or afterwards: the attribute on an already
    imported ``modules.graphs`` is patched too.  Returns the replacement module.
    """
    import importlib
    from . import nearest_neighbor_graph as replacement
    name = package + ".nearest_neighbor_graph"
    sys.modules[name] = replacement
    try:
        pkg = importlib.import_module(package)
        setattr(pkg, "nearest_neighbor_graph", replacement)
    except ImportError:
        pass
    graphs = sys.modules.get(package + ".graphs")
    if graphs is not None:
        graphs.nearest_neighbor_graph = replacement
    return replacement
