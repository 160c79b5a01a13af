"""ctypes.PyDLL binding of ``libisocon_hostops.so`` (``csrc/hostops.cpp``): the list-walking loops of the
reference-facing functions (lengths, content-keyed slot lookup, gathering strings into the pinned upload buffer,
rebuilding the dict-of-dicts result) at memcpy speed.  Host logic only -- no arithmetic of the path lives here."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libisocon_hostops.so")
EXPORTS = ["iso_host_lengths", "iso_host_lookup", "iso_host_register", "iso_host_gather", "iso_host_prepare_graph",
           "iso_host_fill_graph", "iso_host_permute", "iso_host_contains"]
_LIB = None


def load_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `make -C isocon_b200/csrc`" % LIB_PATH)
    L = ctypes.PyDLL(LIB_PATH)
    po, vp, ll = ctypes.py_object, ctypes.c_void_p, ctypes.c_longlong
    L.iso_host_lengths.argtypes = [po, vp]
    L.iso_host_lengths.restype = ll
    L.iso_host_lookup.argtypes = [po, po, vp]
    L.iso_host_lookup.restype = ll
    L.iso_host_register.argtypes = [po, po, vp, ll, ll]
    L.iso_host_register.restype = ctypes.c_int
    L.iso_host_gather.argtypes = [po, vp, ll, vp, ll, vp]
    L.iso_host_gather.restype = ll
    L.iso_host_permute.argtypes = [po, vp, ll]
    L.iso_host_permute.restype = po
    L.iso_host_contains.argtypes = [po, po, vp]
    L.iso_host_contains.restype = ll
    L.iso_host_prepare_graph.argtypes = [po, ll, ll, vp, vp]
    L.iso_host_prepare_graph.restype = po
    L.iso_host_fill_graph.argtypes = [po, po, ll, ll, vp, vp, vp, ll, vp]
    L.iso_host_fill_graph.restype = ctypes.c_int
    _LIB = L
    return L


def lengths(seqs):
    """int64 array of len(s) for every str of the list."""
    L = load_library()
    out = np.empty(max(len(seqs), 1), np.int64)
    L.iso_host_lengths(seqs, out.ctypes.data)
    return out[:len(seqs)]


def permute(items, order):
    """[items[i] for i in order] as a new list."""
    L = load_library()
    order = np.ascontiguousarray(order, dtype=np.int64)
    return L.iso_host_permute(items, order.ctypes.data, order.size)


def contains(container, keys):
    """uint8 mask: keys[i] in container."""
    L = load_library()
    out = np.empty(max(len(keys), 1), np.uint8)
    L.iso_host_contains(container, keys, out.ctypes.data)
    return out[:len(keys)]


def lookup(store, seqs):
    """(int32 slots, number missing): slots[i] = store[seqs[i]] or -1."""
    L = load_library()
    slots = np.empty(max(len(seqs), 1), np.int32)
    missing = L.iso_host_lookup(store, seqs, slots.ctypes.data)
    return slots[:len(seqs)], int(missing)


def register(store, seqs, sel, first_slot):
    L = load_library()
    sel = np.ascontiguousarray(sel, dtype=np.int32)
    L.iso_host_register(store, seqs, sel.ctypes.data, sel.size, int(first_slot))


def gather(seqs, sel, dst_ptr, cap):
    """Concatenate the ASCII bytes of seqs[sel] at dst_ptr; returns (total, int64 offsets[len(sel) + 1])."""
    L = load_library()
    sel = np.ascontiguousarray(sel, dtype=np.int32)
    off = np.empty(sel.size + 1, np.int64)
    total = L.iso_host_gather(seqs, sel.ctypes.data, sel.size, dst_ptr, int(cap), off.ctypes.data)
    return int(total), off


def prepare_graph(accs, lo, hi, skip):
    """({accs[i]: {}} for the entries of [lo, hi) that are not skipped, in list order; the addresses of those dicts,
    for fill_graph).  Needs no device result."""
    L = load_library()
    sk = None if skip is None else np.ascontiguousarray(skip, dtype=np.uint8)
    dicts = np.empty(max(int(hi) - int(lo), 1), np.int64)
    out = L.iso_host_prepare_graph(accs, int(lo), int(hi), None if sk is None else sk.ctypes.data, dicts.ctypes.data)
    return out, dicts


def fill_graph(prepared, accs, lo, hi, eq, et, ed):
    """Insert the device's unordered edges into the prepared dicts in the reference's scan order; returns the graph."""
    L = load_library()
    out, dicts = prepared
    eq = np.ascontiguousarray(eq, dtype=np.int32); et = np.ascontiguousarray(et, dtype=np.int32)
    ed = np.ascontiguousarray(ed, dtype=np.int32)
    L.iso_host_fill_graph(out, accs, int(lo), int(hi), eq.ctypes.data, et.ctypes.data, ed.ctypes.data, int(eq.size),
                          dicts.ctypes.data)
    return out


def build_graph(accs, lo, hi, skip, eq, et, ed):
    """The reference's dict-of-dicts from unordered device edges (key order = list order, insertion = scan order)."""
    return fill_graph(prepare_graph(accs, lo, hi, skip), accs, lo, hi, eq, et, ed)
