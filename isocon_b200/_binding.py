"""ctypes binding of ``libisocon_nn.so`` (C ABI: ``include/isocon_nn.h``).

The library is the product: there is no Python or CPU fallback.  Loading fails loudly when
the shared object is missing, and every call fails loudly when no CUDA device is present.
"""
import ctypes
import os

import numpy as np

from . import _hostops

_HERE = os.path.dirname(os.path.abspath(__file__))
# ISOCON_NN_LIB: another build of the SAME library (kernel tuning experiments, tools/ab_*.sh); never a fallback
LIB_PATH = os.environ.get("ISOCON_NN_LIB") or os.path.join(_HERE, "libisocon_nn.so")

ALGO_AUTO, ALGO_TILE, ALGO_SCAN = 0, 1, 2
PHASE_SEED, PHASE_MAIN, PHASE_WIDE, PHASE_PILOT, PHASE_ALL = 1, 2, 4, 8, 15

EXPORTS = [
    "isocon_nn_device_count", "isocon_nn_create", "isocon_nn_destroy", "isocon_nn_last_error",
    "isocon_nn_set_reads", "isocon_nn_graph_begin", "isocon_nn_graph_run", "isocon_nn_best_dev",
    "isocon_nn_graph_finalize", "isocon_nn_graph_fetch", "isocon_nn_edges_dev", "isocon_nn_ed_pairs",
    "isocon_nn_get_stats", "isocon_nn_last_ms", "isocon_nn_sync", "isocon_nn_int32_peak",
    "isocon_nn_timer_start", "isocon_nn_timer_stop", "isocon_nn_ipc_handles", "isocon_nn_set_peers",
    "isocon_nn_release_retired", "isocon_nn_last_run_rows",
    "isocon_nn_store_reset", "isocon_nn_store_add", "isocon_nn_set_list", "isocon_nn_host_buffer",
    "isocon_nn_store_info", "isocon_nn_reserve_edges", "isocon_nn_pilot_near_dev",
    "isocon_nn_best_agree", "isocon_nn_can_fuse",
]
IPC_BYTES = 256
ERR_ALPHABET, ERR_OVERFLOW = 3, 5


class IsoconNNError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "libisocon_nn error %d: %s" % (code, message))
        self.code = code


class _Params(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int32), ("algo", ctypes.c_int32), ("depth", ctypes.c_int64),
                ("is_query", ctypes.c_void_p), ("is_target", ctypes.c_void_p),
                ("symmetric", ctypes.c_int32), ("rank", ctypes.c_int32), ("world", ctypes.c_int32)]


class _Stats(ctypes.Structure):
    _fields_ = [(name, ctypes.c_uint64) for name in
                ("pairs", "word_columns", "groups", "wide_pairs", "items", "edges_raw", "launches", "bins",
                 "pilot_rows", "unresolved_rows", "useful_cells", "columns", "main_passes", "clusters")]


class _StoreStats(ctypes.Structure):
    _fields_ = [(name, ctypes.c_uint64) for name in
                ("slots", "arena_words", "list_entries", "uploaded_reads", "uploaded_bytes", "upload_calls",
                 "resets", "lists", "foreign_reads")]


_LIB = None


def load_library():
    """dlopen the CUDA library; raises if it has not been built (``__graft_entry__.build()``)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `make -C isocon_b200/csrc` (needs nvcc, sm_100a). "
                          "isocon_b200 has no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    L.isocon_nn_device_count.argtypes = [ctypes.POINTER(ctypes.c_int)]
    L.isocon_nn_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    L.isocon_nn_destroy.argtypes = [vp]
    L.isocon_nn_destroy.restype = None
    L.isocon_nn_last_error.argtypes = [vp]
    L.isocon_nn_last_error.restype = ctypes.c_char_p
    L.isocon_nn_set_reads.argtypes = [vp, vp, vp, i64]
    L.isocon_nn_graph_begin.argtypes = [vp, ctypes.POINTER(_Params)]
    L.isocon_nn_graph_run.argtypes = [vp, ctypes.c_int]
    L.isocon_nn_best_dev.argtypes = [vp, ctypes.POINTER(vp)]
    L.isocon_nn_graph_finalize.argtypes = [vp, ctypes.POINTER(i64)]
    L.isocon_nn_graph_fetch.argtypes = [vp, vp, vp, vp, vp]
    L.isocon_nn_edges_dev.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp)]
    L.isocon_nn_ed_pairs.argtypes = [vp, vp, vp, vp, i64, vp]
    L.isocon_nn_get_stats.argtypes = [vp, ctypes.POINTER(_Stats)]
    L.isocon_nn_last_ms.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
    L.isocon_nn_sync.argtypes = [vp]
    L.isocon_nn_timer_start.argtypes = [vp]
    L.isocon_nn_timer_stop.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    L.isocon_nn_int32_peak.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    L.isocon_nn_ipc_handles.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_uint64)]
    L.isocon_nn_set_peers.argtypes = [vp, vp, i32, i32]
    L.isocon_nn_release_retired.argtypes = [vp]
    L.isocon_nn_last_run_rows.argtypes = [vp, ctypes.POINTER(i64)]
    L.isocon_nn_store_reset.argtypes = [vp, vp]
    L.isocon_nn_store_add.argtypes = [vp, vp, vp, i64, ctypes.POINTER(i64)]
    L.isocon_nn_set_list.argtypes = [vp, vp, i64]
    L.isocon_nn_host_buffer.argtypes = [vp, i64, ctypes.POINTER(vp)]
    L.isocon_nn_store_info.argtypes = [vp, ctypes.POINTER(_StoreStats)]
    L.isocon_nn_reserve_edges.argtypes = [vp, i64]
    L.isocon_nn_pilot_near_dev.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(i64)]
    L.isocon_nn_best_agree.argtypes = [vp]
    L.isocon_nn_can_fuse.argtypes = [vp, ctypes.POINTER(i32)]
    _LIB = L
    return L


def device_count():
    L = load_library()
    c = ctypes.c_int(0)
    rc = L.isocon_nn_device_count(ctypes.byref(c))
    if rc:
        raise IsoconNNError(rc, L.isocon_nn_last_error(None).decode())
    return c.value


class _DevArray(object):
    """Zero-copy view of a device buffer for ``torch.as_tensor`` (CUDA array interface v2)."""

    def __init__(self, ptr, n, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class NNContext(object):
    """One device context: resident packed reads + graph state (isocon_nn_ctx)."""

    def __init__(self, device=0):
        self._L = load_library()
        h = ctypes.c_void_p()
        rc = self._L.isocon_nn_create(int(device), ctypes.byref(h))
        if rc:
            raise IsoconNNError(rc, self._L.isocon_nn_last_error(None).decode())
        self._h = h
        self.on_run = None
        self.device = int(device)
        self.n = 0
        self._slot_of = None
        self._list_slots = None
        self._alphabet = b"ACGT"
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.isocon_nn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise IsoconNNError(rc, self._L.isocon_nn_last_error(self._h).decode())

    # ------------------------------------------------------------------ reads
    # The resident read store: a sequence is uploaded once and found again BY CONTENT (a dict str -> slot; Python
    # caches a str's hash, so a lookup costs a pointer compare for the same object and one memcmp otherwise).  A
    # graph over mostly the same sequences -- the next correction round, isocon_get_candidates.py:141-214 -- uploads
    # only what the correction changed.
    STORE_SLACK = 3          # reset the store when it holds more than STORE_SLACK x the list's entries (+ 4096)

    def store_reset(self, alphabet=b"ACGT"):
        abc = np.frombuffer(bytes(alphabet), dtype=np.uint8).copy()
        assert abc.size == 4
        self._check(self._L.isocon_nn_store_reset(self._h, abc.ctypes.data))
        self._slot_of = {}
        self._list_slots = None
        self._alphabet = bytes(alphabet)
        self.n = 0

    @staticmethod
    def _pick_alphabet(seqs):
        """The four symbols the store packs in 2 bits: ACGT unless the first reads say otherwise (soft-masked
        lower-case input, RNA): the most frequent symbols of a sample, padded from ACGT."""
        sample = "".join(seqs[:64])[:1 << 16]
        if not sample or not sample.strip("ACGT"):
            return b"ACGT"
        if not sample.isascii():
            raise ValueError("reads must be ASCII strings")
        counts = np.bincount(np.frombuffer(sample.encode(), dtype=np.uint8), minlength=256)
        top = [int(c) for c in np.argsort(-counts, kind="stable")[:4] if counts[c] > 0]
        for c in b"ACGTacgtNn":
            if len(top) < 4 and c not in top:
                top.append(c)
        return bytes(sorted(top))

    def use_list(self, seqs, lens=None):
        """Make ``seqs`` (list of str sorted by length) the list the next graphs / ed_pairs calls index into.
        Sequences already resident are found by content; only the others are uploaded.  Returns how many were."""
        n = len(seqs)
        if getattr(self, "_slot_of", None) is None:
            self.store_reset(self._pick_alphabet(seqs))
        if len(self._slot_of) > self.STORE_SLACK * n + 4096:
            self.store_reset(self._alphabet)                   # mostly dead sequences of earlier rounds: start over
        slots, missing = _hostops.lookup(self._slot_of, seqs)
        if missing:
            if lens is None:
                lens = _hostops.lengths(seqs)
            sel = np.flatnonzero(slots < 0).astype(np.int32)
            total = int(lens[sel].sum())
            buf = ctypes.c_void_p()
            self._check(self._L.isocon_nn_host_buffer(self._h, total, ctypes.byref(buf)))
            got, off = _hostops.gather(seqs, sel, buf.value, total)
            assert got == total
            first = ctypes.c_int64(0)
            self._check(self._L.isocon_nn_store_add(self._h, buf.value, off.ctypes.data, sel.size, ctypes.byref(first)))
            _hostops.register(self._slot_of, seqs, sel, first.value)
            slots[sel] = first.value + np.arange(sel.size, dtype=np.int32)
        if self._list_slots is None or self._list_slots.size != n or not np.array_equal(self._list_slots, slots):
            self._check(self._L.isocon_nn_set_list(self._h, slots.ctypes.data if n else None, n))
            self._list_slots = slots.copy()
        self.n = n
        return missing

    def set_reads(self, seqs, key=None):
        """Upload the length-sorted list of sequences (str) from scratch (store reset + whole upload)."""
        self.store_reset(self._pick_alphabet(seqs))
        self.use_list(list(seqs))
        return True

    def store_info(self):
        s = _StoreStats()
        self._check(self._L.isocon_nn_store_info(self._h, ctypes.byref(s)))
        return {name: int(getattr(s, name)) for name, _ in _StoreStats._fields_}

    def reserve_edges(self, capacity):
        self._check(self._L.isocon_nn_reserve_edges(self._h, int(capacity)))

    # ------------------------------------------------------------------ graph
    def graph_begin(self, mode, depth, is_query, is_target=None, algo=ALGO_AUTO, symmetric=True, rank=0, world=1):
        isq = np.ascontiguousarray(is_query, dtype=np.uint8)
        ist = None if is_target is None else np.ascontiguousarray(is_target, dtype=np.uint8)
        assert isq.size == self.n and (ist is None or ist.size == self.n)
        if isq.size == 0:
            isq = np.zeros(1, np.uint8)
        p = _Params(mode=mode, algo=algo, depth=int(min(max(int(depth), 0), 2 ** 62)),   # <= 0: see graph_begin
                    is_query=isq.ctypes.data, is_target=None if ist is None else ist.ctypes.data,
                    symmetric=1 if symmetric else 0, rank=rank, world=world)
        self._keep = (isq, ist)
        self._check(self._L.isocon_nn_graph_begin(self._h, ctypes.byref(p)))

    def graph_run(self, phases=PHASE_ALL):
        # on_run: called once, right before the library call that releases the GIL for the length of the device work
        # (nearest_neighbor_graph._build_graph lets its helper thread go here, so the helper's GIL-bound work falls
        # into the time this thread spends inside the library and not in front of it)
        hook, self.on_run = self.on_run, None
        if hook is not None:
            hook()
        self._check(self._L.isocon_nn_graph_run(self._h, int(phases)))

    def last_run_rows(self):
        """Rows (queries) the last graph_run scheduled over all ranks together; the same on every rank."""
        v = ctypes.c_int64(0)
        self._check(self._L.isocon_nn_last_run_rows(self._h, ctypes.byref(v)))
        return int(v.value)

    def best_dev(self):
        p = ctypes.c_void_p()
        self._check(self._L.isocon_nn_best_dev(self._h, ctypes.byref(p)))
        return _DevArray(p.value, self.n)

    def best_agree(self):
        self._check(self._L.isocon_nn_best_agree(self._h))

    def pilot_near_dev(self):
        """Device view (int64[2n]) of the nearest-pilot-row records of the PILOT phase, or None."""
        p, c = ctypes.c_void_p(), ctypes.c_int64(0)
        self._check(self._L.isocon_nn_pilot_near_dev(self._h, ctypes.byref(p), ctypes.byref(c)))
        return _DevArray(p.value, c.value, "<i8") if c.value else None

    def ipc_handles(self):
        """(handle record of IPC_BYTES: best[], the tile queues, the share block + its layout; generation number)."""
        h = np.zeros(IPC_BYTES, np.uint8)
        gen = ctypes.c_uint64(0)
        self._check(self._L.isocon_nn_ipc_handles(self._h, h.ctypes.data, ctypes.byref(gen)))
        return h, int(gen.value)

    def set_peers(self, handles, world, rank):
        """handles: uint8[world, IPC_BYTES] in rank order; world <= 1 closes the peer mappings."""
        h = np.ascontiguousarray(handles, dtype=np.uint8) if world > 1 else np.zeros(IPC_BYTES, np.uint8)
        self._check(self._L.isocon_nn_set_peers(self._h, h.ctypes.data, int(world), int(rank)))

    def can_fuse(self):
        v = ctypes.c_int32(0)
        self._check(self._L.isocon_nn_can_fuse(self._h, ctypes.byref(v)))
        return bool(v.value)

    def release_retired(self):
        self._check(self._L.isocon_nn_release_retired(self._h))

    def graph_finalize(self):
        ne = ctypes.c_int64(0)
        self._check(self._L.isocon_nn_graph_finalize(self._h, ctypes.byref(ne)))
        self._n_edges = ne.value
        return ne.value

    def graph_fetch(self):
        ne = self._n_edges
        best = np.empty(max(self.n, 1), np.int32)
        eq = np.empty(max(ne, 1), np.int32); et = np.empty(max(ne, 1), np.int32); ed = np.empty(max(ne, 1), np.int32)
        self._check(self._L.isocon_nn_graph_fetch(self._h, best.ctypes.data, eq.ctypes.data, et.ctypes.data,
                                                  ed.ctypes.data))
        return best[:self.n], eq[:ne], et[:ne], ed[:ne]

    def edges_dev(self):
        q, t, d = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        self._check(self._L.isocon_nn_edges_dev(self._h, ctypes.byref(q), ctypes.byref(t), ctypes.byref(d)))
        ne = self._n_edges
        return _DevArray(q.value, ne), _DevArray(t.value, ne), _DevArray(d.value, ne)

    def graph(self, mode, depth, is_query, is_target=None, algo=ALGO_AUTO, symmetric=True):
        """Single-GPU graph: (best[n], edge_q, edge_t, edge_d) with edges unordered.  A candidate-edge buffer that
        proves too small (tie-heavy input) is regrown and the graph rebuilt."""
        for _ in range(8):
            self.graph_begin(mode, depth, is_query, is_target, algo, symmetric)
            self.graph_run(PHASE_ALL)
            try:
                self.graph_finalize()
            except IsoconNNError as e:
                if e.code != ERR_OVERFLOW:
                    raise
                self.reserve_edges(self.stats()["edges_raw"] * 3 // 2 + 4096)
                continue
            return self.graph_fetch()
        raise IsoconNNError(ERR_OVERFLOW, "candidate edge buffer still too small after regrowing")

    # ------------------------------------------------------------------ misc
    def ed_pairs(self, a, b, k=None):
        a = np.ascontiguousarray(a, dtype=np.int32); b = np.ascontiguousarray(b, dtype=np.int32)
        kk = None if k is None else np.ascontiguousarray(k, dtype=np.int32)
        out = np.empty(max(a.size, 1), np.int32)
        self._check(self._L.isocon_nn_ed_pairs(self._h, a.ctypes.data, b.ctypes.data,
                                               None if kk is None else kk.ctypes.data, a.size, out.ctypes.data))
        return out[:a.size]

    def stats(self):
        s = _Stats()
        self._check(self._L.isocon_nn_get_stats(self._h, ctypes.byref(s)))
        return {name: int(getattr(s, name)) for name, _ in _Stats._fields_}

    def last_ms(self, which):
        """Device time of the last set_reads (0), graph_run (1), finalize (2), ed_pairs (3), probe (4)."""
        ms = ctypes.c_float(0)
        self._check(self._L.isocon_nn_last_ms(self._h, which, ctypes.byref(ms)))
        return ms.value

    def sync(self):
        self._check(self._L.isocon_nn_sync(self._h))

    def timer_start(self):
        self._check(self._L.isocon_nn_timer_start(self._h))

    def timer_stop(self):
        """Device milliseconds since timer_start (CUDA events on the library's stream)."""
        ms = ctypes.c_float(0)
        self._check(self._L.isocon_nn_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    def int32_peak(self):
        v = ctypes.c_double(0)
        self._check(self._L.isocon_nn_int32_peak(self._h, ctypes.byref(v)))
        return v.value


_CONTEXTS = {}


def default_device():
    for var in ("ISOCON_NN_DEVICE", "LOCAL_RANK"):
        if os.environ.get(var, "") != "":
            return int(os.environ[var])
    return 0


def get_context(device=None):
    """Process-global context per device, created lazily and reused across graph builds."""
    if device is None:
        device = default_device()
    ctx = _CONTEXTS.get(device)
    if ctx is None:
        ctx = _CONTEXTS[device] = NNContext(device)
    return ctx
