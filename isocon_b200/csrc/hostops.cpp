// hostops.cpp -- host-side helpers of the Python mirror (isocon_b200/nearest_neighbor_graph.py), loaded with
// ctypes.PyDLL (the GIL is held; arguments are PyObject* / plain pointers).  No arithmetic of the path lives
// here: these are the loops over Python lists that the reference does in the interpreter
// (/root/reference/modules/nearest_neighbor_graph.py:243-246, :202-208 list building; :75-79, :328-332 dict merging),
// done at memcpy speed so that the reference-facing call costs little more than the device step:
//
//   iso_host_lengths      len() of every string of a list                          (sort key of :246 / :208)
//   iso_host_lookup       slot of every string in the resident read store          (content-keyed residency)
//   iso_host_register     enter freshly uploaded strings into the store's dict
//   iso_host_permute      a list reordered by the stable length sort             (:246 / :208)
//   iso_host_contains     membership mask of a list in a dict / set               (has_converged :121, targets :350)
//   iso_host_gather       ASCII bytes of selected strings, concatenated, straight into the pinned upload buffer
//   iso_host_prepare_graph / iso_host_fill_graph
//                         dict-of-dicts result in the reference's key and insertion order from the device's
//                         unordered edge list (scan order: per query by |t - q|, down before up); the empty dicts are
//                         made while the device works, the edges filled in afterwards
//
// Errors are Python exceptions (set here, raised by ctypes.PyDLL after the call).
#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {

// out[i] = len(seqs[i]); returns the sum, or -1 with an exception set.
long long iso_host_lengths(PyObject* seqs, long long* out) {
    PyObject* fast = PySequence_Fast(seqs, "expected a sequence of str");
    if (!fast) return -1;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject** items = PySequence_Fast_ITEMS(fast);
    long long total = 0;
    for (Py_ssize_t i = 0; i < n; ++i) {
        if (!PyUnicode_Check(items[i])) {
            Py_DECREF(fast);
            PyErr_Format(PyExc_TypeError, "entry %zd of the read list is not a str", i);
            return -1;
        }
        const long long l = (long long)PyUnicode_GET_LENGTH(items[i]);
        out[i] = l; total += l;
    }
    Py_DECREF(fast);
    return total;
}

// slots[i] = store[seqs[i]] if present else -1; returns the number of missing entries, or -1 on error.
long long iso_host_lookup(PyObject* store, PyObject* seqs, int32_t* slots) {
    if (!PyDict_Check(store)) { PyErr_SetString(PyExc_TypeError, "store must be a dict"); return -1; }
    PyObject* fast = PySequence_Fast(seqs, "expected a sequence of str");
    if (!fast) return -1;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject** items = PySequence_Fast_ITEMS(fast);
    long long missing = 0;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* v = PyDict_GetItemWithError(store, items[i]);   // borrowed
        if (v) {
            slots[i] = (int32_t)PyLong_AsLong(v);
        } else {
            if (PyErr_Occurred()) { Py_DECREF(fast); return -1; }
            slots[i] = -1; ++missing;
        }
    }
    Py_DECREF(fast);
    return missing;
}

// store[seqs[sel[k]]] = first_slot + k  (sel == NULL: k-th entry of seqs).  A string that is selected twice keeps
// the slot of its LAST occurrence (both slots hold the same bytes).  Returns 0, or -1 on error.
int iso_host_register(PyObject* store, PyObject* seqs, const int32_t* sel, long long nsel, long long first_slot) {
    if (!PyDict_Check(store)) { PyErr_SetString(PyExc_TypeError, "store must be a dict"); return -1; }
    PyObject* fast = PySequence_Fast(seqs, "expected a sequence of str");
    if (!fast) return -1;
    PyObject** items = PySequence_Fast_ITEMS(fast);
    for (long long k = 0; k < nsel; ++k) {
        PyObject* v = PyLong_FromLongLong(first_slot + k);
        if (!v || PyDict_SetItem(store, items[sel ? sel[k] : k], v) < 0) { Py_XDECREF(v); Py_DECREF(fast); return -1; }
        Py_DECREF(v);
    }
    Py_DECREF(fast);
    return 0;
}

// [items[order[0]], items[order[1]], ...] as a new list (the stable length sort of :246 / :208 applied to a parallel list).
PyObject* iso_host_permute(PyObject* items, const int64_t* order, long long n) {
    PyObject* fast = PySequence_Fast(items, "expected a sequence");
    if (!fast) return NULL;
    const long long have = (long long)PySequence_Fast_GET_SIZE(fast);
    PyObject** src = PySequence_Fast_ITEMS(fast);
    PyObject* out = PyList_New((Py_ssize_t)n);
    if (!out) { Py_DECREF(fast); return NULL; }
    for (long long i = 0; i < n; ++i) {
        if (order[i] < 0 || order[i] >= have) {
            Py_DECREF(out); Py_DECREF(fast);
            PyErr_SetString(PyExc_IndexError, "permutation index out of range");
            return NULL;
        }
        PyObject* o = src[order[i]];
        Py_INCREF(o);
        PyList_SET_ITEM(out, (Py_ssize_t)i, o);
    }
    Py_DECREF(fast);
    return out;
}

// out[i] = 1 if keys[i] in container else 0 (container: dict, set or anything with __contains__).  Returns the number of
// hits, or -1 on error.
long long iso_host_contains(PyObject* container, PyObject* keys, unsigned char* out) {
    PyObject* fast = PySequence_Fast(keys, "expected a sequence");
    if (!fast) return -1;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject** items = PySequence_Fast_ITEMS(fast);
    long long hits = 0;
    for (Py_ssize_t i = 0; i < n; ++i) {
        const int r = PySequence_Contains(container, items[i]);
        if (r < 0) { Py_DECREF(fast); return -1; }
        out[i] = (unsigned char)r; hits += r;
    }
    Py_DECREF(fast);
    return hits;
}

// dst <- bytes of seqs[sel[0]], seqs[sel[1]], ... (sel == NULL: all nsel leading entries); offsets[0..nsel] are the
// running byte offsets.  Strings must be ASCII (one byte per character).  Returns the total, or -1 on error.
// Large inputs are copied by a few threads with the GIL released (the strings stay alive: the caller holds the list).
long long iso_host_gather(PyObject* seqs, const int32_t* sel, long long nsel, unsigned char* dst, long long cap,
                          long long* offsets) {
    PyObject* fast = PySequence_Fast(seqs, "expected a sequence of str");
    if (!fast) return -1;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject** items = PySequence_Fast_ITEMS(fast);
    std::vector<const unsigned char*> src((size_t)nsel);
    long long at = 0;
    offsets[0] = 0;
    for (long long k = 0; k < nsel; ++k) {
        const long long i = sel ? sel[k] : k;
        if (i < 0 || i >= n || !PyUnicode_Check(items[i])) {
            Py_DECREF(fast);
            PyErr_Format(PyExc_TypeError, "entry %lld of the read list is not a str", i);
            return -1;
        }
        PyObject* s = items[i];
        if (!PyUnicode_IS_ASCII(s)) {
            Py_DECREF(fast);
            PyErr_Format(PyExc_ValueError, "read %lld is not an ASCII string", i);
            return -1;
        }
        src[(size_t)k] = PyUnicode_1BYTE_DATA(s);
        at += (long long)PyUnicode_GET_LENGTH(s);
        offsets[k + 1] = at;
    }
    if (at > cap) {
        Py_DECREF(fast);
        PyErr_SetString(PyExc_BufferError, "upload buffer too small");
        return -1;
    }
    auto copy_range = [&](long long k0, long long k1) {
        for (long long k = k0; k < k1; ++k) memcpy(dst + offsets[k], src[(size_t)k], (size_t)(offsets[k + 1] - offsets[k]));
    };
    const int threads = at >= (64ll << 20) ? 4 : (at >= (24ll << 20) ? 2 : 1);   // several ranks share the host's cores
    if (threads == 1) {
        copy_range(0, nsel);
    } else {
        Py_BEGIN_ALLOW_THREADS
        std::vector<std::thread> pool;
        long long k0 = 0;
        for (int t = 0; t < threads; ++t) {          // equal byte shares
            const long long want = at * (t + 1) / threads;
            long long k1 = std::upper_bound(offsets, offsets + nsel + 1, want) - offsets - 1;
            if (t == threads - 1) k1 = nsel;
            k1 = std::max(k1, k0);
            pool.emplace_back(copy_range, k0, k1);
            k0 = k1;
        }
        for (auto& th : pool) th.join();
        Py_END_ALLOW_THREADS
    }
    Py_DECREF(fast);
    return at;
}

// The reference's result, step 1 (independent of the device: runs while the kernels do): out[accs[i]] = {} for every
// i in [lo, hi) with skip[i] == 0 (skip may be NULL), in list order (nearest_neighbor_graph.py:120, :355).
// dicts (optional, hi - lo entries): the address of each entry's dict (0 = skipped), so that iso_host_fill_graph
// need not look every query up again; all 0 when an accession occurs twice (then the dict that won the key is found
// by lookup, as before).  Borrowed pointers: valid while `out` is alive and unchanged.
PyObject* iso_host_prepare_graph(PyObject* accs, long long lo, long long hi, const unsigned char* skip, int64_t* dicts) {
    PyObject* fast = PySequence_Fast(accs, "expected a sequence of accessions");
    if (!fast) return NULL;
    const long long n = (long long)PySequence_Fast_GET_SIZE(fast);
    PyObject** items = PySequence_Fast_ITEMS(fast);
    if (lo < 0 || hi > n || lo > hi) {
        Py_DECREF(fast);
        PyErr_SetString(PyExc_ValueError, "key range outside the list");
        return NULL;
    }
    PyObject* out = PyDict_New();
    if (!out) { Py_DECREF(fast); return NULL; }
    // (no cyclic-GC passes over a hundred thousand fresh dicts that cannot be garbage: a third of this loop's time)
    const int gc_was_on = PyGC_Disable();
    long long made = 0;
    for (long long i = lo; i < hi; ++i) {
        if (dicts) dicts[i - lo] = 0;
        if (skip && skip[i]) continue;
        PyObject* d = PyDict_New();
        if (dicts) dicts[i - lo] = (int64_t)(intptr_t)d;
        ++made;
        // a repeated accession keeps ONE dict, like the reference's assignment to the same key
        if (!d || PyDict_SetItem(out, items[i], d) < 0) {
            Py_XDECREF(d); Py_DECREF(out); Py_DECREF(fast);
            if (gc_was_on) PyGC_Enable();
            return NULL;
        }
        Py_DECREF(d);
    }
    if (gc_was_on) PyGC_Enable();
    if (dicts && (long long)PyDict_GET_SIZE(out) != made) memset(dicts, 0, (size_t)(hi - lo) * sizeof(int64_t));
    Py_DECREF(fast);
    return out;
}

// Step 2: out[accs[q]][accs[t]] = d for the edges in SCAN ORDER: per query by offset |t - q|, down (t < q) before
// up (nearest_neighbor_graph.py:134-185 / :359-410).  The device reports edges unordered and possibly twice.
// Returns 0, or -1 with an exception set.
int iso_host_fill_graph(PyObject* out, PyObject* accs, long long lo, long long hi, const int32_t* eq, const int32_t* et,
                        const int32_t* ed, long long ne, const int64_t* dicts) {
    if (!PyDict_Check(out)) { PyErr_SetString(PyExc_TypeError, "graph must be a dict"); return -1; }
    PyObject* fast = PySequence_Fast(accs, "expected a sequence of accessions");
    if (!fast) return -1;
    const long long n = (long long)PySequence_Fast_GET_SIZE(fast);
    PyObject** items = PySequence_Fast_ITEMS(fast);
    // sort key (q, |t - q|, up) in one 64-bit word; list indices are < 2**31.  Most queries have one or two edges:
    // a counting sort by query, then a sort inside each query's few entries.
    std::vector<uint32_t> first((size_t)(hi - lo) + 2, 0);
    for (long long e = 0; e < ne; ++e) {
        const int64_t q = eq[e], t = et[e];
        if (q < lo || q >= hi || t < 0 || t >= n) {
            Py_DECREF(fast);
            PyErr_Format(PyExc_ValueError, "edge %lld (%lld -> %lld) lies outside the key range", e, (long long)q, (long long)t);
            return -1;
        }
        ++first[(size_t)(q - lo) + 2];
    }
    for (size_t i = 2; i < first.size(); ++i) first[i] += first[i - 1];      // first[q - lo + 1] = start of q's entries
    std::vector<std::pair<uint64_t, int32_t>> order((size_t)ne);
    for (long long e = 0; e < ne; ++e) {
        const int64_t q = eq[e], t = et[e];
        const uint64_t off = (uint64_t)(t > q ? t - q : q - t);
        order[first[(size_t)(q - lo) + 1]++] = std::make_pair(((uint64_t)q << 33) | (off << 1) | (uint64_t)(t > q), ed[e]);
    }
    for (size_t i = 0; i + 1 < first.size(); ++i)                              // now first[q - lo] .. first[q - lo + 1]
        if (first[i + 1] - first[i] > 1) std::sort(order.begin() + first[i], order.begin() + first[i + 1]);
    uint64_t prev = ~0ull;
    int64_t cur_q = -1;
    PyObject* target = NULL;      // borrowed from `out`
    for (const auto& kv : order) {
        if (kv.first == prev) continue;       // an edge may be reported twice
        prev = kv.first;
        const int64_t q = (int64_t)(kv.first >> 33);
        const int64_t off = (int64_t)((kv.first >> 1) & 0xffffffffull);
        const int64_t t = (kv.first & 1) ? q + off : q - off;
        if (q != cur_q) {
            // when two entries of the list carry the same accession the later dict won the key: write there, as the
            // reference's best_edit_distances[acc1][acc2] = ... does
            target = (dicts && dicts[q - lo]) ? (PyObject*)(intptr_t)dicts[q - lo] : PyDict_GetItemWithError(out, items[q]);
            cur_q = q;
            if (!target) {
                if (!PyErr_Occurred()) PyErr_Format(PyExc_ValueError, "edge of entry %lld, which is not a query of this call", (long long)q);
                Py_DECREF(fast);
                return -1;
            }
        }
        PyObject* v = PyLong_FromLong((long)kv.second);
        if (!v || PyDict_SetItem(target, items[t], v) < 0) { Py_XDECREF(v); Py_DECREF(fast); return -1; }
        Py_DECREF(v);
    }
    Py_DECREF(fast);
    return 0;
}

}  // extern "C"
