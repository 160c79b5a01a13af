/*
 * oracle/edlib_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A CPython module named `edlib` exposing the one call the reference's hot path makes:
 *     edlib.align(query, target, mode="NW", task="distance", k=-1) -> {"editDistance": d, ...}
 * (/root/reference/modules/nearest_neighbor_graph.py:105).  It lets the UNMODIFIED
 * reference module be imported and run in the authoring container, where the real edlib
 * (PyPI, >=1.1.2) is absent.  Only mode="NW", task="distance" is implemented; anything else
 * raises.  Arithmetic: oracle/levenshtein.h (ed_myers64, or the plain DP when the
 * environment variable ISOCON_ORACLE_PLAIN=1 is set -- used to cross-check the two).
 *
 * It is an edlib-COMPATIBLE STAND-IN, never reported as "edlib".
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include "levenshtein.h"

static int g_use_plain = 0;
static unsigned long long g_calls = 0, g_neg = 0, g_cells_full = 0, g_cells_band = 0;

static PyObject* shim_align(PyObject*, PyObject* args, PyObject* kwargs) {
    static const char* kw[] = {"query", "target", "mode", "task", "k", "additionalEqualities", NULL};
    PyObject *q, *t, *extra = Py_None;
    const char *mode = "NW", *task = "distance";
    long k = -1;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "OO|sslO", (char**)kw, &q, &t, &mode, &task, &k, &extra))
        return NULL;
    if (strcmp(mode, "NW") != 0 || strcmp(task, "distance") != 0 || extra != Py_None) {
        PyErr_SetString(PyExc_NotImplementedError, "edlib shim: only mode='NW', task='distance' is provided");
        return NULL;
    }
    Py_ssize_t m, n;
    const char *x, *y;
    if (PyUnicode_Check(q)) { x = PyUnicode_AsUTF8AndSize(q, &m); if (!x) return NULL; }
    else if (PyBytes_Check(q)) { x = PyBytes_AS_STRING(q); m = PyBytes_GET_SIZE(q); }
    else { PyErr_SetString(PyExc_TypeError, "query must be str or bytes"); return NULL; }
    if (PyUnicode_Check(t)) { y = PyUnicode_AsUTF8AndSize(t, &n); if (!y) return NULL; }
    else if (PyBytes_Check(t)) { y = PyBytes_AS_STRING(t); n = PyBytes_GET_SIZE(t); }
    else { PyErr_SetString(PyExc_TypeError, "target must be str or bytes"); return NULL; }

    int d;
    Py_BEGIN_ALLOW_THREADS
    if (g_use_plain) {
        d = isocon_oracle::ed_plain((const uint8_t*)x, (int)m, (const uint8_t*)y, (int)n);
        if (k >= 0 && d > k) d = -1;
    } else {
        d = isocon_oracle::ed_myers64((const uint8_t*)x, (int)m, (const uint8_t*)y, (int)n, (int)k);
    }
    Py_END_ALLOW_THREADS

    /* implementation-independent work counters */
    g_calls++;
    if (d < 0) g_neg++;
    g_cells_full += (unsigned long long)m * (unsigned long long)n;
    {
        const long kk = k < 0 ? (long)std::max(m, n) : k;
        const long ad = labs((long)n - (long)m);
        if (ad <= kk) {
            const long w = ad + 2 * ((kk - ad) / 2) + 1;
            g_cells_band += (unsigned long long)n * (unsigned long long)std::min<long>((long)m, w);
        }
    }
    return Py_BuildValue("{s:i,s:O,s:O,s:i}", "editDistance", d, "locations", Py_None, "cigar", Py_None,
                         "alphabetLength", 4);
}

static PyObject* shim_counters(PyObject*, PyObject*) {
    return Py_BuildValue("{s:K,s:K,s:K,s:K}", "calls", g_calls, "neg", g_neg,
                         "cells_full", g_cells_full, "cells_band", g_cells_band);
}

static PyObject* shim_reset(PyObject*, PyObject*) {
    g_calls = g_neg = g_cells_full = g_cells_band = 0;
    Py_RETURN_NONE;
}

static PyMethodDef methods[] = {
    {"align", (PyCFunction)(void (*)(void))shim_align, METH_VARARGS | METH_KEYWORDS,
     "edlib-compatible align(query, target, mode='NW', task='distance', k=-1)"},
    {"_counters", shim_counters, METH_NOARGS, "work counters since last _reset()"},
    {"_reset", shim_reset, METH_NOARGS, "zero the work counters"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "edlib",
                                    "edlib-compatible stand-in (oracle; NW distance only)", -1, methods};

PyMODINIT_FUNC PyInit_edlib(void) {
    const char* e = getenv("ISOCON_ORACLE_PLAIN");
    g_use_plain = (e && e[0] == '1');
    PyObject* mod = PyModule_Create(&moddef);
    if (mod) PyModule_AddStringConstant(mod, "__shim__", "isocon_b200 oracle stand-in");
    return mod;
}
