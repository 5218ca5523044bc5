/* _fastpack: CPython helper of the host packing path (plumbing, no arithmetic).
 *
 * The reference keeps every trial / segment as its own NumPy array (a list of dicts).  Handing thousands of small
 * arrays to the C ABI one ctypes pointer at a time costs ~2 us each in pure Python; this module walks the sequence
 * through the buffer protocol in C (~50 ns each) and returns the data pointers and row counts as two packed byte
 * strings that vlgp_trials_set_y_parts / vlgp_trials_{set,get}_state_parts take directly.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

/* pointers(seq, itemsize, ncols, writable, fmt) -> (ptrs: bytes of uint64, rows: bytes of int64)
 * Every item must expose a C-contiguous 2-D buffer with the given itemsize, second dimension and (when fmt is a
 * non-empty string) struct format character, e.g. "d" for float64, "B" for uint8. */
static PyObject *fp_pointers(PyObject *self, PyObject *args) {
    PyObject *seq;
    Py_ssize_t itemsize, ncols;
    int writable = 0;
    const char *fmt = "";
    if (!PyArg_ParseTuple(args, "Onn|ps", &seq, &itemsize, &ncols, &writable, &fmt)) return NULL;
    PyObject *fast = PySequence_Fast(seq, "expected a sequence of arrays");
    if (!fast) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject *ptrs = PyBytes_FromStringAndSize(NULL, n * (Py_ssize_t)sizeof(uint64_t));
    PyObject *rows = PyBytes_FromStringAndSize(NULL, n * (Py_ssize_t)sizeof(int64_t));
    if (!ptrs || !rows) goto fail;
    {
        uint64_t *pp = (uint64_t *)PyBytes_AS_STRING(ptrs);
        int64_t *pr = (int64_t *)PyBytes_AS_STRING(rows);
        const int flags = PyBUF_C_CONTIGUOUS | PyBUF_FORMAT | (writable ? PyBUF_WRITABLE : 0);
        for (Py_ssize_t i = 0; i < n; ++i) {
            Py_buffer view;
            if (PyObject_GetBuffer(PySequence_Fast_GET_ITEM(fast, i), &view, flags) != 0) goto fail;
            int ok = view.ndim == 2 && view.itemsize == itemsize && view.shape[1] == ncols;
            if (ok && fmt[0]) {
                const char *f = view.format ? view.format : "B";
                while (*f == '<' || *f == '>' || *f == '=' || *f == '@' || *f == '!') ++f;
                ok = (f[0] == fmt[0]);
            }
            if (!ok) {
                PyBuffer_Release(&view);
                PyErr_Format(PyExc_TypeError, "item %zd: expected a C-contiguous (rows, %zd) array of %zd-byte items", i,
                             ncols, itemsize);
                goto fail;
            }
            pp[i] = (uint64_t)(uintptr_t)view.buf;
            pr[i] = (int64_t)view.shape[0];
            PyBuffer_Release(&view);      /* the caller keeps the arrays alive for the duration of the C-ABI call */
        }
    }
    Py_DECREF(fast);
    return Py_BuildValue("(NN)", ptrs, rows);
fail:
    Py_XDECREF(ptrs);
    Py_XDECREF(rows);
    Py_DECREF(fast);
    return NULL;
}

/* format_char(obj) -> the struct format character of a buffer ('d', 'B', ...), or '?' */
static PyObject *fp_format_char(PyObject *self, PyObject *obj) {
    Py_buffer view;
    if (PyObject_GetBuffer(obj, &view, PyBUF_FORMAT | PyBUF_ND) != 0) return NULL;
    const char *f = view.format ? view.format : "B";
    while (*f == '<' || *f == '>' || *f == '=' || *f == '@' || *f == '!') ++f;
    PyObject *r = PyUnicode_FromStringAndSize(f, 1);
    PyBuffer_Release(&view);
    return r;
}

/* block_owners(seq, ndim, axis, extent) -> list of the DISTINCT arrays owning the memory of the blocks in seq (None
 * items are skipped): item.base when that is an object of the item's own type (a view), else the item itself.
 * Every block must have `ndim` dimensions and shape[axis] == extent, else ValueError.  Used for the regressor check:
 * thousands of segments are views of a few hundred trial-level arrays, and only those need to be scanned. */
static PyObject *fp_block_owners(PyObject *self, PyObject *args) {
    PyObject *seq;
    int ndim, axis;
    Py_ssize_t extent;
    if (!PyArg_ParseTuple(args, "Oiin", &seq, &ndim, &axis, &extent)) return NULL;
    if (axis < 0 || axis >= ndim) {
        PyErr_SetString(PyExc_ValueError, "axis out of range");
        return NULL;
    }
    PyObject *fast = PySequence_Fast(seq, "expected a sequence of arrays");
    if (!fast) return NULL;
    PyObject *seen = PyDict_New();
    PyObject *out = NULL;
    static PyObject *base_name = NULL;
    if (!base_name) base_name = PyUnicode_InternFromString("base");
    if (!seen || !base_name) goto fail;
    {
        const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
        PyObject *last = NULL;      /* borrowed: kept alive by `seen`; consecutive blocks usually share their owner */
        for (Py_ssize_t i = 0; i < n; ++i) {
            PyObject *item = PySequence_Fast_GET_ITEM(fast, i);
            if (item == Py_None) continue;
            Py_buffer view;
            if (PyObject_GetBuffer(item, &view, PyBUF_STRIDES) != 0) goto fail;
            const int ok = view.ndim == ndim && view.shape[axis] == extent;
            PyBuffer_Release(&view);
            if (!ok) {
                PyErr_Format(PyExc_ValueError, "item %zd: expected %d dimensions with shape[%d] == %zd", i, ndim, axis,
                             extent);
                goto fail;
            }
            PyObject *owner = PyObject_GetAttr(item, base_name);
            if (!owner) {
                PyErr_Clear();
                owner = item;
                Py_INCREF(owner);
            } else if (Py_TYPE(owner) != Py_TYPE(item)) {
                Py_DECREF(owner);
                owner = item;
                Py_INCREF(owner);
            }
            if (owner == last) {
                Py_DECREF(owner);
                continue;
            }
            PyObject *key = PyLong_FromVoidPtr((void *)owner);
            if (!key) {
                Py_DECREF(owner);
                goto fail;
            }
            const int rc = PyDict_SetItem(seen, key, owner);      /* the dict holds its own references */
            Py_DECREF(key);
            Py_DECREF(owner);
            if (rc != 0) goto fail;
            last = owner;
        }
    }
    out = PyDict_Values(seen);
fail:
    Py_XDECREF(seen);
    Py_DECREF(fast);
    return out;
}

static PyMethodDef methods[] = {
    {"block_owners", fp_block_owners, METH_VARARGS, "block_owners(seq, ndim, axis, extent) -> distinct owning arrays"},
    {"pointers", fp_pointers, METH_VARARGS, "pointers(seq, itemsize, ncols, writable=False) -> (ptr bytes, row bytes)"},
    {"format_char", fp_format_char, METH_O, "struct format character of a buffer"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastpack", "host packing helper of vlgp_b200", -1, methods};

PyMODINIT_FUNC PyInit__fastpack(void) { return PyModule_Create(&moddef); }
