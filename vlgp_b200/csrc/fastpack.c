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

static PyMethodDef methods[] = {
    {"pointers", fp_pointers, METH_VARARGS, "pointers(seq, itemsize, ncols, writable=False) -> (ptr bytes, row bytes)"},
    {"format_char", fp_format_char, METH_O, "struct format character of a buffer"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastpack", "host packing helper of vlgp_b200", -1, methods};

PyMODINIT_FUNC PyInit__fastpack(void) { return PyModule_Create(&moddef); }
