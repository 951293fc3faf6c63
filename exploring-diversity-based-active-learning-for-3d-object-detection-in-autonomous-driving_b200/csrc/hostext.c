/* hostext.c -- CPython helper for the host-side extraction of the selector inputs.
 *
 * Replaces, for speed only, the per-frame python loops of det3d/selectors/spatial_temporal_selector.py:78-91
 * (ego location from ``car_from_global``, logfile name from ``cam_front_path``) and base_selector.py:85 (box count)
 * that adsel_b200/pool.py otherwise performs with list comprehensions: at ~1 us per frame and field they are most
 * of the end-to-end time once the selection itself takes 10 ms.  Pure host code, no CUDA; pool.py falls back to
 * its python implementation when this module is not built, and checks the ego locations of a probe of frames
 * bit-for-bit against the reference's literal numpy expression before trusting them.
 *
 *   extract(infos, lo, hi, fused) -> (loc: bytes f64[m,2], boxes: bytes i32[m], names: list[str], change: bytes u8[m])
 *
 * change[i] = 1 when names[i] != names[i-1] (change[0] = 1): the consecutive-run boundaries of
 * spatial_temporal_selector.py:114-129, found while the names are parsed.
 *
 * loc[j] = -(t0*R0j + t1*R1j + t2*R2j) with t = cal[:3,3], R = cal[:3,:3], evaluated in index order either with
 * fused multiply-adds (what numpy's vector-matrix product does on FMA hosts) or with separately rounded
 * products (fused = 0).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef ADSEL_HAVE_NUMPY_HEADERS
/* Direct access to the fields of a numpy.ndarray (data, shape, strides, dtype number): the buffer protocol costs
 * more than everything else this function does per frame.  Only exact ``numpy.ndarray`` instances take this path, it
 * is validated against the buffer protocol on the first array it sees, and pool.py checks the resulting ego
 * locations against the reference expression before trusting the helper at all. */
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/ndarraytypes.h>
static int g_fast = -1; /* -1 not validated yet, 0 disabled, 1 enabled */
static int is_exact_ndarray(PyObject *o) { return strcmp(Py_TYPE(o)->tp_name, "numpy.ndarray") == 0; }
static int read_cal_fields(PyObject *cal, double m[4][4]) {
    const PyArrayObject_fields *a = (const PyArrayObject_fields *)cal;
    if (a->nd != 2 || a->dimensions[0] != 4 || a->dimensions[1] != 4) return -1;
    if (a->strides[0] != 32 || a->strides[1] != 8) return -1;   /* C order only, see read_cal */
    const PyArray_Descr *d = a->descr;
    if (d->type_num != NPY_DOUBLE || d->byteorder == '>') return -1;
    const char *base = a->data;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) memcpy(&m[i][j], base + i * a->strides[0] + j * a->strides[1], 8);
    return 0;
}
#endif

static int read_cal(PyObject *cal, double m[4][4]) {
    Py_buffer view;
    if (PyObject_GetBuffer(cal, &view, PyBUF_STRIDES | PyBUF_FORMAT) != 0) return -1;
    /* Only the one layout real pools have -- a C-ordered 4x4 float64 matrix -- is accepted: numpy's vector-matrix
     * product takes a different BLAS kernel (different rounding order) for Fortran-ordered or strided operands, and a
     * probe of a few dozen frames could not vouch for a pool that mixes layouts.  Anything else raises TypeError and
     * pool.py evaluates the reference's literal expression instead. */
    int ok = view.ndim == 2 && view.shape[0] == 4 && view.shape[1] == 4 && view.itemsize == 8 && view.format &&
             view.strides[0] == 32 && view.strides[1] == 8 &&
             (strcmp(view.format, "d") == 0 || strcmp(view.format, "<d") == 0 || strcmp(view.format, "=d") == 0);
    if (ok) {
        const char *base = (const char *)view.buf;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 4; ++j) memcpy(&m[i][j], base + i * view.strides[0] + j * view.strides[1], 8);
    }
    PyBuffer_Release(&view);
    if (!ok) {
        PyErr_SetString(PyExc_TypeError, "car_from_global is not a C-ordered 4x4 float64 matrix");
        return -1;
    }
    return 0;
}

/* basename up to the first "__": path.split("/")[-1].split("__")[0] */
static PyObject *logfile_of(PyObject *path) {
    if (!PyUnicode_Check(path)) {
        PyErr_SetString(PyExc_TypeError, "cam_front_path is not a str");
        return NULL;
    }
    const Py_ssize_t len = PyUnicode_GET_LENGTH(path);
    Py_ssize_t slash = PyUnicode_FindChar(path, '/', 0, len, -1);
    if (slash == -2) return NULL;
    const Py_ssize_t start = slash + 1; /* -1 + 1 = 0 when there is no "/" */
    Py_ssize_t end = len;
    for (Py_ssize_t p = start; p + 1 < len;) {
        const Py_ssize_t u = PyUnicode_FindChar(path, '_', p, len, 1);
        if (u == -2) return NULL;
        if (u < 0 || u + 1 >= len) break;
        if (PyUnicode_READ_CHAR(path, u + 1) == '_') {
            end = u;
            break;
        }
        p = u + 1;
    }
    return PyUnicode_Substring(path, start, end);
}

static PyObject *extract(PyObject *self, PyObject *args) {
    PyObject *infos;
    Py_ssize_t lo, hi;
    int fused;
    if (!PyArg_ParseTuple(args, "Onnp", &infos, &lo, &hi, &fused)) return NULL;
    if (!PyList_Check(infos)) {
        PyErr_SetString(PyExc_TypeError, "infos must be a list");
        return NULL;
    }
    const Py_ssize_t n = PyList_GET_SIZE(infos);
    if (lo < 0 || hi > n || lo > hi) {
        PyErr_SetString(PyExc_ValueError, "bad range");
        return NULL;
    }
    const Py_ssize_t m = hi - lo;
    PyObject *loc_b = PyBytes_FromStringAndSize(NULL, m * 16), *box_b = PyBytes_FromStringAndSize(NULL, m * 4);
    PyObject *chg_b = PyBytes_FromStringAndSize(NULL, m);
    PyObject *pose_b = PyBytes_FromStringAndSize(NULL, m * 72); /* the nine pose entries the location uses, per frame */
    PyObject *names = PyList_New(m);
    PyObject *k_cal = PyUnicode_InternFromString("car_from_global"), *k_path = PyUnicode_InternFromString("cam_front_path"),
             *k_gt = PyUnicode_InternFromString("gt_names");
    if (!loc_b || !box_b || !chg_b || !pose_b || !names || !k_cal || !k_path || !k_gt) goto fail;
    double *loc = (double *)PyBytes_AS_STRING(loc_b);
    int32_t *box = (int32_t *)PyBytes_AS_STRING(box_b);
    unsigned char *chg = (unsigned char *)PyBytes_AS_STRING(chg_b);
    double *pose = (double *)PyBytes_AS_STRING(pose_b);
    PyObject *last_name = NULL; /* names[i-1], borrowed from the list */
    PyObject *prev_name = NULL, *prev_path = NULL;
    Py_ssize_t prev_prefix = 0; /* length of "<dir>/<name>__" in prev_path */
    /* The infos of a freshly unpickled pool are cold: a frame costs ~8 dependent cache misses (dict, its key table,
     * the three value objects, the matrix data, the shape/stride blocks).  The loop is therefore software-pipelined
     * with prefetches: the dict of frame i+24, the key table of frame i+16, the value objects of frame i+8 (their
     * lookups are done then and kept in a small ring), the matrix data and shape blocks of frame i+4. */
    enum { RING = 32, D_DICT = 24, D_KEYS = 16, D_VALS = 8, D_DATA = 4 };
    PyObject *r_cal[RING], *r_path[RING], *r_gt[RING];
    for (Py_ssize_t s = -D_DICT; s < m; ++s) {
        if (s + D_DICT < m) __builtin_prefetch(PyList_GET_ITEM(infos, lo + s + D_DICT));
        if (s + D_KEYS >= 0 && s + D_KEYS < m) {
            PyObject *d = PyList_GET_ITEM(infos, lo + s + D_KEYS);
            if (PyDict_Check(d)) {
                const char *keys = (const char *)((PyDictObject *)d)->ma_keys;
                __builtin_prefetch(keys);
                __builtin_prefetch(keys + 64);
                __builtin_prefetch(keys + 128);
                __builtin_prefetch(keys + 192);
                __builtin_prefetch(keys + 256);
            }
        }
        if (s + D_VALS >= 0 && s + D_VALS < m) {
            const Py_ssize_t q = s + D_VALS;
            PyObject *d = PyList_GET_ITEM(infos, lo + q);
            PyObject *vc = NULL, *vp = NULL, *vg = NULL;
            if (PyDict_Check(d)) {
                vc = PyDict_GetItemWithError(d, k_cal);
                vp = PyDict_GetItemWithError(d, k_path);
                vg = PyDict_GetItemWithError(d, k_gt);
                if (PyErr_Occurred()) goto fail;
                if (vc) {
                    __builtin_prefetch(vc);
                    __builtin_prefetch((const char *)vc + 64);
                }
                if (vp) {
                    __builtin_prefetch(vp);
                    __builtin_prefetch((const char *)vp + 64);
                }
                if (vg) __builtin_prefetch(vg);
            }
            r_cal[q % RING] = vc;
            r_path[q % RING] = vp;
            r_gt[q % RING] = vg;
        }
#ifdef ADSEL_HAVE_NUMPY_HEADERS
        if (g_fast == 1 && s + D_DATA >= 0 && s + D_DATA < m) {
            PyObject *vc = r_cal[(s + D_DATA) % RING], *vg = r_gt[(s + D_DATA) % RING];
            if (vc && is_exact_ndarray(vc)) {
                const PyArrayObject_fields *a = (const PyArrayObject_fields *)vc;
                __builtin_prefetch(a->data);
                __builtin_prefetch(a->data + 64);
                __builtin_prefetch(a->dimensions);
            }
            if (vg && is_exact_ndarray(vg)) __builtin_prefetch(((const PyArrayObject_fields *)vg)->dimensions);
        }
#endif
        if (s < 0) continue;
        const Py_ssize_t i = s;
        PyObject *info = PyList_GET_ITEM(infos, lo + i);
        if (!PyDict_Check(info)) {
            PyErr_SetString(PyExc_TypeError, "infos entries must be dicts");
            goto fail;
        }
        PyObject *cal = r_cal[i % RING], *path = r_path[i % RING], *gt = r_gt[i % RING];
        if (!cal || !path || !gt) {
            PyErr_SetObject(PyExc_KeyError, !cal ? k_cal : !path ? k_path : k_gt);
            goto fail;
        }
        double c[4][4];
#ifdef ADSEL_HAVE_NUMPY_HEADERS
        int done = 0;
        if (g_fast != 0 && is_exact_ndarray(cal)) {
            if (g_fast < 0) { /* first array: the field access must agree with the buffer protocol */
                double ref[4][4];
                memset(ref, 0, sizeof(ref));
                memset(c, 0, sizeof(c));
                if (read_cal(cal, ref) != 0) goto fail;
                g_fast = (read_cal_fields(cal, c) == 0 && memcmp(ref, c, sizeof(double) * 12) == 0) ? 1 : 0;
                memcpy(c, ref, sizeof(ref));
                done = 1;
            } else if (read_cal_fields(cal, c) == 0) {
                done = 1;
            }
        }
        if (!done && read_cal(cal, c) != 0) goto fail;
#else
        if (read_cal(cal, c) != 0) goto fail;
#endif
        for (int r = 0; r < 3; ++r) { /* handed to the device, which redoes every frame (csrc/ego_check.cu) */
            pose[9 * i + 3 * r] = c[r][0];
            pose[9 * i + 3 * r + 1] = c[r][1];
            pose[9 * i + 3 * r + 2] = c[r][3];
        }
        for (int j = 0; j < 2; ++j) {
            double v;
            if (fused) v = fma(c[2][3], c[2][j], fma(c[1][3], c[1][j], c[0][3] * c[0][j]));
            else v = (c[0][3] * c[0][j] + c[1][3] * c[1][j]) + c[2][3] * c[2][j];
            loc[2 * i + j] = -v;
        }
        Py_ssize_t nb;
#ifdef ADSEL_HAVE_NUMPY_HEADERS
        if (g_fast == 1 && is_exact_ndarray(gt) && ((const PyArrayObject_fields *)gt)->nd >= 1)
            nb = ((const PyArrayObject_fields *)gt)->dimensions[0];
        else
#endif
            nb = PyObject_Length(gt);
        if (nb < 0) goto fail;
        box[i] = (int32_t)nb;
        /* same "<dir>/<name>__" prefix as the previous frame and no further "/": the name is the previous one */
        PyObject *name = NULL;
        if (prev_name && PyUnicode_Check(path) && PyUnicode_GET_LENGTH(path) >= prev_prefix) {
            int same = 1;
            if (PyUnicode_KIND(path) == PyUnicode_KIND(prev_path)) {
                same = memcmp(PyUnicode_DATA(path), PyUnicode_DATA(prev_path), (size_t)prev_prefix * PyUnicode_KIND(path)) == 0;
            } else {
                for (Py_ssize_t q = 0; q < prev_prefix && same; ++q)
                    same = PyUnicode_READ_CHAR(path, q) == PyUnicode_READ_CHAR(prev_path, q);
            }
            if (same) {
                const Py_ssize_t s2 = PyUnicode_FindChar(path, '/', prev_prefix, PyUnicode_GET_LENGTH(path), 1);
                if (s2 == -2) goto fail;
                if (s2 == -1) {
                    name = prev_name;
                    Py_INCREF(name);
                }
            }
        }
        if (!name) {
            name = logfile_of(path);
            if (!name) goto fail;
            /* the fast path is valid only if the basename really continues with "__" after the name */
            const Py_ssize_t len = PyUnicode_GET_LENGTH(path), nl = PyUnicode_GET_LENGTH(name);
            Py_ssize_t slash = PyUnicode_FindChar(path, '/', 0, len, -1);
            if (slash == -2) {
                Py_DECREF(name);
                goto fail;
            }
            const Py_ssize_t start = slash + 1;
            if (start + nl + 1 < len && PyUnicode_READ_CHAR(path, start + nl) == '_' && PyUnicode_READ_CHAR(path, start + nl + 1) == '_') {
                prev_name = name;
                prev_path = path;
                prev_prefix = start + nl + 2;
            } else {
                prev_name = NULL;
            }
        }
        if (last_name == NULL) {
            chg[i] = 1;
        } else if (name == last_name) {
            chg[i] = 0;
        } else {
            const int cmp = PyUnicode_Compare(name, last_name);
            if (cmp == -1 && PyErr_Occurred()) {
                Py_DECREF(name);
                goto fail;
            }
            chg[i] = cmp != 0;
        }
        last_name = name;
        PyList_SET_ITEM(names, i, name);
    }
    Py_DECREF(k_cal);
    Py_DECREF(k_path);
    Py_DECREF(k_gt);
    return Py_BuildValue("(NNNNN)", loc_b, box_b, names, chg_b, pose_b);
fail:
    Py_XDECREF(loc_b);
    Py_XDECREF(box_b);
    Py_XDECREF(chg_b);
    Py_XDECREF(pose_b);
    Py_XDECREF(names);
    Py_XDECREF(k_cal);
    Py_XDECREF(k_path);
    Py_XDECREF(k_gt);
    return NULL;
}

static PyMethodDef methods[] = {{"extract", extract, METH_VARARGS, "extract(infos, lo, hi, fused) -> (loc bytes, boxes bytes, names, change bytes, pose9 bytes)"},
                                {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_adsel_hostext", "host-side extraction helper", -1, methods};
PyMODINIT_FUNC PyInit__adsel_hostext(void) { return PyModule_Create(&moddef); }
