/* Host ingest helper: pd.factorize(col, sort=False) for an object column of Python str.
 *
 * The reference scans label columns of Python strings with pandas (Trajectory.py:402-425: unique() and one boolean
 * mask per sample and per type); the GPU path needs integer codes.  pandas' object hashtable costs ~220 ns per cell
 * when every cell is its own str object (PyObject_Hash + PyObject_RichCompare per probe): 1.1 s for a 5 M-cell
 * column, more than every GPU stage together.  This routine walks the object pointers once: same pointer as the
 * previous cell -> same code; otherwise the str's cached hash (computed once per object by CPython) indexes a small
 * open-addressing table whose entries compare by pointer, then by (hash, kind, length, bytes).  Codes are assigned in
 * order of first appearance, exactly like pd.factorize(sort=False).
 *
 * Called through ctypes.PyDLL (the GIL is held; nothing here releases it).  Built by the Makefile next to
 * libpilot_b200.so; when the helper is missing the caller uses pandas.
 *
 * Returns the number of distinct labels, -1 if a cell is not an exact str (None / NaN / other types: the caller
 * falls back to pandas, which defines how those are treated), -2 if there are more than max_unique labels.
 * uniques[] receives BORROWED references to the first object of every label.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    PyObject *obj;   /* representative */
    Py_hash_t hash;
    int32_t code;
} slot_t;

static int same_str(PyObject *a, PyObject *b)
{
    if (PyUnicode_KIND(a) != PyUnicode_KIND(b)) return 0;
    const Py_ssize_t n = PyUnicode_GET_LENGTH(a);
    if (n != PyUnicode_GET_LENGTH(b)) return 0;
    return memcmp(PyUnicode_DATA(a), PyUnicode_DATA(b), (size_t)n * PyUnicode_KIND(a)) == 0;
}

int64_t pilot_factorize_str(PyObject **cells, int64_t n, int32_t *codes, PyObject **uniques, int64_t max_unique)
{
    size_t cap = 1024;
    while (cap < (size_t)max_unique * 2) cap <<= 1;
    slot_t *tab = (slot_t *)calloc(cap, sizeof(slot_t));
    if (!tab) return -3;
    const size_t mask = cap - 1;
    int64_t nuniq = 0;
    PyObject *prev = NULL;
    int32_t prev_code = -1;
    for (int64_t i = 0; i < n; ++i) {
        PyObject *o = cells[i];
        if (o == prev) { codes[i] = prev_code; continue; }
        if (!PyUnicode_CheckExact(o)) { free(tab); return -1; }
        Py_hash_t h = ((PyASCIIObject *)o)->hash;
        if (h == -1) {
            h = PyObject_Hash(o);
            if (h == -1) { PyErr_Clear(); free(tab); return -1; }
        }
        size_t p = ((size_t)h * 0x9E3779B97F4A7C15ull >> 20) & mask;
        for (;;) {
            slot_t *s = &tab[p];
            if (s->obj == NULL) {
                if (nuniq >= max_unique) { free(tab); return -2; }
                s->obj = o; s->hash = h; s->code = (int32_t)nuniq;
                uniques[nuniq++] = o;
                prev_code = s->code;
                break;
            }
            if (s->obj == o || (s->hash == h && same_str(s->obj, o))) { prev_code = s->code; break; }
            p = (p + 1) & mask;
        }
        codes[i] = prev_code;
        prev = o;
    }
    free(tab);
    return nuniq;
}
