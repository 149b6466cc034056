/* _upb_pyptr: pointer table of a list of reference-layout states, built with the CPython buffer protocol.
 *
 * `pack_states` needs 9 raw pointers per rollout state (include/upb200.h, upb_pack_fill).  Collecting them in
 * Python costs ~1.5 us per array (2,304 arrays per 256-state minibatch); this helper walks the lists in C
 * and checks item type / contiguity on the way: numpy arrays are read straight from their object header (one cache
 * line per array; the objects of a large rollout buffer are cold in memory, so touching less matters), anything else
 * goes through the buffer protocol.  Host glue only: no CUDA.
 *
 *   pointer_table(states, out, n_cap, e_cap) -> -1 on success, or the index of the first state that needs the slow path
 *     states : list/tuple of list/tuple of 9 objects exporting C-contiguous buffers
 *     out    : writable buffer of 9 * len(states) uint64
 *     n_cap, e_cap : padded widths the packer will read; an array with a different element count raises ValueError
 *                    (the packer trusts the pointers: a short array would be an out-of-bounds host read)
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>

static const int kTypeNum[9] = {NPY_FLOAT32, NPY_FLOAT32, NPY_INT64, NPY_FLOAT32, NPY_BOOL, NPY_BOOL, NPY_BOOL, NPY_BOOL,
                                NPY_FLOAT32};

static const Py_ssize_t kItem[9] = {4, 4, 8, 4, 1, 1, 1, 1, 4};   /* f32 f32 i64 f32 bool bool bool bool f32 */
static const char kKind[9] = {'f', 'f', 'i', 'f', '?', '?', '?', '?', 'f'};

static int kind_ok(const char* fmt, char want) {
  if (!fmt) return 0;
  while (*fmt == '<' || *fmt == '=' || *fmt == '@' || *fmt == '|') ++fmt;
  switch (want) {
    case 'f': return *fmt == 'f';
    case 'i': return *fmt == 'q' || *fmt == 'l';
    default: return *fmt == '?';
  }
}

static int size_ok(int j, Py_ssize_t size, const Py_ssize_t want[9]) {
  return j == 8 ? size >= 2 : size == want[j];     /* stage: 3 in the reference, the packer reads entries 0 and 1 */
}

static PyObject* pointer_table(PyObject* self, PyObject* args) {
  PyObject* states;
  Py_buffer out;
  long n_cap = 0, e_cap = 0;
  if (!PyArg_ParseTuple(args, "Ow*ll", &states, &out, &n_cap, &e_cap)) return NULL;
  const Py_ssize_t want[9] = {52, (Py_ssize_t)n_cap * 23, (Py_ssize_t)e_cap * 2, 23, n_cap, e_cap, e_cap, n_cap, 3};
  long bad_size_state = -1;
  int bad_size_arr = -1;
  PyObject* seq = PySequence_Fast(states, "states must be a sequence");
  if (!seq) { PyBuffer_Release(&out); return NULL; }
  const Py_ssize_t count = PySequence_Fast_GET_SIZE(seq);
  long bad = -1;
  if (out.len < (Py_ssize_t)(9 * count * sizeof(uint64_t))) {
    PyBuffer_Release(&out);
    Py_DECREF(seq);
    PyErr_SetString(PyExc_ValueError, "pointer_table: output buffer too small");
    return NULL;
  }
  uint64_t* dst = (uint64_t*)out.buf;
  for (Py_ssize_t i = 0; i < count && bad < 0; ++i) {
    PyObject* st = PySequence_Fast_GET_ITEM(seq, i);
    if (!(PyList_Check(st) || PyTuple_Check(st)) || PySequence_Fast_GET_SIZE(st) != 9) { bad = (long)i; break; }
    for (int j = 0; j < 9; ++j) {
      PyObject* a = PySequence_Fast_GET_ITEM(st, j);
      if (PyArray_CheckExact(a)) {
        PyArrayObject* arr = (PyArrayObject*)a;
        if (!size_ok(j, (Py_ssize_t)PyArray_SIZE(arr), want)) { bad_size_state = (long)i; bad_size_arr = j; bad = (long)i; break; }
        if (PyArray_TYPE(arr) == kTypeNum[j] && PyArray_IS_C_CONTIGUOUS(arr) && PyArray_ISNOTSWAPPED(arr)) {
          dst[9 * i + j] = (uint64_t)(uintptr_t)PyArray_DATA(arr);
          continue;
        }
        bad = (long)i;
        break;
      }
      Py_buffer v;
      if (PyObject_GetBuffer(a, &v, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0) { PyErr_Clear(); bad = (long)i; break; }
      const int ok = v.itemsize == kItem[j] && kind_ok(v.format, kKind[j]);
      const int sz_ok = v.itemsize > 0 && size_ok(j, v.len / v.itemsize, want);
      dst[9 * i + j] = (uint64_t)(uintptr_t)v.buf;
      PyBuffer_Release(&v);         /* the caller's list keeps the array alive */
      if (!sz_ok) { bad_size_state = (long)i; bad_size_arr = j; bad = (long)i; break; }
      if (!ok) { bad = (long)i; break; }
    }
  }
  PyBuffer_Release(&out);
  Py_DECREF(seq);
  if (bad_size_state >= 0) {
    PyErr_Format(PyExc_ValueError,
                 "state %ld, array %d: element count does not match the padded widths (n_cap=%ld, e_cap=%ld): expected "
                 "52, n_cap*23, e_cap*2, 23, n_cap, e_cap, e_cap, n_cap, 3",
                 bad_size_state, bad_size_arr, n_cap, e_cap);
    return NULL;
  }
  return PyLong_FromLong(bad);
}

static PyMethodDef kMethods[] = {{"pointer_table", pointer_table, METH_VARARGS, "raw pointers of 9-array states"},
                                 {NULL, NULL, 0, NULL}};
static struct PyModuleDef kModule = {PyModuleDef_HEAD_INIT, "_upb_pyptr", NULL, -1, kMethods};
PyMODINIT_FUNC PyInit__upb_pyptr(void) {
  import_array();
  return PyModule_Create(&kModule);
}
