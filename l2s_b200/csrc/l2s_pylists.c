/*
 * l2s_b200 -- CPython glue for the list-shaped part of the reference's API.
 *
 * The reference hands feasible moves to the policy as nested Python lists (env/environment.py:411-423,
 * model/actor.py:216-226) and takes actions back as a list of [a0, a1] (env/environment.py:356).  At large
 * batch sizes building ~10 small lists per instance dominates a step of the drop-in, so the two conversions
 * are done here with the C API (called through ctypes.PyDLL, GIL held).  No computation happens here: the
 * inputs are the pinned host records that the step kernel wrote (l2s_host_records).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

/* Host records of l2s_step_host ([B][stride] uint32, see include/l2s_b200.h: word 0 = npairs | ..., words 4.. =
 * a0 | a1 << 16) -> list[B] of list[P] of [a0, a1]; an instance without moves gets [[0, 0]] exactly like
 * JsspN5.feasible_actions. */
PyObject* l2s_build_move_lists(const uint32_t* records, int stride, int B) {
  const int pmax = stride - 4;
  /* ~10 container objects per instance: keep the cyclic collector from re-scanning them while they are built */
  const int gc_was_enabled = PyGC_Disable();
  PyObject* outer = PyList_New(B);
  if (!outer) { if (gc_was_enabled) PyGC_Enable(); return NULL; }
  for (int b = 0; b < B; ++b) {
    const uint32_t* row = records + (size_t)b * stride + 4;
    int c = (int)(row[-4] & 0xffffu);
    if (c > pmax) c = pmax;
    int m = c > 0 ? c : 1;
    PyObject* inner = PyList_New(m);
    if (!inner) { Py_DECREF(outer); if (gc_was_enabled) PyGC_Enable(); return NULL; }
    for (int i = 0; i < m; ++i) {
      long a0 = c > 0 ? (long)(row[i] & 0xffffu) : 0, a1 = c > 0 ? (long)(row[i] >> 16) : 0;
      PyObject* pair = PyList_New(2);
      PyObject* x0 = PyLong_FromLong(a0);
      PyObject* x1 = PyLong_FromLong(a1);
      if (!pair || !x0 || !x1) {
        Py_XDECREF(pair); Py_XDECREF(x0); Py_XDECREF(x1); Py_DECREF(inner); Py_DECREF(outer);
        if (gc_was_enabled) PyGC_Enable();
        return NULL;
      }
      PyList_SET_ITEM(pair, 0, x0);
      PyList_SET_ITEM(pair, 1, x1);
      PyList_SET_ITEM(inner, i, pair);
    }
    PyList_SET_ITEM(outer, b, inner);
  }
  if (gc_was_enabled) PyGC_Enable();
  return outer;
}

/* actions: sequence of B items, each a sequence of two integers -> out [B, 2] int32.  Returns 0, or -1 with a
 * Python exception set. */
int l2s_parse_actions(PyObject* actions, int32_t* out, int B) {
  PyObject* fast = PySequence_Fast(actions, "actions must be a sequence of [a0, a1] pairs");
  if (!fast) return -1;
  if (PySequence_Fast_GET_SIZE(fast) != B) {
    Py_DECREF(fast);
    PyErr_SetString(PyExc_AssertionError, "one action per instance");
    return -1;
  }
  for (int b = 0; b < B; ++b) {
    PyObject* item = PySequence_Fast_GET_ITEM(fast, b);
    /* common case: an exact list of two exact ints -- no temporaries */
    if (PyList_CheckExact(item) && PyList_GET_SIZE(item) == 2 && PyLong_CheckExact(PyList_GET_ITEM(item, 0)) &&
        PyLong_CheckExact(PyList_GET_ITEM(item, 1))) {
      const long v0 = PyLong_AsLong(PyList_GET_ITEM(item, 0)), v1 = PyLong_AsLong(PyList_GET_ITEM(item, 1));
      if ((v0 == -1 || v1 == -1) && PyErr_Occurred()) { Py_DECREF(fast); return -1; }
      out[2 * b] = (int32_t)v0;
      out[2 * b + 1] = (int32_t)v1;
      continue;
    }
    PyObject* pf = PySequence_Fast(item, "every action must be a pair [a0, a1]");
    if (!pf) { Py_DECREF(fast); return -1; }
    if (PySequence_Fast_GET_SIZE(pf) != 2) {
      Py_DECREF(pf); Py_DECREF(fast);
      PyErr_SetString(PyExc_ValueError, "every action must be a pair [a0, a1]");
      return -1;
    }
    for (int k = 0; k < 2; ++k) {
      PyObject* num = PyNumber_Index(PySequence_Fast_GET_ITEM(pf, k));   /* accepts int and numpy integers */
      if (!num) { Py_DECREF(pf); Py_DECREF(fast); return -1; }
      long v = PyLong_AsLong(num);
      Py_DECREF(num);
      if (v == -1 && PyErr_Occurred()) { Py_DECREF(pf); Py_DECREF(fast); return -1; }
      out[2 * b + k] = (int32_t)v;
    }
    Py_DECREF(pf);
  }
  Py_DECREF(fast);
  return 0;
}
