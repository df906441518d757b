/* _b2rfast: CPython fast-call shims for the two calls an agent's loop makes per
 * environment step / per update (OutOfGraph*ReplayBuffer.add, ReplayTrainer.step).
 *
 * ctypes costs ~2 us per call plus ~1 us to get at a numpy array's address; at the
 * agent's batch of 32 the device step takes ~25 us, so four add() calls per update
 * through ctypes are most of the host time.  These shims take the observation through
 * the buffer protocol and forward to the same C-ABI entry points of libb200replay.so
 * (resolved with dlsym at bind time: this module has no link-time dependency on it).
 * No logic lives here: argument checks that the reference performs stay in Python.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <dlfcn.h>
#include <stdint.h>

typedef int (*add_atari_fn)(void *, const void *, int32_t, float, uint8_t, double, int,
                            void *);
typedef int (*step_host_fn)(void *, const float *, const float *, float *, int64_t *,
                            void *);

static add_atari_fn p_add_atari = NULL;
static step_host_fn p_step_host = NULL;

static PyObject *fast_bind(PyObject *self, PyObject *args) {
  const char *path;
  if (!PyArg_ParseTuple(args, "s", &path)) return NULL;
  void *lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    PyErr_Format(PyExc_OSError, "dlopen(%s): %s", path, dlerror());
    return NULL;
  }
  p_add_atari = (add_atari_fn)dlsym(lib, "b2r_add_atari");
  p_step_host = (step_host_fn)dlsym(lib, "b2r_trainer_step_host");
  if (!p_add_atari || !p_step_host) {
    PyErr_SetString(PyExc_OSError, "libb200replay.so lacks b2r_add_atari / "
                                   "b2r_trainer_step_host");
    return NULL;
  }
  Py_RETURN_NONE;
}

/* add_atari(handle, obs_bytes, observation, action, reward, terminal, priority, mode,
 *           stream) -> status, or -1 when the observation is not a C-contiguous
 * buffer of exactly obs_bytes bytes (the caller then takes the general path). */
static PyObject *fast_add_atari(PyObject *self, PyObject *const *args, Py_ssize_t n) {
  if (n != 9) {
    PyErr_SetString(PyExc_TypeError, "add_atari takes 9 arguments");
    return NULL;
  }
  void *handle = PyLong_AsVoidPtr(args[0]);
  const Py_ssize_t obs_bytes = PyLong_AsSsize_t(args[1]);
  const long action = PyLong_AsLong(args[3]);
  const double reward = PyFloat_AsDouble(args[4]);
  const long terminal = PyLong_AsLong(args[5]);
  const double priority = PyFloat_AsDouble(args[6]);
  const long mode = PyLong_AsLong(args[7]);
  void *stream = PyLong_AsVoidPtr(args[8]);
  if (PyErr_Occurred()) return NULL;
  Py_buffer view;
  if (PyObject_GetBuffer(args[2], &view, PyBUF_C_CONTIGUOUS) != 0) {
    PyErr_Clear();
    return PyLong_FromLong(-1);
  }
  int status = -1;
  if (view.len == obs_bytes && p_add_atari != NULL)
    status = p_add_atari(handle, view.buf, (int32_t)action, (float)reward,
                         (uint8_t)terminal, priority, (int)mode, stream);
  PyBuffer_Release(&view);
  return PyLong_FromLong(status);
}

/* trainer_step(handle, online_ptr, target_ptr, loss_ptr, step_ptr, stream) -> status */
static PyObject *fast_trainer_step(PyObject *self, PyObject *const *args, Py_ssize_t n) {
  if (n != 6) {
    PyErr_SetString(PyExc_TypeError, "trainer_step takes 6 arguments");
    return NULL;
  }
  void *p[6];
  for (int k = 0; k < 6; ++k) p[k] = PyLong_AsVoidPtr(args[k]);
  if (PyErr_Occurred()) return NULL;
  if (p_step_host == NULL) {
    PyErr_SetString(PyExc_RuntimeError, "_b2rfast is not bound");
    return NULL;
  }
  int status;
  Py_BEGIN_ALLOW_THREADS
  status = p_step_host(p[0], (const float *)p[1], (const float *)p[2], (float *)p[3],
                       (int64_t *)p[4], p[5]);
  Py_END_ALLOW_THREADS
  return PyLong_FromLong(status);
}

static PyMethodDef methods[] = {
    {"bind", fast_bind, METH_VARARGS, "bind(path to libb200replay.so)"},
    {"add_atari", (PyCFunction)(void (*)(void))fast_add_atari, METH_FASTCALL,
     "b2r_add_atari with the observation taken through the buffer protocol"},
    {"trainer_step", (PyCFunction)(void (*)(void))fast_trainer_step, METH_FASTCALL,
     "b2r_trainer_step_host from raw addresses"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_b2rfast",
                                    "fast-call shims for libb200replay", -1, methods};

PyMODINIT_FUNC PyInit__b2rfast(void) { return PyModule_Create(&module); }
