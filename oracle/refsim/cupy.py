"""TEST INFRASTRUCTURE ONLY — a minimal ``cupy`` stand-in backed by NumPy.

Lets the *unmodified* reference solver modules (``/root/reference/solver/*.py``,
which do ``import cupy as cp`` and launch ``numba.cuda`` kernels) execute on the
CPU under Numba's CUDA simulator (``NUMBA_ENABLE_CUDASIM=1``).  It is only put on
``sys.path`` by ``oracle/refsim/run_reference.py`` when golden fixtures are
(re)generated in the build container; nothing in the product imports it.

Only what the reference's hot path touches is provided: array creation returning
an ``ndarray`` subclass with ``.get()``, and NumPy's elementwise / reduction API.
"""
import numpy as _np
from numpy import *  # noqa: F401,F403  (sum, prod, abs, float64, int64, ...)

float64 = _np.float64
float32 = _np.float32
int64 = _np.int64
int32 = _np.int32
bool_ = _np.bool_


class ndarray(_np.ndarray):
    """NumPy array that also answers CuPy's ``.get()`` (device -> host copy)."""

    def get(self):
        return _np.asarray(self)


def _wrap(a):
    return _np.asarray(a).view(ndarray)


def array(obj, dtype=None, **kw):
    return _wrap(_np.array(obj, dtype=dtype, **kw))


def asarray(obj, dtype=None):
    return _wrap(_np.asarray(obj, dtype=dtype))


def zeros(shape, dtype=float64):
    return _wrap(_np.zeros(tuple(int(s) for s in _np.atleast_1d(shape)), dtype=dtype))


def ones(shape, dtype=float64):
    return _wrap(_np.ones(tuple(int(s) for s in _np.atleast_1d(shape)), dtype=dtype))


def empty(shape, dtype=float64):
    return zeros(shape, dtype)


def asnumpy(a):
    return _np.asarray(a)
