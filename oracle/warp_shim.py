"""oracle/warp_shim.py -- TEST INFRASTRUCTURE.  A tiny pure-Python stand-in for the `warp` module, just large
enough to EXECUTE the reference's own kernel sources unmodified, one simulated thread at a time, on the CPU.

Why: the reference's MPPI kernels are `@wp.kernel` Python functions (thesis_master/warp_implementation/
{sampling,projection,critics}_warp.py) driven by `MPPI_Controller` (MPPI_isaac.py).  warp-lang is absent here (no
network, no GPU in the build container), so the real thing cannot run -- but the kernel *source* is ordinary Python
once `wp.*` resolves to something.  With this module installed as `sys.modules["warp"]`, importing MPPI_isaac.py and
calling `MPPI_Controller.run()` interprets the reference's code line by line.  tests/golden/make_golden_warp.py
does exactly that (in the build container, where /root/reference exists) and commits the inputs and outputs as
fixtures; tests then hold BOTH the oracle and the CUDA path to those fixtures.  Nothing here is part of the product.

Semantics reproduced (Warp's documented scalar model):
  * kernel scalars are float32 / int32: `float` arguments and array elements are numpy.float32, and NumPy >= 2 keeps
    float32 when a float32 meets a Python literal (NEP 50 weak scalars) -- every +, -, *, / is one correctly rounded
    binary32 operation, evaluated in the source's own order.  No FMA contraction (NVRTC may contract on a GPU: that is
    what the 1e-4 tolerance of the specification absorbs).
  * `wp.int(x)` truncates toward zero (C cast); `wp.trunc`, `wp.sqrt`, `wp.abs`, `wp.clamp = min(max(x, lo), hi)`.
  * vec2f / vec3f / mat22f have value semantics (indexing an array yields a copy; assignment stores a copy);
    vec * scalar, vec / scalar are component-wise; dot = a0*b0 + a1*b1 + a2*b2 in that order; cross as usual.
  * `wp.launch(kernel, dim, inputs)` runs tid = 0..dim-1 sequentially.  A kernel whose source calls
    `wp.atomic_min` runs its threads concurrently with a rendezvous inside atomic_min, i.e. the schedule in which
    every thread's atomic lands before any thread reads the result back -- the race-free outcome the reference
    intends for `_compute_weights` (critics_warp.py:338-347; host version old_files/run_mppi.py:222-226).
  * `wp.randn(state)` is NOT Warp's PCG/Box-Muller (third-party arithmetic, unpinned): it is a deterministic standard
    normal keyed by the integer state (`randn_from_state`), so that the generator script can hand the same noise to
    the oracle -- the "shared injected noise" route of the specification.
"""
from __future__ import annotations

import inspect
import threading

import numpy as np

f32 = np.float32
float32 = np.float32
_tls = threading.local()
_rendezvous = None


def init():
    return None


def tid():
    return _tls.tid


# ------------------------------------------------------------------ scalar helpers
def float(x):                     # noqa: A001  (wp.float)
    return f32(x)


def int(x):                       # noqa: A001  (wp.int): C-style truncation
    import builtins
    return builtins.int(x)


def uint32(x):
    import builtins
    return builtins.int(x) & 0xFFFFFFFF


def sqrt(x):
    return np.sqrt(f32(x))


def sin(x):
    return np.sin(f32(x))


def cos(x):
    return np.cos(f32(x))


def exp(x):
    with np.errstate(over="ignore", under="ignore"):
        return np.exp(f32(x))


def atan(x):
    return np.arctan(f32(x))


def pow(x, y):                    # noqa: A001
    return np.power(f32(x), f32(y))


def trunc(x):
    return np.trunc(f32(x))


def abs(x):                       # noqa: A001
    return np.abs(x)


def min(a, b):                    # noqa: A001
    return a if a < b else b


def max(a, b):                    # noqa: A001
    return a if a > b else b


def clamp(x, lo, hi):
    x, lo, hi = f32(x), f32(lo), f32(hi)
    m = x if x > lo else lo       # max(x, lo)
    return m if m < hi else hi    # min(., hi)


def randn_from_state(state) -> np.float32:
    """Deterministic N(0,1) float32 for an integer RNG state (stands in for wp.randn; see module docstring)."""
    import builtins
    return f32(np.random.default_rng(builtins.int(state) + 0x9E3779B9).standard_normal())


def randn(state):
    return randn_from_state(state)


# ------------------------------------------------------------------ vector / matrix value types
class _Vec:
    N = 0
    __slots__ = ("c",)
    __array_ufunc__ = None            # numpy.float32 * vec must defer to vec.__rmul__

    def __init__(self, *a):
        if len(a) == 0:
            self.c = [f32(0)] * self.N
        elif len(a) == 1:
            self.c = [f32(v) for v in a[0]]
        else:
            self.c = [f32(v) for v in a]
        assert len(self.c) == self.N

    def __len__(self):
        return self.N

    def __iter__(self):
        return iter(self.c)

    def __getitem__(self, i):
        return self.c[i]

    def __setitem__(self, i, v):
        self.c[i] = f32(v)

    def copy(self):
        return type(self)(self.c)

    def __add__(self, o):
        return type(self)([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        return type(self)([a - b for a, b in zip(self.c, o.c)])

    def __neg__(self):
        return type(self)([-a for a in self.c])

    def __mul__(self, s):
        s = f32(s)
        return type(self)([a * s for a in self.c])

    def __rmul__(self, s):
        s = f32(s)
        return type(self)([s * a for a in self.c])

    def __truediv__(self, s):
        s = f32(s)
        return type(self)([a / s for a in self.c])

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(str(x) for x in self.c)})"


class vec2f(_Vec):
    N = 2
    __slots__ = ()


class vec3f(_Vec):
    N = 3
    __slots__ = ()


class mat22f:
    __slots__ = ("m",)
    __array_ufunc__ = None

    def __init__(self, *a):
        if len(a) == 0:
            self.m = np.zeros((2, 2), np.float32)
        elif len(a) == 1:
            self.m = np.array(a[0], np.float32).reshape(2, 2).copy()
        else:
            self.m = np.array(a, np.float32).reshape(2, 2)

    def __getitem__(self, ij):
        return self.m[ij[0], ij[1]]

    def __setitem__(self, ij, v):
        self.m[ij[0], ij[1]] = f32(v)

    def copy(self):
        return mat22f(self.m)


def dot(a, b):
    r = a.c[0] * b.c[0]
    for i in range(1, a.N):
        r = r + a.c[i] * b.c[i]
    return r


def cross(a, b):
    return vec3f(a.c[1] * b.c[2] - a.c[2] * b.c[1],
                 a.c[2] * b.c[0] - a.c[0] * b.c[2],
                 a.c[0] * b.c[1] - a.c[1] * b.c[0])


def length(a):
    return np.sqrt(dot(a, a))


# ------------------------------------------------------------------ arrays
_SHAPES = {vec2f: (2,), vec3f: (3,), mat22f: (2, 2)}


class _ArrayAnnotation:
    def __init__(self, dtype):
        self.dtype = dtype


class array:                       # noqa: N801  (wp.array)
    def __new__(cls, data=None, dtype=None, device=None, **kw):
        if data is None:
            return _ArrayAnnotation(dtype)          # `x: wp.array(dtype=float)` in a kernel signature
        return super().__new__(cls)

    def __init__(self, data=None, dtype=None, device=None, **kw):
        import builtins
        self.vtype = dtype if dtype in _SHAPES else None
        if self.vtype is not None:
            rows = [np.asarray(v.m if isinstance(v, mat22f) else (v.c if isinstance(v, _Vec) else v), np.float32)
                    for v in data]
            self.data = np.stack(rows).reshape((len(rows),) + _SHAPES[self.vtype]).copy()
        else:
            self.data = np.array(data, dtype=np.float32).reshape(-1).copy()
        self.shape = (builtins.int(self.data.shape[0]),)

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, i):
        if self.vtype is None:
            return self.data[i]                      # numpy.float32 scalar
        return self.vtype(self.data[i])             # value copy

    def __setitem__(self, i, v):
        if self.vtype is None:
            self.data[i] = f32(v)
        elif self.vtype is mat22f:
            self.data[i] = v.m
        else:
            self.data[i] = np.array(v.c, np.float32)

    def numpy(self):
        return self.data.copy()

    def zero_(self):
        self.data[...] = 0

    def assign(self, src):
        self.data[...] = np.asarray(src.data if isinstance(src, array) else src, np.float32).reshape(self.data.shape)

    @property
    def ptr(self):
        return self.data.ctypes.data


def zeros(n, dtype=None, device=None, **kw):
    import builtins
    n = builtins.int(n if not isinstance(n, tuple) else n[0])
    if dtype in _SHAPES:
        a = array.__new__(array, data=[])
        a.vtype = dtype
        a.data = np.zeros((n,) + _SHAPES[dtype], np.float32)
        a.shape = (n,)
        return a
    return array(np.zeros(n, np.float32), dtype=np.float32)


def atomic_min(arr, i, v):
    with _lock:
        if v < arr.data[i]:
            arr.data[i] = f32(v)
    if _rendezvous is not None:
        _rendezvous.wait()


def atomic_add(arr, i, v):
    with _lock:
        arr.data[i] = arr.data[i] + f32(v)


_lock = threading.Lock()


# ------------------------------------------------------------------ kernels
def func(f):
    return f


def kernel(f):
    f._wp_uses_atomic_min = "atomic_min" in inspect.getsource(f)
    return f


def _convert(fn, inputs):
    import builtins
    out = []
    params = list(inspect.signature(fn).parameters.values())
    assert len(params) == len(inputs), f"{fn.__name__}: expected {len(params)} arguments, got {len(inputs)}"
    for p, v in zip(params, inputs):
        ann = p.annotation
        if ann is builtins.float or ann is np.float32:
            v = f32(v)
        elif ann is builtins.int:
            v = builtins.int(v)
        out.append(v)
    return out


def launch(kernel=None, dim=None, inputs=(), device=None, **kw):   # noqa: A002
    global _rendezvous
    import builtins
    args = _convert(kernel, list(inputs))
    dim = builtins.int(dim)
    if getattr(kernel, "_wp_uses_atomic_min", False) and dim > 1:
        _rendezvous = threading.Barrier(dim)
        errs = []

        def body(t):
            _tls.tid = t
            try:
                kernel(*args)
            except BaseException as e:       # noqa: BLE001
                errs.append(e)
                _rendezvous.abort()
        ths = [threading.Thread(target=body, args=(t,)) for t in range(dim)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        _rendezvous = None
        if errs:
            raise errs[0]
        return
    for t in range(dim):
        _tls.tid = t
        kernel(*args)
