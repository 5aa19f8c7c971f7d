"""A tiny pure-Python/NumPy stand-in for ``warp-lang`` -- TEST TOOLING ONLY.

Purpose: ``warp-lang==1.7.0`` (the reference's only runtime) is not installable in this image, so
the reference's modules cannot be imported.  This shim implements just enough of the ``wp`` API
for the UNMODIFIED reference sources (/root/reference/{forward,backward,optimizer,loss}.py, ...)
to import and for their ``@wp.kernel`` functions to execute, one thread at a time, as ordinary
Python.  tests/golden/make_golden.py uses it to turn the reference's own kernel source into golden
vectors for the CPU oracle.

What it pins: the reference's control flow, indexing, constants, argument plumbing and host
orchestration.  What it cannot pin: the bit-level behaviour of Warp built-ins (matrix products,
quat_to_matrix, normalize, randf, exp) -- those are restated here from upstream knowledge of Warp's
native headers and are marked [Warp] in the oracle.

Scalar semantics: kernel ``float`` arguments and array elements are ``numpy.float32``; Python float
literals are "weak" under NumPy 2 promotion, so arithmetic stays in binary32 like Warp's codegen.
Threads run serially in launch-index order (last dimension fastest), like Warp's CPU device.
"""
from __future__ import annotations

import inspect
import itertools
import types as _pytypes

import numpy as np

f32 = np.float32

# ------------------------------------------------------------------------------------------------
# scalar types
# ------------------------------------------------------------------------------------------------
float32 = np.float32
float16 = np.float16
float64 = np.float64
int32 = np.int32
int64 = np.int64
uint32 = np.uint32
uint64 = np.uint64
int8 = np.int8
uint8 = np.uint8

_SCALAR_NP = {float: np.float32, int: np.int32, bool: np.bool_, np.float32: np.float32, np.int32: np.int32,
              np.int64: np.int64, np.uint32: np.uint32, np.float16: np.float16}


def _is_scalar(x):
    return isinstance(x, (int, float, bool, np.generic))


def _s(x):
    """Python/NumPy scalar -> float32 scalar (ints stay exact up to 2^24 like Warp's float())."""
    return x if isinstance(x, np.float32) else np.float32(x)


# ------------------------------------------------------------------------------------------------
# vectors / matrices
# ------------------------------------------------------------------------------------------------
class _VecBase:
    """Fixed-length vector backed by a NumPy view (so ``arr[i][k] = v`` writes through)."""
    _length = 0
    _np = np.float32
    __array_ufunc__ = None
    __slots__ = ("d",)

    def __init__(self, *args):
        n = self._length
        if len(args) == 0:
            self.d = np.zeros(n, self._np)
        elif len(args) == 1 and isinstance(args[0], _VecBase):
            self.d = args[0].d.astype(self._np)
        elif len(args) == 1 and isinstance(args[0], (np.ndarray, list, tuple)):
            self.d = np.asarray(args[0]).astype(self._np).reshape(n).copy()
        elif len(args) == 1 and n > 1:
            self.d = np.full(n, args[0], self._np)
        else:
            assert len(args) == n, f"{type(self).__name__} needs {n} components"
            self.d = np.array([self._np(a) for a in args], self._np)

    @classmethod
    def _view(cls, arr):
        v = cls.__new__(cls)
        v.d = arr
        return v

    def __len__(self):
        return self._length

    def __getitem__(self, i):
        return self.d[i]

    def __setitem__(self, i, v):
        self.d[i] = v

    def __iter__(self):
        return iter(self.d)

    def _new(self, data):
        return type(self)._view(np.asarray(data, self._np))

    def _bin(self, o, op):
        if isinstance(o, _VecBase):
            assert o._length == self._length
            return self._new([op(a, b) for a, b in zip(self.d, o.d)])
        return NotImplemented

    def __add__(self, o):
        return self._bin(o, lambda a, b: a + b)

    def __sub__(self, o):
        return self._bin(o, lambda a, b: a - b)

    def __neg__(self):
        return self._new([-a for a in self.d])

    def __mul__(self, o):
        if _is_scalar(o):
            s = self._np(o)
            return self._new([a * s for a in self.d])
        if isinstance(o, _MatBase):          # [Warp] row-vector * matrix
            return o._rvecmul(self)
        return NotImplemented

    def __rmul__(self, o):
        if _is_scalar(o):
            s = self._np(o)
            return self._new([s * a for a in self.d])
        return NotImplemented

    def __truediv__(self, o):
        if _is_scalar(o):
            s = self._np(o)
            return self._new([a / s for a in self.d])
        return NotImplemented

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(repr(float(a)) for a in self.d)})"


def _make_vec(name, length, np_dtype):
    return type(name, (_VecBase,), {"_length": length, "_np": np_dtype, "__slots__": ()})


vec2 = _make_vec("vec2", 2, np.float32)
vec3 = _make_vec("vec3", 3, np.float32)
vec4 = _make_vec("vec4", 4, np.float32)
vec2i = _make_vec("vec2i", 2, np.int32)
vec3i = _make_vec("vec3i", 3, np.int32)
vec2h = _make_vec("vec2h", 2, np.float16)
vec3f = vec3
quat = _make_vec("quat", 4, np.float32)


def quaternion(x, y, z, w):
    return quat(x, y, z, w)


class _MatBase:
    _rows = 0
    _cols = 0
    __array_ufunc__ = None
    __slots__ = ("d",)

    def __init__(self, *args):
        r, c = self._rows, self._cols
        if len(args) == 0:
            self.d = np.zeros((r, c), np.float32)
        elif len(args) == 1 and isinstance(args[0], (np.ndarray, list, tuple)):
            self.d = np.asarray(args[0], dtype=np.float32).reshape(r, c).copy()   # row-major
        elif len(args) == 1 and isinstance(args[0], _MatBase):
            self.d = args[0].d.copy()
        else:
            assert len(args) == r * c
            self.d = np.array([np.float32(a) for a in args], np.float32).reshape(r, c)

    @classmethod
    def _from(cls, d):
        m = cls.__new__(cls)
        m.d = d
        return m

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            return self.d[idx[0], idx[1]]
        rowcls = {2: vec2, 3: vec3, 4: vec4}[self._cols]
        return rowcls._view(self.d[idx])          # m[i] = row i

    def __setitem__(self, idx, v):
        if isinstance(idx, tuple):
            self.d[idx[0], idx[1]] = v
        else:
            self.d[idx] = v.d if isinstance(v, _VecBase) else v

    def __mul__(self, o):
        if isinstance(o, _MatBase):               # [Warp] t[i][j] = 0; t[i][j] += a[i][k]*b[k][j]
            assert self._cols == o._rows
            out = np.zeros((self._rows, o._cols), np.float32)
            for i in range(self._rows):
                for j in range(o._cols):
                    s = np.float32(0.0)
                    for k in range(self._cols):
                        s = s + self.d[i, k] * o.d[k, j]
                    out[i, j] = s
            return _mat_cls(self._rows, o._cols)._from(out)
        if isinstance(o, _VecBase):               # [Warp] r = col(0)*v[0]; r += col(i)*v[i]
            assert self._cols == o._length
            res = []
            for i in range(self._rows):
                s = self.d[i, 0] * o.d[0]
                for k in range(1, self._cols):
                    s = s + self.d[i, k] * o.d[k]
                res.append(s)
            return {2: vec2, 3: vec3, 4: vec4}[self._rows](*res)
        if _is_scalar(o):
            return type(self)._from((self.d * np.float32(o)).astype(np.float32))
        return NotImplemented

    def __rmul__(self, o):
        if _is_scalar(o):
            return type(self)._from((np.float32(o) * self.d).astype(np.float32))
        return NotImplemented

    def _rvecmul(self, v):                         # [Warp] r = row(0)*v[0]; r += row(i)*v[i]
        assert self._rows == v._length
        res = []
        for j in range(self._cols):
            s = self.d[0, j] * v.d[0]
            for i in range(1, self._rows):
                s = s + self.d[i, j] * v.d[i]
            res.append(s)
        return {2: vec2, 3: vec3, 4: vec4}[self._cols](*res)

    def __add__(self, o):
        return type(self)._from(self.d + o.d)

    def __sub__(self, o):
        return type(self)._from(self.d - o.d)


_MAT_CACHE = {}


def _mat_cls(r, c):
    if (r, c) not in _MAT_CACHE:
        _MAT_CACHE[(r, c)] = type(f"mat{r}{c}", (_MatBase,), {"_rows": r, "_cols": c, "__slots__": ()})
    return _MAT_CACHE[(r, c)]


mat22 = _mat_cls(2, 2)
mat33 = _mat_cls(3, 3)
mat44 = _mat_cls(4, 4)


def transpose(m):
    return _mat_cls(m._cols, m._rows)._from(np.ascontiguousarray(m.d.T))


def quat_rotate(q, v):
    """[Warp] native/quat.h quat_rotate."""
    x, y, z, w = q.d
    two = np.float32(2.0)
    c = two * w * w - np.float32(1.0)
    d = two * (x * v[0] + y * v[1] + z * v[2])
    return vec3(v[0] * c + x * d + (y * v[2] - z * v[1]) * w * two,
                v[1] * c + y * d + (z * v[0] - x * v[2]) * w * two,
                v[2] * c + z * d + (x * v[1] - y * v[0]) * w * two)


def quat_to_matrix(q):
    """[Warp] native/quat.h quat_to_matrix: columns are the rotated basis vectors."""
    c1 = quat_rotate(q, vec3(1.0, 0.0, 0.0))
    c2 = quat_rotate(q, vec3(0.0, 1.0, 0.0))
    c3 = quat_rotate(q, vec3(0.0, 0.0, 1.0))
    return mat33(c1[0], c2[0], c3[0], c1[1], c2[1], c3[1], c1[2], c2[2], c3[2])


# ------------------------------------------------------------------------------------------------
# built-in functions
# ------------------------------------------------------------------------------------------------
def dot(a, b):
    s = a.d[0] * b.d[0]
    for k in range(1, a._length):
        s = s + a.d[k] * b.d[k]
    return s


def length(a):
    return np.sqrt(dot(a, a))


def normalize(a):
    l = length(a)
    if l > np.float32(0.0):
        return a / l
    return type(a)()


def sqrt(x):
    return np.sqrt(_s(x))


def exp(x):
    return np.exp(_s(x))


def ceil(x):
    return np.ceil(_s(x))


def floor(x):
    return np.floor(_s(x))


def pow(a, b):  # noqa: A001
    return np.float32(np.power(_s(a), _s(b)))


def sign(x):
    """[Warp] sign(x) = x < 0 ? -1 : 1  (so sign(0) = +1)."""
    return np.float32(-1.0) if x < 0 else np.float32(1.0)


def abs(x):  # noqa: A001
    if isinstance(x, _VecBase):
        return x._new([np.abs(a) for a in x.d])
    return np.abs(x)


def min(a, b):  # noqa: A001
    return a if a < b else b


def max(a, b):  # noqa: A001
    return a if a > b else b


def clamp(x, lo, hi):
    return min(max(x, lo), hi)


def _pcg(state):
    state = int(state) & 0xFFFFFFFF
    b = (state * 747796405 + 2891336453) & 0xFFFFFFFF
    c = ((((b >> ((b >> 28) + 4)) ^ b) & 0xFFFFFFFF) * 277803737) & 0xFFFFFFFF
    return ((c >> 22) ^ c) & 0xFFFFFFFF


def randf(state):
    """[Warp] native/rand.h randf(uint32): PCG hash, top 24 bits / 2^24."""
    return np.float32(_pcg(state) >> 8) * np.float32(1.0 / 16777216.0)


def atomic_add(arr, idx, value):
    old = arr[idx]
    arr[idx] = old + value
    return old


def constant(x):
    return x


def init():
    return None


# ------------------------------------------------------------------------------------------------
# arrays
# ------------------------------------------------------------------------------------------------
def _veclen(dtype):
    if isinstance(dtype, type) and issubclass(dtype, _VecBase):
        return dtype._length
    return 0


def _np_dtype(dtype):
    if isinstance(dtype, type) and issubclass(dtype, _VecBase):
        return dtype._np
    return _SCALAR_NP.get(dtype, dtype)


# train.py:479-492 reads a stale, too-short ``avg_grads`` after cloning: an
# out-of-bounds READ whose value is undefined on a real device.  make_golden.py sets this flag for
# the densify case so that such reads yield 0 (the convention the oracle and the CUDA path adopt).
OOB_SCALAR_READ_IS_ZERO = False
# clone/compact write one element past the end when the last flag is set (quirk G5: the count is
# the last entry of an EXCLUSIVE scan).  With this flag such writes are dropped and logged.
OOB_WRITE_IS_DROPPED = False
OOB_WRITES = []


class array:  # noqa: N801
    """``wp.array``.  Called with only ``dtype=`` it is a type annotation stub."""

    def __init__(self, data=None, dtype=None, shape=None, device=None, ndim=1, **_):
        self.dtype = dtype
        self.device = device
        self._vl = _veclen(dtype)
        if data is None:
            self.data = None
            self.shape = (0,) * ndim
            return
        a = np.array(data, dtype=_np_dtype(dtype))      # copy, like wp.array(host data)
        if self._vl:
            if a.ndim >= 1 and a.shape[-1] == self._vl and a.ndim >= 2:
                pass
            else:
                a = a.reshape(-1, self._vl)
            self.shape = a.shape[:-1]
        else:
            self.shape = a.shape
        self.data = np.ascontiguousarray(a)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape))

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        if self._vl:
            return self.dtype._view(self.data[idx])
        if OOB_SCALAR_READ_IS_ZERO and self.ndim == 1 and isinstance(idx, (int, np.integer)) \
                and idx >= self.data.shape[0]:
            return self.data.dtype.type(0)      # reference reads past the end here (undefined on a GPU)
        return self.data[idx]

    def __setitem__(self, idx, v):
        if OOB_WRITE_IS_DROPPED and isinstance(idx, (int, np.integer)) and idx >= self.data.shape[0]:
            OOB_WRITES.append((int(idx), int(self.data.shape[0])))
            return                              # reference writes past the end here (quirk G5)
        if self._vl:
            self.data[idx] = v.d if isinstance(v, _VecBase) else v
        else:
            self.data[idx] = v

    def numpy(self):
        return self.data.copy()

    def reshape(self, *shape):
        raise NotImplementedError("shim: wp.array.reshape")


def array2d(dtype=None, **kw):
    return array(dtype=dtype, ndim=2, **kw)


def zeros(shape, dtype=float, device=None, **_):
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    shape = tuple(int(s) for s in shape)
    a = array(dtype=dtype, device=device)
    vl = _veclen(dtype)
    a.shape = shape
    a.data = np.zeros(shape + ((vl,) if vl else ()), _np_dtype(dtype))
    return a


def zeros_like(a):
    return zeros(a.shape, dtype=a.dtype, device=a.device)


def empty(shape, dtype=float, device=None, **_):
    return zeros(shape, dtype=dtype, device=device)


def copy(dest, src, dest_offset=0, src_offset=0, count=0):
    n = count if count else builtins_min(dest.data.shape[0] - dest_offset, src.data.shape[0] - src_offset)
    dest.data[dest_offset:dest_offset + n] = src.data[src_offset:src_offset + n]


def builtins_min(a, b):
    return a if a < b else b


def to_torch(a):
    import torch
    return torch.from_numpy(a.data)


def synchronize():
    return None


# ------------------------------------------------------------------------------------------------
# kernels / launch
# ------------------------------------------------------------------------------------------------
_TID = [None]


def tid():
    return _TID[0]


class Kernel:
    def __init__(self, fn):
        self.fn = fn
        self.params = list(inspect.signature(fn).parameters.values())
        self.__name__ = fn.__name__

    def convert(self, args):
        out = []
        assert len(args) == len(self.params), f"{self.__name__}: got {len(args)} args, want {len(self.params)}"
        for p, a in zip(self.params, args):
            ann = p.annotation
            if ann is float or ann is np.float32:
                a = np.float32(a)
            elif ann is int or ann is np.int32:
                a = int(a)
            elif ann is bool:
                a = bool(a)
            elif isinstance(ann, type) and issubclass(ann, _VecBase) and not isinstance(a, _VecBase):
                a = ann(*a)
            out.append(a)
        return out


def kernel(fn):
    return Kernel(fn)


def func(fn):
    return fn


def func_native(snippet):
    def deco(fn):
        if "reinterpret_cast<uint32_t&>" in snippet:
            return lambda x: np.float32(x).view(np.uint32)
        raise NotImplementedError("shim: unknown native snippet")
    return deco


def launch(kernel=None, dim=None, inputs=(), outputs=(), device=None, **_):  # noqa: A002
    k = kernel
    args = k.convert(list(inputs) + list(outputs))
    if isinstance(dim, (int, np.integer)):
        dim = (int(dim),)
    dim = tuple(int(d) for d in dim)
    if len(dim) == 1:
        for i in range(dim[0]):
            _TID[0] = i
            k.fn(*args)
    else:
        for idx in itertools.product(*[range(d) for d in dim]):
            _TID[0] = idx
            k.fn(*args)
    _TID[0] = None


# ------------------------------------------------------------------------------------------------
# wp.types / wp.utils
# ------------------------------------------------------------------------------------------------
def _vector(length, dtype):  # noqa: A002
    return _make_vec(f"vec{length}_{np.dtype(_np_dtype(dtype)).name}", length, _np_dtype(dtype))


types = _pytypes.SimpleNamespace(vector=_vector, matrix=lambda shape, dtype: _mat_cls(*shape))


def _radix_sort_pairs(keys, values, count):
    """[Warp] wp.utils.radix_sort_pairs: ascending, stable, in place on the first ``count``."""
    count = int(count)
    order = np.argsort(keys.data[:count], kind="stable")
    keys.data[:count] = keys.data[:count][order]
    values.data[:count] = values.data[:count][order]


def _array_scan(in_array, out_array, inclusive=True):
    c = np.cumsum(in_array.data, dtype=in_array.data.dtype)
    if inclusive:
        out_array.data[:] = c
    else:
        out_array.data[0] = 0
        out_array.data[1:] = c[:-1]


utils = _pytypes.SimpleNamespace(radix_sort_pairs=_radix_sort_pairs, array_scan=_array_scan)
