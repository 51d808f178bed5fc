"""Minimal TensorFlow-2.10-API stand-in over torch (CPU).  TEST INFRASTRUCTURE ONLY -- see ../README.md."""
import types

import numpy as np
import torch

newaxis = None


class DType:
    def __init__(self, name, torch_dtype):
        self.name, self.torch = name, torch_dtype

    def __repr__(self):
        return f"tf.{self.name}"


float32 = DType("float32", torch.float32)
float64 = DType("float64", torch.float64)
int32 = DType("int32", torch.int32)
int64 = DType("int64", torch.int64)
bool_ = DType("bool", torch.bool)
string = DType("string", None)
_BY_TORCH = {torch.float32: float32, torch.float64: float64, torch.int32: int32, torch.int64: int64, torch.bool: bool_}


def _td(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, DType):
        return dtype.torch
    return dtype


class TensorShape(tuple):
    def __eq__(self, other):
        return list(self) == list(other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = tuple.__hash__

    def __getitem__(self, item):
        r = tuple.__getitem__(self, item)
        return TensorShape(r) if isinstance(item, slice) else r


def _raw(x):
    """Tensor / Variable -> torch tensor; everything else unchanged."""
    return x.t if isinstance(x, Tensor) else x


def _idx(i):
    if isinstance(i, Tensor):
        return i.t.long() if i.t.dtype in (torch.int32, torch.int64) else i.t
    if isinstance(i, tuple):
        return tuple(_idx(j) for j in i)
    return i


def _as_int(v):
    if isinstance(v, Tensor):
        return int(v.t.item())
    return int(v)


def _ints(seq):
    return [_as_int(v) for v in seq]


class Tensor:
    """Eager tensor.  Binary operators convert the other operand to THIS tensor's dtype, like TensorFlow's
    operator overloads do (ops.convert_to_tensor(y, dtype_hint=x.dtype.base_dtype))."""
    __array_priority__ = 100

    def __init__(self, t):
        self.t = t

    # -- introspection
    @property
    def shape(self):
        return TensorShape(self.t.shape)

    @property
    def dtype(self):
        return _BY_TORCH[self.t.dtype]

    def numpy(self):
        return self.t.detach().numpy()

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        for i in range(self.t.shape[0]):
            yield Tensor(self.t[i])

    def __float__(self):
        return float(self.t.item())

    def __int__(self):
        return int(self.t.item())

    def __index__(self):
        return int(self.t.item())

    def __bool__(self):
        return bool(self.t.item())

    def __repr__(self):
        return f"tf_shim.Tensor({self.t!r})"

    def __array__(self, dtype=None, copy=None):
        a = self.t.detach().numpy()
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, item):
        return Tensor(self.t[_idx(item)])

    # -- arithmetic
    def _conv(self, other):
        if isinstance(other, Tensor):
            if other.t.dtype != self.t.dtype:
                raise TypeError(f"dtype mismatch {self.t.dtype} vs {other.t.dtype} (TensorFlow would raise InvalidArgumentError)")
            return other.t
        return torch.as_tensor(np.asarray(other)).to(self.t.dtype)

    def __add__(self, o): return Tensor(self.t + self._conv(o))
    def __radd__(self, o): return Tensor(self._conv(o) + self.t)
    def __sub__(self, o): return Tensor(self.t - self._conv(o))
    def __rsub__(self, o): return Tensor(self._conv(o) - self.t)
    def __mul__(self, o): return Tensor(self.t * self._conv(o))
    def __rmul__(self, o): return Tensor(self._conv(o) * self.t)
    def __truediv__(self, o): return Tensor(self.t / self._conv(o))
    def __rtruediv__(self, o): return Tensor(self._conv(o) / self.t)
    def __pow__(self, o): return Tensor(self.t ** self._conv(o))
    def __neg__(self): return Tensor(-self.t)
    def __matmul__(self, o): return Tensor(self.t @ self._conv(o))
    def __rmatmul__(self, o): return Tensor(self._conv(o) @ self.t)
    def __lt__(self, o): return Tensor(self.t < self._conv(o))
    def __le__(self, o): return Tensor(self.t <= self._conv(o))
    def __gt__(self, o): return Tensor(self.t > self._conv(o))
    def __ge__(self, o): return Tensor(self.t >= self._conv(o))
    def __eq__(self, o): return Tensor(self.t == self._conv(o))
    def __ne__(self, o): return Tensor(self.t != self._conv(o))
    __hash__ = object.__hash__


class StringTensor:
    """tf.string tensor: only construction, indexing with newaxis and == are needed."""
    def __init__(self, a):
        self.a = np.asarray(a, dtype=object)

    def __getitem__(self, item):
        return StringTensor(self.a[item])

    def __eq__(self, other):
        other = other.a if isinstance(other, StringTensor) else other
        return Tensor(torch.as_tensor(np.asarray(self.a == other, dtype=bool)))

    __hash__ = object.__hash__

    def value(self):
        return str(self.a.item()) if self.a.ndim == 0 else self.a

    def numpy(self):
        return self.a.item().encode() if self.a.ndim == 0 else self.a

    @property
    def shape(self):
        return TensorShape(self.a.shape)


class Variable(Tensor):
    def __init__(self, initial_value, trainable=True, dtype=None, shape=None, name=None):
        super().__init__(_to_torch(initial_value, dtype).clone())

    def assign(self, value):
        self.t = _to_torch(value, _BY_TORCH[self.t.dtype]).clone()
        return self


def _to_torch(value, dtype=None):
    td = _td(dtype)
    if isinstance(value, Tensor):
        return value.t if td is None or td == value.t.dtype else value.t.to(td)
    if isinstance(value, torch.Tensor):
        return value if td is None else value.to(td)
    if isinstance(value, (list, tuple)) and len(value) > 0 and any(isinstance(v, Tensor) for v in value):
        return torch.stack([_to_torch(v, dtype) for v in value])
    arr = np.asarray(value)
    if td is None:                                # TensorFlow's defaults: python floats -> float32, ints -> int32
        if arr.dtype == np.float64 and not isinstance(value, np.ndarray):
            td = torch.float32
        elif arr.dtype == np.int64 and not isinstance(value, np.ndarray):
            td = torch.int32
    t = torch.as_tensor(arr)
    return t if td is None else t.to(td)


def constant(value, dtype=None, shape=None, name=None):
    if dtype is string or isinstance(value, (str, bytes)) or (
            isinstance(value, (list, tuple, np.ndarray)) and np.asarray(value).dtype.kind in "USO"):
        return StringTensor(value)
    return Tensor(_to_torch(value, dtype))


def convert_to_tensor(value, dtype=None, name=None):
    return constant(value, dtype)


def cast(x, dtype):
    return Tensor(_to_torch(x).to(_td(dtype)))


class TensorSpec:
    def __init__(self, shape=None, dtype=None, name=None):
        self.shape, self.dtype = shape, dtype


def function(func=None, input_signature=None, **kwargs):
    """@tf.function / @tf.function(input_signature=...): eager execution, signature ignored."""
    if func is not None and callable(func):
        return func
    return lambda f: f


# ---------------------------------------------------------------------------------------- shapes
def shape(x):
    return Tensor(torch.tensor(list(_to_torch(x).shape), dtype=torch.int32))


def rank(x):
    return Tensor(torch.tensor(_to_torch(x).dim(), dtype=torch.int32))


def reshape(x, shape, name=None):
    return Tensor(_to_torch(x).reshape(_ints(shape)))


def squeeze(x, axis=None):
    t = _to_torch(x)
    return Tensor(t.squeeze() if axis is None else t.squeeze(axis))


def expand_dims(x, axis):
    return Tensor(_to_torch(x).unsqueeze(axis))


def stack(values, axis=0, name=None):
    ts = [_to_torch(v) for v in values]
    dt = next((t.dtype for t in ts if t.dtype.is_floating_point), ts[0].dtype)
    return Tensor(torch.stack([t.to(dt) for t in ts], dim=axis))


def unstack(x, axis=0):
    return [Tensor(t) for t in torch.unbind(_to_torch(x), dim=axis)]


def concat(values, axis, name=None):
    ts = [_to_torch(v) for v in values]
    dt = next((t.dtype for t in ts if t.dtype.is_floating_point), ts[0].dtype)
    return Tensor(torch.cat([t.to(dt) for t in ts], dim=axis))


def zeros(shape, dtype=float32):
    return Tensor(torch.zeros(_ints(shape), dtype=_td(dtype)))


def ones(shape, dtype=float32):
    return Tensor(torch.ones(_ints(shape), dtype=_td(dtype)))


def zeros_like(x):
    return Tensor(torch.zeros_like(_to_torch(x)))


def ones_like(x):
    return Tensor(torch.ones_like(_to_torch(x)))


def eye(num_rows, num_columns=None, batch_shape=None, dtype=float32, name=None):
    n = _as_int(num_rows)
    m = n if num_columns is None else _as_int(num_columns)
    e = torch.eye(n, m, dtype=_td(dtype))
    if batch_shape is not None:
        bs = _ints(batch_shape)
        e = e.expand(*bs, n, m).clone()
    return Tensor(e)


def broadcast_to(x, shape):
    return Tensor(_to_torch(x).broadcast_to(_ints(shape)))


def repeat(x, repeats, axis=None):
    return Tensor(torch.repeat_interleave(_to_torch(x), _as_int(repeats), dim=axis))


def gather(params, indices, axis=0, name=None):
    p, i = _to_torch(params), _to_torch(indices).long()
    return Tensor(torch.index_select(p, axis % p.dim() if p.dim() else 0, i.reshape(-1)).reshape(
        p.shape[:axis % p.dim()] + i.shape + p.shape[axis % p.dim() + 1:]) if i.dim() != 1 else torch.index_select(p, axis % p.dim(), i))


def pad(tensor, paddings, constant_values=0):
    t = _to_torch(tensor)
    flat = []
    for before, after in reversed([list(_ints(p)) for p in paddings]):
        flat += [before, after]
    return Tensor(torch.nn.functional.pad(t, flat, value=_as_int(constant_values) if not t.dtype.is_floating_point else float(constant_values)))


def where(condition, x=None, y=None):
    return Tensor(torch.where(_to_torch(condition), _to_torch(x), _to_torch(y)))


def transpose(x, perm=None):
    t = _to_torch(x)
    return Tensor(t.permute(*perm) if perm is not None else t.permute(*reversed(range(t.dim()))))


def stop_gradient(x):
    return Tensor(_to_torch(x).detach())


# ------------------------------------------------------------------------------------------ math
def einsum(eq, *ops):
    return Tensor(torch.einsum(eq.replace(" ", ""), *[_to_torch(o) for o in ops]))


def tensordot(a, b, axes):
    return Tensor(torch.tensordot(_to_torch(a), _to_torch(b), dims=axes))


def norm(x, ord="euclidean", axis=None, keepdims=False):
    t = _to_torch(x)
    if axis is None:
        return Tensor(torch.sqrt((t * t).sum()))
    return Tensor(torch.sqrt((t * t).sum(dim=axis, keepdim=keepdims)))


def _unary(fn):
    return lambda x, name=None: Tensor(fn(_to_torch(x)))


abs = _unary(torch.abs)
sign = _unary(torch.sign)
exp = _unary(torch.exp)
sigmoid = _unary(torch.sigmoid)
cos = _unary(torch.cos)
sin = _unary(torch.sin)
asin = _unary(torch.asin)
atan = _unary(torch.atan)


def atan2(y, x):
    return Tensor(torch.atan2(_to_torch(y), _to_torch(x)))


def _binary(fn):
    def op(a, b, name=None):
        ta = _to_torch(a)
        tb = _to_torch(b)
        dt = ta.dtype if isinstance(a, Tensor) or not isinstance(b, Tensor) else tb.dtype
        return Tensor(fn(ta.to(dt), tb.to(dt)))
    return op


maximum = _binary(torch.maximum)
minimum = _binary(torch.minimum)
less = _binary(torch.lt)


def reduce_sum(x, axis=None, keepdims=False):
    t = _to_torch(x)
    return Tensor(t.sum() if axis is None else t.sum(dim=axis, keepdim=keepdims))


def reduce_min(x, axis=None, keepdims=False):
    t = _to_torch(x)
    return Tensor(t.min() if axis is None else t.min(dim=axis, keepdim=keepdims).values)


def reduce_max(x, axis=None, keepdims=False):
    t = _to_torch(x)
    return Tensor(t.max() if axis is None else t.max(dim=axis, keepdim=keepdims).values)


def while_loop(cond, body, loop_vars, shape_invariants=None, **kwargs):
    vars_ = list(loop_vars)
    while bool(_to_torch(cond(*vars_))):
        vars_ = list(body(*vars_))
    return vars_


def assert_equal(x, y, *args, **kwargs):
    a = np.asarray(list(x) if isinstance(x, TensorShape) else _np(x))
    b = np.asarray(list(y) if isinstance(y, TensorShape) else _np(y))
    if not np.array_equal(a, b):
        raise ValueError(f"tf.assert_equal failed: {a} vs {b}")


def _np(x):
    return x.numpy() if isinstance(x, Tensor) else np.asarray(x)


def assert_rank(x, rank_, *args, **kwargs):
    if _to_torch(x).dim() != rank_:
        raise ValueError("tf.assert_rank failed")


def assert_greater(x, y, *args, **kwargs):
    if not np.all(_np(x) > _np(y)):
        raise ValueError("tf.assert_greater failed")


math = types.SimpleNamespace(
    sin=sin, cos=cos, exp=exp, log=_unary(torch.log), maximum=maximum, minimum=minimum, sigmoid=sigmoid, abs=abs,
    reciprocal=_unary(torch.reciprocal), mod=_binary(torch.remainder), sqrt=_unary(torch.sqrt))


# ---------------------------------------------------------------------------------------- linalg
def _matvec(a, b, **kwargs):
    return Tensor((_to_torch(a) @ _to_torch(b)[..., None])[..., 0])


def _matmul(a, b, **kwargs):
    return Tensor(_to_torch(a) @ _to_torch(b))


def _diag(x):
    return Tensor(torch.diag_embed(_to_torch(x)))


def _diag_part(x):
    return Tensor(torch.diagonal(_to_torch(x), dim1=-2, dim2=-1))


def _normalize(x, ord="euclidean", axis=None):
    t = _to_torch(x)
    n = torch.sqrt((t * t).sum(dim=axis, keepdim=True))
    return Tensor(t / n), Tensor(n)


def _pinv(a, rcond=None, validate_args=False, name=None):
    """tensorflow/python/ops/linalg/linalg_impl.py (2.10) pinv: rcond = 10 * max(rows, cols) * eps(dtype);
    singular values <= rcond * max(s) are replaced by inf; a_pinv = (V / s) @ U^H."""
    t = _to_torch(a)
    if rcond is None:
        rcond = 10.0 * max(t.shape[-2], t.shape[-1]) * float(torch.finfo(t.dtype).eps)
    u, s, vh = torch.linalg.svd(t, full_matrices=False)
    cutoff = rcond * s.max(dim=-1).values
    s = torch.where(s > cutoff[..., None], s, torch.full_like(s, float("inf")))
    return Tensor((vh.transpose(-1, -2) / s[..., None, :]) @ u.transpose(-1, -2))


linalg = types.SimpleNamespace(matvec=_matvec, matmul=_matmul, diag=_diag, diag_part=_diag_part, norm=norm,
                               normalize=_normalize, pinv=_pinv)


def _assert_near(x, y, rtol=None, atol=None, **kwargs):
    a, b = _np(x), _np(y)
    eps = float(np.finfo(a.dtype).eps) if a.dtype.kind == "f" else 0.0
    if not np.allclose(a, b, rtol=10 * eps if rtol is None else rtol, atol=10 * eps if atol is None else atol):
        raise ValueError("tf.debugging.assert_near failed")


debugging = types.SimpleNamespace(assert_near=_assert_near)


# ---------------------------------------------------------------------------------------- lookup
class _KeyValueTensorInitializer:
    def __init__(self, keys, values, *args, **kwargs):
        self.keys = [k.value() if isinstance(k, StringTensor) else k for k in keys]
        self.values = [int(v) for v in values]


class _StaticHashTable:
    def __init__(self, initializer, default_value, name=None):
        self.map = dict(zip(initializer.keys, initializer.values))
        self.default = int(default_value)

    def lookup(self, key):
        k = key.value() if isinstance(key, StringTensor) else key
        if isinstance(k, bytes):
            k = k.decode()
        return Tensor(torch.tensor(self.map.get(k, self.default), dtype=torch.int32))

    __getitem__ = lookup


lookup = types.SimpleNamespace(KeyValueTensorInitializer=_KeyValueTensorInitializer, StaticHashTable=_StaticHashTable)


# -------------------------------------------------------------------------------------- autodiff
class GradientTape:
    """tf.GradientTape over torch autograd.  torch records every op on tensors that require grad, so the
    tape only has to mark watched tensors; gradient() is torch.autograd.grad with create_graph=True (the
    reference differentiates its gradients again, helper/rmp_helper.py:50-60)."""

    def __init__(self, persistent=False, watch_accessed_variables=True):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, x):
        if not x.t.requires_grad:
            x.t.requires_grad_(True)

    @staticmethod
    def _grad(target, source):
        if not target.requires_grad:
            return torch.zeros_like(source)
        g, = torch.autograd.grad(target, source, create_graph=True, allow_unused=True)
        return torch.zeros_like(source) if g is None else g

    def gradient(self, target, sources, unconnected_gradients="none"):
        return Tensor(self._grad(_to_torch(target).sum(), _to_torch(sources)))

    def jacobian(self, target, sources, experimental_use_pfor=True, **kwargs):
        y, x = _to_torch(target), _to_torch(sources)
        rows = [self._grad(yi, x) for yi in y.reshape(-1)]
        return Tensor(torch.stack(rows).reshape(tuple(y.shape) + tuple(x.shape)))

    def batch_jacobian(self, target, source, experimental_use_pfor=True, **kwargs):
        y, x = _to_torch(target), _to_torch(source)
        cols = [self._grad(y[:, i].sum(), x) for i in range(y.shape[1])]         # rows are independent
        if not cols:
            return Tensor(torch.zeros(y.shape[0], 0, x.shape[1], dtype=x.dtype))
        return Tensor(torch.stack(cols, dim=1))
