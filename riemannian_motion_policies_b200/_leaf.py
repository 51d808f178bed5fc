"""Base class of all leaf policies (reference: rmp.py:184-206 and rmp2.py:6-29)."""
import numpy as np
import torch

from . import _native
from ._tensor import current_stream_ptr, like_input, require_cuda, to_device, unwrap


def as_float_list(value, length=None, what="vector"):
    value = unwrap(value)
    if isinstance(value, torch.Tensor):
        value = value.detach().cpu().numpy()
    arr = np.asarray(value, dtype=np.float64).reshape(-1)
    if length is not None and arr.shape[0] != length:
        raise ValueError(f"{what} has {arr.shape[0]} entries, expected {length}")
    return [float(v) for v in arr]


import inspect


class RiemannianMotionPolicy:
    """Abstract leaf: a task map plus ``evaluate(x, xd) -> (xdd, M)``."""

    leaf_type = None

    def __init__(self, name, taskmap):
        self.name = name
        self.taskmap = taskmap

    def __setattr__(self, key, value):
        # Every attribute assignment (the reference idiom ``leaf.goal = ...``, ``leaf.repulsion_gain = ...``) bumps a
        # version counter; a compiled tree re-derives a leaf's kernel parameters only when it moved (or when the leaf
        # holds vector parameters, which may also be mutated in place).
        object.__setattr__(self, key, value)
        object.__setattr__(self, "_version", self.__dict__.get("_version", 0) + 1)

    @classmethod
    def _has_vector_parameters(cls):
        return cls._vec is not RiemannianMotionPolicy._vec

    # -- description handed to the C ABI --------------------------------------------------------
    def _params(self):
        """Constructor arguments in the order documented in include/rmp2_b200.h."""
        raise NotImplementedError

    def _vec(self, dim):
        """Vector parameter (goal / q0 / limits) for a task space of dimension ``dim``."""
        return []

    def leaf_desc(self, dim, space=_native.SPACE_CONFIG, frame=-1, goal_slot=-1):
        if self.leaf_type is None:
            raise NotImplementedError(f"{type(self).__name__} is not implemented by the CUDA engine")
        d = _native.LeafDesc()
        d.type, d.space, d.frame, d.goal_slot = self.leaf_type, space, frame, goal_slot
        params = [float(p) for p in self._params()]
        for i, p in enumerate(params):
            d.params[i] = p
        vec = [] if goal_slot >= 0 else self._vec(dim)
        if len(vec) > len(d.vec):
            raise NotImplementedError("task space too large")
        for i, v in enumerate(vec):
            d.vec[i] = v
        return d

    def _aux(self, K, dev):
        """Extra per-row data a leaf holds itself (only the v1 CollisionAvoidance: distance, normal)."""
        return None

    # -- stand-alone evaluation, same signature as the reference --------------------------------
    def evaluate(self, x, xd, *args, **kwargs):
        """x, xd [K,m] -> xdd [K,m], M [K,m,m] through the CUDA leaf kernel
        (reference: rmp.py:202-206 / rmp2.py:25-29)."""
        dev = require_cuda()
        xt, xdt = to_device(x, dev), to_device(xd, dev)
        if xt.dim() != 2 or xt.shape != xdt.shape:
            raise ValueError("x and xd must both be [K, m]")
        K, m = xt.shape
        desc = self.leaf_desc(m)
        xdd = torch.empty(K, m, device=dev)
        M = torch.empty(K, m, m, device=dev)
        aux = self._aux(K, dev)
        _native.check(_native.lib().rmp2_leaf_evaluate(desc, m, K, xt.data_ptr(), xdt.data_ptr(),
                                                       None if aux is None else aux.data_ptr(), xdd.data_ptr(),
                                                       M.data_ptr(), current_stream_ptr(dev)))
        return like_input(xdd, x), like_input(M, x)


class DeclaredLeaf(RiemannianMotionPolicy):
    """Leaf whose constructor is declared as data: ``FIELDS`` lists the constructor arguments in the
    reference's order, as ``name`` or ``(name, default)``; every argument becomes an attribute of the same
    name, exactly as the reference's constructors do by hand.  ``taskmap`` / ``name`` may be among them;
    a leaf without a ``taskmap`` argument lives on the identity task map."""

    FIELDS = ()
    PARAMS = ()          # attribute names handed to the kernel, in the order of include/rmp2_b200.h

    @classmethod
    def _signature(cls):
        ps = []
        for f in cls.FIELDS:
            name, default = (f, inspect.Parameter.empty) if isinstance(f, str) else f
            ps.append(inspect.Parameter(name, inspect.Parameter.POSITIONAL_OR_KEYWORD, default=default))
        return inspect.Signature(ps)

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        if cls.FIELDS:
            cls.__signature__ = cls._signature()

    def __init__(self, *args, **kwargs):
        bound = self._signature().bind(*args, **kwargs)
        bound.apply_defaults()
        values = dict(bound.arguments)
        taskmap = values.pop("taskmap", None)
        if taskmap is None:
            from .taskmap import IdentityTaskmap
            taskmap = IdentityTaskmap()
        super().__init__(values.pop("name"), taskmap)
        for key, value in values.items():
            setattr(self, key, value)
        self._post_init()

    def _post_init(self):
        pass

    def _params(self):
        return [getattr(self, k) for k in self.PARAMS]
