"""Symbolic torch value: operator overloads append IR statements while the user function runs once.

Behaviour follows ``stgraph/compiler/val/pytorch/torch_val.py:6-227``:
``*``, ``+``, ``-``, ``/`` append ``Mul/Add/Sub/TrueDiv``; Python's builtin ``sum`` over the
one-element neighbour list calls ``0 + val`` -> ``__radd__`` -> ``AggSum`` typed DEST; ``.sum`` /
``.view`` append node-wise ``Sum`` / ``View`` statements executed through torch.  Shape inference
runs the op on a *meta* tensor of the per-element shape (the reference runs it on the mean over
axis 0 of the real data, ``torch_val.py:13-16``): same shapes, no compute.
"""
from __future__ import annotations

import torch

from ...program import Stmt, Var
from ...schema import Schema
from ...utils import ValType, infer_val_type
from ..val import Val


class TorchVal(Val):
    def __init__(self, trace, tensor, val_type, vid=None, reduce_dim=False, meta=None):
        super().__init__(tensor, vid, trace.fprog)
        self.trace = trace
        self._val_type = val_type
        if meta is not None:
            self._v = meta
            dtype, device, rg = meta.dtype, trace.device, False
        else:
            shape = tuple(tensor.shape[1:]) if reduce_dim else tuple(tensor.shape)
            self._v = torch.empty(shape, dtype=tensor.dtype, device="meta")
            dtype, device, rg = tensor.dtype, tensor.device, tensor.requires_grad
        self.var = Var.create_var(trace.ids, list(self._v.shape), dtype, val_type, var_id=vid, device=device,
                                  requires_grad=rg)

    # -- metadata ---------------------------------------------------------------
    @property
    def val_type(self):
        return self._val_type

    @property
    def dtype(self):
        return self._v.dtype

    @property
    def size(self):
        return list(self._v.shape)

    @property
    def requires_grad(self):
        return self.var.requires_grad

    @property
    def backend(self):
        return ("torch", torch)

    # -- helpers ----------------------------------------------------------------
    def _emit(self, name, other, fn, callback, swap=False, **params):
        operands = (other, self) if swap else (self, other)
        vtype = infer_val_type([o for o in operands if isinstance(o, TorchVal)])
        metas = [o.v if isinstance(o, TorchVal) else o for o in operands]
        ret = TorchVal(self.trace, None, vtype, meta=fn(*metas))
        args = [o.var if isinstance(o, TorchVal) else o for o in operands]
        self.fprog.append_stmt(Stmt(Schema(name, **params), args, ret.var, callback))
        return ret

    def __mul__(self, other):
        return self._emit("Mul", other, lambda a, b: a * b, lambda a, b: a * b)

    def __rmul__(self, other):
        return self.__mul__(other)

    def __add__(self, other):
        if not isinstance(other, TorchVal):
            raise NotImplementedError("Add of a constant is not supported by the vertex-program IR")
        return self._emit("Add", other, lambda a, b: a + b, lambda a, b: a + b)

    def __radd__(self, other):
        # builtin sum([...]) starts with int 0: this is the edge aggregation (torch_val.py:117-127)
        assert isinstance(other, int) and other == 0, "only sum(<neighbour list>) maps to AggSum"
        assert self.val_type in (ValType.SRC, ValType.EDGE), "AggSum aggregates neighbour or edge values"
        ret = TorchVal(self.trace, None, ValType.DEST, meta=self.v)
        self.fprog.append_stmt(Stmt(Schema("AggSum"), [self.var], ret.var))
        return ret

    def __sub__(self, other):
        if not isinstance(other, TorchVal):
            raise NotImplementedError("Sub of a constant is not supported by the vertex-program IR")
        return self._emit("Sub", other, lambda a, b: a - b, lambda a, b: a - b)

    def __truediv__(self, other):
        if not isinstance(other, TorchVal):
            raise NotImplementedError("TrueDiv by a constant is not supported by the vertex-program IR")
        return self._emit("TrueDiv", other, lambda a, b: a / b, lambda a, b: a / b)

    def __floordiv__(self, other):
        raise NotImplementedError("__floordiv__ Op not supported")

    def sum(self, *args, **kargs):
        ret = TorchVal(self.trace, None, self.val_type, meta=self.v.sum(*args, **kargs))

        def call(t, *rest):
            return t.sum(*rest, **kargs)

        self.fprog.append_stmt(Stmt(Schema("Sum", **kargs), [self.var] + list(args), ret.var, call))
        return ret

    def view(self, *args, **kargs):
        ret = TorchVal(self.trace, None, self.val_type, meta=self.v.view(*args, **kargs))

        def call(t, *rest):
            return t.view(-1, *rest, **kargs)   # the real tensor still has its leading node/edge axis

        self.fprog.append_stmt(Stmt(Schema("View", **kargs), [self.var] + list(args), ret.var, call))
        return ret
