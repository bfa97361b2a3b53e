"""Symbolic stand-in for a torch function or sub-module during tracing.

``stgraph/compiler/op/pytorch/torch_op.py:5-9`` + ``op/op.py:16-42``: calling it on
symbolic values appends one statement whose schema name is the module's class name (with its
public attributes as parameters, e.g. ``LeakyReLU(negative_slope=0.2)``) or the builtin's name
(``exp``); registry lookup is case-insensitive on that name.
"""
from __future__ import annotations

import torch

from ...program import Stmt
from ...schema import Schema
from ...utils import infer_val_type
from ...val.pytorch.torch_val import TorchVal


class TorchOp:
    def __init__(self, op, trace):
        self._op = op
        self.trace = trace

    def to_schema(self) -> Schema:
        if isinstance(self._op, torch.nn.Module):
            params = {k: v for k, v in self._op.__dict__.items() if not k.startswith("_") and k != "training"}
            return Schema(type(self._op).__name__, **params)
        return Schema(getattr(self._op, "__name__", str(self._op)))

    def __call__(self, *args, **kargs):
        vals = [a for a in args if isinstance(a, TorchVal)]
        if not vals:
            return self._op(*args, **kargs)          # not part of the vertex program: run the real op
        if kargs:
            raise NotImplementedError("keyword arguments are not supported on traced ops")
        meta = self._op(*[a.v if isinstance(a, TorchVal) else a for a in args])
        if isinstance(meta, (tuple, list)):
            raise NotImplementedError(f"ops that return several tensors are not supported: {self._op}")
        ret = TorchVal(self.trace, None, infer_val_type(vals), meta=meta)
        op = self._op

        def call(*tensors):
            return op(*tensors)

        self.trace.fprog.append_stmt(Stmt(self.to_schema(), [a.var if isinstance(a, TorchVal) else a for a in args],
                                          ret.var, call))
        return ret

    def __getattr__(self, name):
        # e.g. self.leaky_relu.negative_slope read inside the vertex function
        return getattr(self.__dict__["_op"], name)
