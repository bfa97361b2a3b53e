"""Explicit aggregation ops (``stgraph/compiler/op/agg/agg_op.py:9-37``).

The reference registers AggMax / AggMin / AggMean in the IR but leaves no public way to trace
them (its import is commented out, ``stgraph/compiler/stgraph.py:6``).  These helpers take the
one-element neighbour list a vertex program builds and emit the aggregation typed DEST.
"""
from ...program import Stmt
from ...schema import Schema
from ...utils import ValType
from ...val.pytorch.torch_val import TorchVal


def _agg(name, vals):
    val = vals[0] if isinstance(vals, (list, tuple)) else vals
    assert isinstance(val, TorchVal) and val.val_type in (ValType.SRC, ValType.EDGE)
    ret = TorchVal(val.trace, None, ValType.DEST, meta=val.v)
    val.fprog.append_stmt(Stmt(Schema(name), [val.var], ret.var))
    return ret


def agg_sum(vals):
    return _agg("AggSum", vals)


def agg_max(vals):
    return _agg("AggMax", vals)


def agg_min(vals):
    return _agg("AggMin", vals)


def agg_mean(vals):
    return _agg("AggMean", vals)
