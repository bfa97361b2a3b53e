"""Enums and small type-inference helpers of the vertex-program IR.

Observable rules follow ``stgraph/compiler/utils.py:1-96``:
value types SRC / DEST / EDGE / PARAM; the result of an op is EDGE-typed when its
non-PARAM operands disagree, else their common type; op types S / D / E follow the
result type and any schema whose name contains "agg" is an aggregation (A).
"""
from __future__ import annotations

from collections.abc import Iterable
from enum import Enum

var_prefix = "V"
cen_attr_postfix = "cen"
inb_attr_postfix = "inb"


class EdgeDirection(Enum):
    IN = 0
    OUT = 1


class ValType(Enum):
    SRC = 0
    DEST = 1
    EDGE = 2
    PARAM = 3


class OpType(Enum):
    S = 0
    E = 1
    A = 2
    D = 3


class ParallelMode(Enum):
    SrcParallel = 0
    DstParallel = 1


def is_const_scalar(val) -> bool:
    return type(val) in (str, int, float, bool)


def infer_val_type(vals) -> ValType:
    """EDGE if the non-PARAM operands have different types, else their common type (PARAM if all are)."""
    assert isinstance(vals, Iterable)
    kinds = [v.val_type for v in vals if not is_const_scalar(v) and v.val_type != ValType.PARAM]
    if not kinds:
        return ValType.PARAM
    return kinds[0] if all(k == kinds[0] for k in kinds) else ValType.EDGE


def infer_op_type(op_name: str, args) -> OpType | None:
    if "agg" in op_name.lower():
        return OpType.A
    t = infer_val_type(args)
    return {ValType.EDGE: OpType.E, ValType.SRC: OpType.S, ValType.DEST: OpType.D}.get(t)


def bcast_dim(var_list) -> list:
    """Element-wise maximum of the operand shapes (operands must have the same rank)."""
    shapes = [list(v.var_shape) for v in var_list if not is_const_scalar(v)]
    assert shapes, "need at least one non-constant operand"
    rank = len(shapes[0])
    out = list(shapes[0])
    for s in shapes[1:]:
        assert len(s) == rank, f"operands must have the same rank: {shapes}"
        out = [max(a, b) for a, b in zip(out, s)]
    return out


def numel(shape) -> int:
    n = 1
    for d in shape:
        n *= int(d)
    return n
