"""Op table of the IR: per-op gradient rule, VM opcode and torch restatement.

This table *is* the op set the pre-compiled kernels cover.  It mirrors the
``OpImpl`` registry of ``stgraph/compiler/registry.py:195-406`` (Add, Sub, Mul,
TrueDiv, Exp, LeakyRelu, Relu, AggSum, AggMax, BackwardAMax, BackwardRelu,
BackwardLeakyRelu, GTypeCast; lookup is case-insensitive on the schema name,
``registry.py:25-36``), but where the reference attaches a C-expression string to
each op (``gen_code``) for its Jinja/nvcc back-end, here each op carries the
opcode interpreted by the fused VM kernel (``csrc/vm.cu``) and a torch callable
used for uncompiled (node-wise) units.

Gradient rules are restated exactly, including the reference's ``Sub`` rule that
propagates ``+1`` to *both* operands (``registry.py:210-213``; SURVEY.md trap T2) --
stock ``GATConv`` gradients depend on it.
"""
from __future__ import annotations

import os
import warnings

import torch

from .program import Stmt, Var
from .schema import Schema
from .utils import ValType, infer_val_type, is_const_scalar

# VM opcodes -- keep in sync with StgVmOp in include/stgraph_b200.h
OP_LOAD, OP_CONST, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_EXP, OP_LRELU, OP_LRELU_BWD, OP_RELU, OP_RELU_BWD, \
    OP_AMAX_BWD, OP_ACC_SUM, OP_ACC_MAX, OP_ACC_MIN, OP_ACC_READ, OP_STORE, OP_GSUM = range(18)


class GradCtx:
    """Helpers handed to gradient rules (the reference passes ``create_var`` / ``create_stmt`` callbacks)."""

    def __init__(self, ids):
        self.ids = ids

    def var_like(self, x: Var, val_type=None, shape=None):
        return Var.create_var(self.ids, shape if shape is not None else x.var_shape, x.var_dtype,
                              val_type if val_type is not None else x.val_type, device=x.device)

    def stmt(self, name, args, ret, **params):
        return Stmt(Schema(name, **params), args, ret)

    def multiply_grad(self, dzdy: Var, dydx, x: Var):
        """``dz/dx = dz/dy * dy/dx`` with broadcasting; reduce back to ``x``'s shape when it is smaller.

        ``registry.py:136-164``.  A multiplication by the constant 1 is not emitted
        (the reference emits it and lets constant folding remove it, ``passes/cf.py:22-30``).
        """
        stmts = []
        if is_const_scalar(dydx) and dydx == 1:
            cur = dzdy
        else:
            ops = [dzdy, dydx]
            shape = list(dzdy.var_shape)
            if not is_const_scalar(dydx):
                assert len(dydx.var_shape) == len(shape), "gradient operands must have the same rank"
                shape = [max(a, b) for a, b in zip(shape, dydx.var_shape)]
            first = dzdy
            ret = Var.create_var(self.ids, shape, first.var_dtype, infer_val_type(ops), device=first.device)
            stmts.append(self.stmt("Mul", ops, ret))
            cur = ret
        return stmts, cur

    def reduce_to(self, cur: Var, x: Var, stmts: list):
        """Bring an (edge- or node-typed) gradient ``cur`` to the type and shape of ``x``.

        node-typed ``x`` fed by an edge-typed gradient -> ``AggSum`` onto ``x``'s side
        (``registry.py:182-188``); a smaller shape is a sum over the broadcast lanes --
        folded into the ``AggSum`` (its ret keeps ``x``'s shape; the reference reaches the same
        kernel through its peephole pass, ``passes/peephole.py:60-80``) or an explicit ``Sum``.
        """
        if x.is_nodevar() and cur.val_type != x.val_type:
            ret = self.var_like(x)
            stmts.append(self.stmt("AggSum", [cur], ret))
            return ret
        if list(cur.var_shape) != list(x.var_shape):
            diff = [i for i, (a, b) in enumerate(zip(cur.var_shape, x.var_shape)) if a != b]
            if len(cur.var_shape) != len(x.var_shape) or len(diff) != 1:
                raise NotImplementedError("gradient broadcast over more than one dimension is not supported")
            ret = self.var_like(x, val_type=cur.val_type)
            stmts.append(self.stmt("Sum", [cur], ret, dim=diff[0], keep_dim=True))
            return ret
        return cur


class OpDef:
    def __init__(self, name, vm_op=None, torch_fn=None, grad=None, is_agg=False, acc_init=0.0):
        self.name = name
        self.vm_op = vm_op
        self.torch_fn = torch_fn
        self.grad = grad
        self.is_agg = is_agg
        self.acc_init = acc_init


def _binary_grad(dydx_of):
    def rule(ctx: GradCtx, fstmt: Stmt, pos: int, x: Var, y: Var, grad_y: Var):
        pre, dydx = dydx_of(ctx, fstmt, pos, x, y)
        stmts, cur = ctx.multiply_grad(grad_y, dydx, x)
        stmts = pre + stmts
        out = ctx.reduce_to(cur, x, stmts)
        return stmts, out
    return rule


def _add_dydx(ctx, fstmt, pos, x, y):
    return [], 1


#: STG_SUB_GRAD=correct gives the subtrahend its mathematical gradient (-1).  The default keeps the reference's rule
#: (+1 for BOTH operands, ``registry.py:210-213``): stock GATConv's ``emb - max([emb])`` traces to ``Sub(V0, V0)`` and the
#: reference's d_el / d_er -- pinned by tests/golden/ref_kernels.npz -- depend on it (SURVEY.md trap T2).
SUB_GRAD_CORRECT = os.environ.get("STG_SUB_GRAD", "reference") == "correct"
_warned_sub = []


def _sub_dydx(ctx, fstmt, pos, x, y):
    if pos == 1 and not is_const_scalar(fstmt.args[0]) and fstmt.args[0] != fstmt.args[1]:
        if SUB_GRAD_CORRECT:
            return [], -1
        if not _warned_sub:
            _warned_sub.append(True)
            warnings.warn("vertex program computes a - b with b requiring a gradient: the reference's Sub rule passes +1 to "
                          "BOTH operands (registry.py:210-213), so d_b has the wrong sign; set STG_SUB_GRAD=correct for -1",
                          RuntimeWarning, stacklevel=2)
    return [], 1


def _mul_dydx(ctx, fstmt, pos, x, y):
    return [], fstmt.args[1 - pos]


def _div_dydx(ctx, fstmt, pos, x, y):
    a, b = fstmt.args
    pre = []
    if pos == 0:
        v = ctx.var_like(b)
        pre.append(ctx.stmt("TrueDiv", [1, b], v))
        return pre, v
    sq = ctx.var_like(b)
    pre.append(ctx.stmt("Mul", [b, b], sq))
    if is_const_scalar(a):
        neg = -a
    else:
        neg = ctx.var_like(a)
        pre.append(ctx.stmt("Mul", [-1, a], neg))
    shape = list(y.var_shape)
    v = Var.create_var(ctx.ids, shape, y.var_dtype, infer_val_type([t for t in (neg, sq) if not is_const_scalar(t)]),
                       device=y.device)
    pre.append(ctx.stmt("TrueDiv", [neg, sq], v))
    return pre, v


def _exp_grad(ctx, fstmt, pos, x, y, grad_y):
    stmts, cur = ctx.multiply_grad(grad_y, y, x)
    return stmts, ctx.reduce_to(cur, x, stmts)


def _lrelu_grad(ctx, fstmt, pos, x, y, grad_y):
    d = ctx.var_like(x)
    stmts = [ctx.stmt("BackwardLeakyRelu", [x], d, **fstmt.op_schema.params)]
    more, cur = ctx.multiply_grad(grad_y, d, x)
    stmts += more
    return stmts, ctx.reduce_to(cur, x, stmts)


def _relu_grad(ctx, fstmt, pos, x, y, grad_y):
    ret = ctx.var_like(x)
    return [ctx.stmt("BackwardRelu", [x, grad_y], ret)], ret


def _aggsum_grad(ctx, fstmt, pos, x, y, grad_y):
    """``y = AggSum(x)``: every edge of the row receives ``grad_y`` (``registry.py:262-276``)."""
    stmts = []
    cur = grad_y
    if not x.is_edgevar():
        ret = ctx.var_like(x)
        stmts.append(ctx.stmt("AggSum", [cur], ret))
        cur = ret
    return stmts, cur


def _aggmax_grad(ctx, fstmt, pos, x, y, grad_y):
    mask = ctx.var_like(x, val_type=ValType.EDGE)
    stmts = [ctx.stmt("BackwardAMax", [x, y], mask)]
    more, cur = ctx.multiply_grad(grad_y, mask, x)
    stmts += more
    if not x.is_edgevar():
        ret = ctx.var_like(x)
        stmts.append(ctx.stmt("AggSum", [cur], ret))
        cur = ret
    return stmts, cur


def _no_grad(name):
    def rule(*a, **k):
        raise NotImplementedError(f"gradient of {name} is not supported (same as the reference)")
    return rule


def _leaky(x, negative_slope=0.01, **_):
    return torch.nn.functional.leaky_relu(x, negative_slope)


_TABLE = [
    OpDef("add", OP_ADD, lambda a, b: a + b, _binary_grad(_add_dydx)),
    OpDef("sub", OP_SUB, lambda a, b: a - b, _binary_grad(_sub_dydx)),
    OpDef("mul", OP_MUL, lambda a, b: a * b, _binary_grad(_mul_dydx)),
    OpDef("truediv", OP_DIV, lambda a, b: a / b, _binary_grad(_div_dydx)),
    OpDef("exp", OP_EXP, torch.exp, _exp_grad),
    OpDef("leakyrelu", OP_LRELU, _leaky, _lrelu_grad),
    OpDef("relu", OP_RELU, torch.relu, _relu_grad),
    OpDef("backwardleakyrelu", OP_LRELU_BWD,
          lambda x, negative_slope=0.01, **_: torch.where(x > 0, torch.ones_like(x), torch.full_like(x, negative_slope)),
          _no_grad("BackwardLeakyRelu")),
    OpDef("backwardrelu", OP_RELU_BWD, lambda x, g: torch.where(x > 0, g, torch.zeros_like(g)), _no_grad("BackwardRelu")),
    OpDef("backwardamax", OP_AMAX_BWD, lambda x, y: (x == y).to(x.dtype), _no_grad("BackwardAMax")),
    OpDef("aggsum", OP_ACC_SUM, None, _aggsum_grad, is_agg=True, acc_init=0.0),
    OpDef("aggmax", OP_ACC_MAX, None, _aggmax_grad, is_agg=True, acc_init=float("-inf")),
    OpDef("aggmin", OP_ACC_MIN, None, _no_grad("AggMin"), is_agg=True, acc_init=float("inf")),
    OpDef("aggmean", OP_ACC_SUM, None, _no_grad("AggMean"), is_agg=True, acc_init=0.0),
    OpDef("sum", OP_GSUM, lambda x, dim=-1, keep_dim=True, **_: x.sum(dim=(dim % (x.dim() - 1)) + 1, keepdim=keep_dim),
          _no_grad("Sum")),
    OpDef("gtypecast", None, None, _no_grad("GTypeCast")),
]

impl_registry = {d.name: d for d in _TABLE}


def look_up_registry(op_name: str):
    """Case-insensitive lookup by schema name; None for ops only torch can run (node-wise callbacks)."""
    return impl_registry.get(op_name.lower())


def is_vm_supported(stmt: Stmt) -> bool:
    d = look_up_registry(stmt.op_name)
    return d is not None and d.vm_op is not None


def torch_eval(stmt: Stmt, tensor_args):
    """Run a (node-wise) statement with torch: the traced callback if there is one, else the table entry."""
    if stmt.callback is not None:
        return stmt.callback(*tensor_args)
    d = look_up_registry(stmt.op_name)
    if d is None or d.torch_fn is None:
        raise NotImplementedError(f"no torch implementation for op {stmt.op_name}")
    return d.torch_fn(*tensor_args, **stmt.op_schema.params)
