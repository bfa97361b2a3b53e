"""The vertex-program IR: ``Var`` / ``Stmt`` / ``Program``.

Surface kept from ``stgraph/compiler/program.py``: a
``Var`` has an id ``V<n>`` (temporaries) or ``V<name>{cen,inb}`` (features), a
``val_type`` in SRC/DEST/EDGE/PARAM, a per-element ``var_shape`` (the leading
node/edge axis removed) and ``dtype_str``; a ``Stmt`` is ``ret = op(args)`` with an
``op_type`` in S/D/E/A; a ``Program`` is an ordered statement list.  The
implementation is a plain Python list in SSA form (the reference uses an
intrusive doubly-linked list with in-place mutation).
"""
from __future__ import annotations

import itertools

from .schema import Schema
from .utils import (OpType, ValType, bcast_dim, cen_attr_postfix, inb_attr_postfix, infer_op_type,
                    infer_val_type, is_const_scalar, var_prefix)


class VarIds:
    """Per-trace id allocator (the reference keeps a module-global counter, ``utils.py:3-4``)."""

    def __init__(self, start: int = 0):
        self._c = itertools.count(start)

    def next(self) -> int:
        return next(self._c)


class Var:
    def __init__(self, var_id, val_type, var_shape, var_dtype, device=None, requires_grad=True):
        self._id = var_prefix + str(var_id)
        self._val_type = val_type
        self._var_shape = list(var_shape)
        self._var_dtype = var_dtype
        self.dtype_str = "float" if "float" in str(var_dtype).lower() else "int"
        self._device = device
        self._requires_grad = bool(requires_grad)
        self.stmt = None          # producing statement (None for inputs)

    @classmethod
    def create_var(cls, ids: VarIds, var_shape, var_dtype, val_type, var_id=None, device=None, requires_grad=True):
        vid = var_id if var_id is not None else ids.next()
        return cls(vid, val_type, var_shape, var_dtype, device, requires_grad)

    # -- reference-compatible accessors -------------------------------------
    @property
    def id(self):
        return self._id

    @property
    def int_id(self):
        return int(self._id[len(var_prefix):])

    @property
    def val_type(self):
        return self._val_type

    @property
    def var_shape(self):
        return self._var_shape

    @var_shape.setter
    def var_shape(self, shape):
        self._var_shape = list(shape)

    @property
    def var_dtype(self):
        return self._var_dtype

    @property
    def device(self):
        return self._device

    @property
    def requires_grad(self):
        return self._requires_grad

    def is_srcvar(self):
        return self._val_type == ValType.SRC

    def is_dstvar(self):
        return self._val_type == ValType.DEST

    def is_edgevar(self):
        return self._val_type == ValType.EDGE

    def is_nodevar(self):
        return self.is_srcvar() or self.is_dstvar()

    def is_param(self):
        return self._val_type == ValType.PARAM

    @property
    def base_name(self):
        """Feature name without the cen/inb view suffix (both views alias one tensor, ``stgraph.py:59-61``)."""
        name = self._id[len(var_prefix):]
        if self.is_srcvar() and name.endswith(inb_attr_postfix):
            return name[: -len(inb_attr_postfix)]
        if self.is_dstvar() and name.endswith(cen_attr_postfix):
            return name[: -len(cen_attr_postfix)]
        return name

    def __eq__(self, other):
        return isinstance(other, Var) and self._id == other._id

    def __hash__(self):
        return hash(self._id)

    def __str__(self):
        return f"{self._id}({self._val_type.name},{self._var_shape})"

    __repr__ = __str__


class Stmt:
    def __init__(self, op_schema: Schema, args: list, ret: Var, callback=None):
        self.op_schema = op_schema
        self.args = list(args)
        self.ret = ret
        self.op_type = infer_op_type(op_schema.op_name, [a for a in self.args]) if ret is not None else None
        self.callback = callback
        if ret is not None:
            ret.stmt = self
            ret._requires_grad = any(a.requires_grad for a in self.args if not is_const_scalar(a))

    @classmethod
    def create_stmt(cls, op_schema=None, args=None, ret=None, callback=None):
        return cls(op_schema, args, ret, callback)

    @classmethod
    def create_binary_bcast_stmt(cls, ids: VarIds, op_schema, args, callback=None):
        first = next(a for a in args if not is_const_scalar(a))
        ret = Var.create_var(ids, bcast_dim(args), first.var_dtype, infer_val_type(args), device=first.device)
        return cls(op_schema, args, ret, callback)

    @property
    def op_name(self):
        return self.op_schema.op_name

    def var_args(self):
        return [a for a in self.args if not is_const_scalar(a)]

    def is_agg(self):
        return self.op_type == OpType.A

    def is_edgewise(self):
        return self.op_type == OpType.E

    def is_nodewise(self):
        return self.op_type in (OpType.S, OpType.D)

    def is_src(self):
        return self.op_type == OpType.S

    def is_dst(self):
        return self.op_type == OpType.D

    def stmt_info(self):
        """CSE key: schema + argument ids with the cen/inb suffix stripped (+ ret type unless node-wise).

        Same rule as ``program.py:238-250``: ``nb.norm`` and ``v.norm`` uses are told
        apart only by the statement's type.
        """
        parts = []
        for a in self.args:
            parts.append(str(a) if is_const_scalar(a) else var_prefix + a.base_name)
        tail = "" if self.is_nodewise() else str(self.ret.val_type)
        return f"{self.op_schema}-{parts}{tail}-{self.op_type}"

    def __str__(self):
        ot = self.op_type.name if self.op_type else "?"
        return f"{ot}: {self.ret} = {self.op_name}({self.args})"

    __repr__ = __str__


class Program:
    """Ordered list of statements (SSA: every Var is assigned once)."""

    def __init__(self, stmts=None):
        self.stmts = list(stmts) if stmts else []

    def append_stmt(self, stmt: Stmt):
        self.stmts.append(stmt)
        return stmt

    def extend(self, stmts):
        self.stmts.extend(stmts)

    def __iter__(self):
        return iter(list(self.stmts))

    def __reversed__(self):
        return reversed(list(self.stmts))

    def __len__(self):
        return len(self.stmts)

    def empty(self):
        return not self.stmts

    def input_vars(self):
        produced, inputs = set(), []
        for s in self.stmts:
            for a in s.var_args():
                if a not in produced and a not in inputs:
                    inputs.append(a)
            produced.add(s.ret)
        return inputs

    def all_vars(self):
        out = []
        for s in self.stmts:
            for v in s.var_args() + [s.ret]:
                if v not in out:
                    out.append(v)
        return out

    def find_var_by_id(self, vid):
        for v in self.all_vars():
            if v.id == vid:
                return v
        return None

    def users_of(self, var):
        return [s for s in self.stmts if var in s.var_args()]

    def replace_uses(self, old: Var, new):
        for s in self.stmts:
            s.args = [new if (not is_const_scalar(a) and a == old) else a for a in s.args]

    def remove(self, stmt):
        self.stmts = [s for s in self.stmts if s is not stmt]

    def __str__(self):
        return "\n".join(str(s) for s in self.stmts)

    __repr__ = __str__
