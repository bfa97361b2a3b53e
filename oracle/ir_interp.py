"""IR-level oracle: evaluates execution units with torch-CPU gather / ``index_add_`` (fp64).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Any traced vertex program therefore has an oracle, and the stock ``GATConv`` trace (with its
``emb - emb`` degeneracy and the reference's gradient rules) is reproduced by construction
(SURVEY.md section 7 step 0, section 8(c)).  Semantics per statement follow the generated
kernels (``/root/reference/stgraph/compiler/code_gen/templates/fa/tpl_fa_csr_unsorted.jinja:20-57``
and the per-op expressions of ``registry.py:195-406``):

* a SRC-typed operand is read at ``src[e]``, a DEST-typed one at ``dst[e]``, an EDGE-typed one at
  the edge id, a PARAM as is (``kernel_context.py:167-204``);
* node-wise statements act on node tensors, edge-wise ones on per-edge tensors;
* ``AggSum/Max/Min/Mean`` reduce per-edge values onto the side of their result type; a result
  shape smaller than the operand's means the lanes are summed too (the reference's atomic
  "cross-lane" write, ``kernel_context.py:126-149``).

Edges are given as ``(src, dst)`` arrays ordered by edge id.
"""
from __future__ import annotations

import torch

from stgraph_b200.compiler.utils import ValType, is_const_scalar


def _reduce_shape(t, shape):
    """Sum ``t`` ([M, *big]) over the dims where ``shape`` has 1 and ``t`` does not."""
    tgt = [t.shape[0]] + list(shape)
    if list(t.shape) == tgt:
        return t
    for d, (a, b) in enumerate(zip(t.shape[1:], shape), start=1):
        if a != b:
            assert b == 1
            t = t.sum(dim=d, keepdim=True)
    return t


class Interp:
    def __init__(self, src, dst, num_nodes, dtype=torch.float64):
        self.src = torch.as_tensor(src, dtype=torch.int64)
        self.dst = torch.as_tensor(dst, dtype=torch.int64)
        self.n = num_nodes
        self.e = self.src.shape[0]
        self.dtype = dtype

    def _edge_view(self, var, t):
        if var.val_type == ValType.SRC:
            return t[self.src]
        if var.val_type == ValType.DEST:
            return t[self.dst]
        if var.val_type == ValType.EDGE:
            return t
        return t.unsqueeze(0)

    def _eval_op(self, stmt, args):
        name = stmt.op_name.lower()
        p = stmt.op_schema.params
        a = args
        if name == "add":
            return a[0] + a[1]
        if name == "sub":
            return a[0] - a[1]
        if name == "mul":
            return a[0] * a[1]
        if name == "truediv":
            return a[0] / a[1]
        if name == "exp":
            return torch.exp(a[0])
        if name == "leakyrelu":
            return torch.where(a[0] > 0, a[0], a[0] * p["negative_slope"])
        if name == "backwardleakyrelu":
            return torch.where(a[0] > 0, torch.ones_like(a[0]), torch.full_like(a[0], p["negative_slope"]))
        if name == "relu":
            return torch.clamp(a[0], min=0)
        if name == "backwardrelu":
            return torch.where(a[0] > 0, a[1], torch.zeros_like(a[1]))
        if name == "backwardamax":
            return (a[0] == a[1]).to(a[0].dtype)
        if name == "sum":
            return a[0].sum(dim=(p["dim"] % (a[0].dim() - 1)) + 1, keepdim=True)
        raise NotImplementedError(name)

    def run_units(self, units, tensors):
        """Evaluate units in order; ``tensors``: {var id: tensor}.  Returns the (updated) map (fp64)."""
        env = {k: (v.to(self.dtype) if torch.is_floating_point(v) else v) for k, v in tensors.items()}
        for u in units:
            for st in u.program:
                name = st.op_name.lower()
                if not u.compiled and st.callback is not None:
                    env[st.ret.id] = st.callback(*[env[x.id] if not is_const_scalar(x) else x for x in st.args])
                    continue
                if st.is_agg():
                    x = st.args[0]
                    ev = self._edge_view(x, env[x.id])
                    idx = self.dst if st.ret.val_type == ValType.DEST else self.src
                    shape = [self.n] + list(ev.shape[1:])
                    if name in ("aggsum", "aggmean"):
                        out = torch.zeros(shape, dtype=self.dtype).index_add_(0, idx, ev)
                        if name == "aggmean":
                            cnt = torch.zeros(self.n, dtype=self.dtype).index_add_(0, idx, torch.ones(self.e, dtype=self.dtype))
                            out = out / cnt.clamp(min=1).reshape([-1] + [1] * (out.dim() - 1))
                    elif name == "aggmax":
                        out = torch.full(shape, -float("inf"), dtype=self.dtype).index_reduce_(0, idx, ev, "amax")
                    elif name == "aggmin":
                        out = torch.full(shape, float("inf"), dtype=self.dtype).index_reduce_(0, idx, ev, "amin")
                    else:
                        raise NotImplementedError(name)
                    env[st.ret.id] = _reduce_shape(out, st.ret.var_shape)
                    continue
                if st.is_nodewise():
                    args = [env[x.id] if not is_const_scalar(x) else x for x in st.args]
                    args = [t.unsqueeze(0) if (not is_const_scalar(x) and x.is_param()) else t
                            for x, t in zip(st.args, args)]
                else:
                    args = [self._edge_view(x, env[x.id]) if not is_const_scalar(x) else x for x in st.args]
                env[st.ret.id] = self._eval_op(st, args)
        return env
