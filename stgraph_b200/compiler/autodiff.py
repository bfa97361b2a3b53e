"""Reverse-mode differentiation of a traced vertex program.

Observable behaviour follows ``stgraph/compiler/autodiff.py:16-142``:
walk the forward statements backwards, ask each op for its gradient statements
(``registry.py``), accumulate several contributions to one variable with ``Add``, produce
gradient units only for inputs with ``requires_grad``.  Forward values the gradient statements
read are either *materialised* (inputs and unit rets -- the executor saves exactly those on its
state stack, ``executor.py:64-72``) or *recomputed* inside the backward kernel from materialised
ones (the reference's ``dep_program`` slice with stopping vars, ``passes/mem_planning.py:20-25``);
aggregation results are never recomputed, they are materialised by the forward kernel.

Reference quirk kept on purpose (it changes gradients): ``Stmt.grad`` returns
``dict(op_impl.grad(...))`` (``program.py:328-329``), so when one variable feeds SEVERAL operands
of the same statement only the LAST operand's contribution survives (``Sub(V0, V0)`` in stock
GATConv gives ``dV0 = dV1`` once; ``x*x`` would give ``g*x``, not ``2*g*x``).
"""
from __future__ import annotations

import os

from .passes import DCE, fuse, optimize, resolve
from .passes.peephole import factor_aggregations
from .program import Program, Stmt, Var
from .registry import GradCtx, look_up_registry
from .schema import Schema
from .utils import is_const_scalar

#: STG_PEEPHOLE=0 keeps inner aggregations as separate units (A/B runs, tests of the unfactored program)
PEEPHOLE = os.environ.get("STG_PEEPHOLE", "1") != "0"


def diff(ids, forward_units, out_vars):
    """Return ``(backward_units, grad_in, grad_out)``.

    ``grad_in``: {forward output Var -> Var holding its incoming gradient};
    ``grad_out``: {forward input Var -> Var holding the gradient to return}.
    Forward units get the extra rets the backward program needs (materialisation).
    """
    compiled = [u for u in forward_units if u.compiled]
    fstmts = [s for u in compiled for s in u.program]
    produced = {s.ret: s for s in fstmts}
    unit_of = {s.ret: u for u in compiled for s in u.program}
    region_inputs = []
    for s in fstmts:
        for a in s.var_args():
            if a not in produced and a not in region_inputs:
                region_inputs.append(a)
    consumed_elsewhere = set()
    for u in forward_units:
        if not u.compiled:
            consumed_elsewhere.update(u._args)
    region_outputs = [v for u in compiled for v in u.unit_rets()
                      if (v in out_vars or v in consumed_elsewhere) and v.requires_grad]

    ctx = GradCtx(ids)
    grad_in = {}
    grad_map = {}
    for y in region_outputs:
        g = Var.create_var(ids, y.var_shape, y.var_dtype, y.val_type, device=y.device, requires_grad=False)
        grad_in[y] = g
        grad_map[y] = g

    bstmts = []
    for s in reversed(fstmts):
        y = s.ret
        if y not in grad_map:
            continue
        opdef = look_up_registry(s.op_name)
        if opdef is None or opdef.grad is None:
            raise NotImplementedError(f"no gradient rule for op {s.op_name}")
        contributions = {}                       # dict semantics: the last operand position wins
        for pos, x in enumerate(s.args):
            if is_const_scalar(x) or not x.requires_grad:
                continue
            contributions[x] = opdef.grad(ctx, s, pos, x, y, grad_map[y])
        for x, (stmts, g) in contributions.items():
            bstmts.extend(stmts)
            if x in grad_map:
                acc = Var.create_var(ids, x.var_shape, x.var_dtype, x.val_type, device=x.device, requires_grad=False)
                bstmts.append(Stmt(Schema("Add"), [grad_map[x], g], acc))
                grad_map[x] = acc
            else:
                grad_map[x] = g

    grad_out = {x: grad_map[x] for x in region_inputs if x.requires_grad and x in grad_map}
    if not grad_out:
        return [], grad_in, grad_out

    # ---- forward values read by the gradient statements: materialise or recompute
    materialised = set(region_inputs)
    for u in compiled:
        materialised.update(u._rets)
    recompute, seen = [], set()

    def need(v):
        if v in materialised or v in seen or v not in produced:
            return
        st = produced[v]
        if st.is_agg():
            unit_of[v].add_ret_val(v)            # aggregation results are stored by the forward kernel
            materialised.add(v)
            return
        seen.add(v)
        for a in st.var_args():
            need(a)
        recompute.append(st)

    for st in bstmts:
        for a in st.var_args():
            need(a)
    order = {s: i for i, s in enumerate(fstmts)}
    recompute.sort(key=lambda s: order[s])
    # recomputed statements are re-created so the backward program owns its Vars' producer links
    remap = {}
    re_stmts = []
    for st in recompute:
        ret = Var.create_var(ids, st.ret.var_shape, st.ret.var_dtype, st.ret.val_type, device=st.ret.device,
                             requires_grad=False)
        re_stmts.append(Stmt(st.op_schema, [remap.get(a, a) if not is_const_scalar(a) else a for a in st.args],
                             ret, st.callback))
        remap[st.ret] = ret
    for st in bstmts:
        st.args = [remap.get(a, a) if not is_const_scalar(a) else a for a in st.args]

    bprog = Program(re_stmts + bstmts)
    DCE(bprog, list(grad_out.values()))
    replaced = optimize(bprog)
    grad_out = {x: resolve(g, replaced) for x, g in grad_out.items()}
    if PEEPHOLE:
        # inner aggregations that equal node-wise arithmetic on a forward aggregation the kernel stored (GAT: K2 is one
        # unit again); a forward aggregation that is only NAMED by the rewrite becomes a stored ret of its unit
        stored = set(materialised)
        for s in fstmts:
            if s.is_agg():
                stored.add(s.ret)
        rewritten = factor_aggregations(ids, bprog, fstmts, stored, list(grad_out.values()))
        if rewritten:
            for st in bprog:
                for a in st.var_args():
                    if a in produced and produced[a].is_agg() and a not in materialised:
                        unit_of[a].add_ret_val(a)
                        materialised.add(a)
            DCE(bprog, list(grad_out.values()))
            replaced = optimize(bprog)
            grad_out = {x: resolve(g, replaced) for x, g in grad_out.items()}
    backward_units = fuse(bprog, list(grad_out.values()))
    return backward_units, grad_in, grad_out
