"""Peephole pass over a backward program: replace an inner aggregation by a forward aggregation that is already stored.

What it is for (``stgraph/compiler/passes/peephole.py:82-134`` has the same purpose): the
gradient of ``out[v] = sum_e w(e) * X[u]`` with respect to something inside ``w`` contains an aggregation of its own,

    A[v] = sum_e  c(v) * w(e) * <g[v], X[u]>            (stock GATConv: the softmax-denominator term),

which starts a second kernel because a unit holds one aggregation stage.  Since ``c`` and ``g`` do not depend on the
edge, ``A[v] = c(v) * <g[v], sum_e w(e) X[u]> = c(v) * <g[v], out[v]>``: node-wise arithmetic on a tensor the forward
kernel stored anyway.  After the rewrite stock GATConv's backward is ONE compiled unit (three output aggregations)
instead of two, i.e. two launches (source-parallel + destination-parallel) instead of four; the node-wise arithmetic
itself is marked for hoisting and runs once per node as a torch unit in front of the kernel (evaluated inside the
source-parallel launch it would be redone, with three 512-byte gathers, for every edge: measured 15.6 ms against
12.7 ms for the unfactored program on config 3).

How (own design; the reference runs sympy over the whole program and accepts any textual shortening): every value is
tracked as a monomial ``coef * prod(var ** exp)`` over the variables the backward kernel reads from memory, through
``Mul`` / ``TrueDiv`` chains; a lane reduction ``Sum`` is transparent (it commutes with the products of its narrow
factors) and remembered.  An aggregation ``A = AggSum(P)`` that is not an output of the program is rewritten when some
stored forward aggregation ``K = AggSum(Q)`` onto the same side satisfies ``mono(P) / mono(Q) = r`` where every
variable of ``r`` lives on that side (or is a parameter) -- the condition under which ``r`` leaves the sum; the new
statements are ``t = K (* or /) the factors of r``, then the remembered ``Sum``.
"""
from __future__ import annotations

from collections import Counter

from ..program import Stmt, Var
from ..schema import Schema
from ..utils import ValType, is_const_scalar


class _Mono:
    __slots__ = ("coef", "fac", "summed")

    def __init__(self, coef=1.0, fac=None, summed=None):
        self.coef = float(coef)
        self.fac = Counter(fac or {})
        self.summed = summed          # the Sum statement the value went through (at most one), or None

    def combine(self, other, sign):
        if self.summed is not None and other.summed is not None:
            return None
        out = _Mono(self.coef * (other.coef ** sign), self.fac, self.summed or other.summed)
        for v, e in other.fac.items():
            out.fac[v] += sign * e
        out.fac = Counter({v: e for v, e in out.fac.items() if e != 0})
        return out


def _monomials(stmts):
    """{Var: _Mono} for the results of ``stmts`` (vars produced elsewhere are atoms)."""
    table = {}

    def of(a):
        if is_const_scalar(a):
            return _Mono(float(a))
        return table.get(a) or _Mono(1.0, {a: 1})

    for st in stmts:
        name = st.op_name.lower()
        m = None
        if name in ("mul", "truediv") and len(st.args) == 2:
            m = of(st.args[0]).combine(of(st.args[1]), 1 if name == "mul" else -1)
        elif name == "sum" and not st.is_agg():
            inner = of(st.args[0])
            if inner.summed is None:
                m = _Mono(inner.coef, inner.fac, st)
        if m is not None:
            table[st.ret] = m
    return table


def factor_aggregations(ids, bprog, forward_stmts, stored, outputs):
    """Rewrite eligible inner aggregations of ``bprog`` in place.  ``forward_stmts``: the forward region's statements;
    ``stored``: forward Vars the backward kernel may read (inputs and unit rets); ``outputs``: the gradient Vars the
    program must deliver.  Returns the list of (old aggregation Var, replacement Var)."""
    fmono = _monomials([s for s in forward_stmts if not s.is_agg()])
    known = []                                    # (K var, side, monomial of its summand)
    for s in forward_stmts:
        if s.is_agg() and s.op_name.lower() == "aggsum" and s.ret in stored:
            q = s.args[0]
            mq = fmono.get(q) or _Mono(1.0, {q: 1})
            # the summand must be expressed in stored variables only (what the backward kernel can name)
            if mq.summed is None and all(v in stored for v in mq.fac):
                known.append((s.ret, s.ret.val_type, mq))
    if not known:
        return []
    done = []
    changed = True
    while changed:
        changed = False
        bmono = _monomials([s for s in bprog if not s.is_agg()])
        for st in list(bprog):
            if not (st.is_agg() and st.op_name.lower() == "aggsum") or st.ret in outputs:
                continue
            p = st.args[0]
            mp = bmono.get(p)
            if mp is None:
                continue
            side = st.ret.val_type
            for kvar, kside, mq in known:
                if kside != side:
                    continue
                r = mp.combine(mq, -1)
                if r is None or not r.fac and mp.summed is None and r.coef == 1.0:
                    continue
                if not all((v.val_type == side or v.is_param()) for v in r.fac):
                    continue                      # something edge- or other-side-typed is left: cannot leave the sum
                if not all(e > 0 for e in mq.fac.values()) and any(v not in mp.fac for v in mq.fac):
                    continue
                if mp.summed is not None and list(kvar.var_shape) != list(mp.summed.args[0].var_shape):
                    continue                      # the remembered lane reduction must act on K's shape
                new = []
                cur = kvar
                for v, e in sorted(r.fac.items(), key=lambda t: t[0].id):
                    for _ in range(abs(e)):
                        s2 = Stmt.create_binary_bcast_stmt(ids, Schema("Mul" if e > 0 else "TrueDiv"), [cur, v])
                        new.append(s2)
                        cur = s2.ret
                if mp.summed is not None:
                    red = mp.summed
                    ret = Var.create_var(ids, red.ret.var_shape, cur.var_dtype, cur.val_type, device=cur.device,
                                         requires_grad=False)
                    new.append(Stmt(red.op_schema, [cur], ret, red.callback))
                    cur = ret
                if r.coef != 1.0:
                    s2 = Stmt.create_binary_bcast_stmt(ids, Schema("Mul"), [r.coef, cur])
                    new.append(s2)
                    cur = s2.ret
                if list(cur.var_shape) != list(st.ret.var_shape) or cur.val_type != side:
                    continue                      # shapes disagree: leave the aggregation alone
                for s2 in new:
                    s2.ret._requires_grad = False
                    s2.hoist = True               # node-wise on stored tensors: run once per node (fusion.py)
                idx = bprog.stmts.index(st)
                bprog.stmts[idx:idx + 1] = new
                bprog.replace_uses(st.ret, cur)
                done.append((st.ret, cur))
                changed = True
                break
            if changed:
                break
    return done
