"""Operator fusion: partition a program into execution units (one fused kernel each).

Observable behaviour follows ``stgraph/compiler/passes/fusion.py:13-59,183-306`` (SURVEY.md
appendix B.4): walking the dataflow, a unit accepts a node-wise prologue (src and/or dst
side), edge-wise ops, **at most one aggregation stage**, and a node-wise epilogue on the
aggregation's side; a second aggregation downstream of the first starts a new unit whose inputs
are materialised (GAT: ``K0`` -> ``K1``); independent aggregations over the same inputs share a
unit (GAT backward ``K2`` has three); units made only of node-wise statements are *not* compiled
and run through torch.

Implementation: instead of the reference's DFS + 6-state machine over a mutable linked list,
every statement is assigned the earliest feasible *slot* in the timeline
``T0, E0, P0, T1, E1, P1, ...`` (T = torch unit, E = edge phase of kernel k, P = post-loop phase
of kernel k) by one forward pass over the SSA program; statements sharing ``k`` form a unit.
"""
from __future__ import annotations

from ..execution_unit import ExecutionUnit
from ..registry import is_vm_supported
from ..utils import ValType, is_const_scalar


def _is_torch_only(stmt) -> bool:
    if stmt.is_agg():
        return False
    if getattr(stmt, "hoist", False) and stmt.is_nodewise():
        # node-wise arithmetic on stored tensors that a pass asked to run ONCE PER NODE before the kernel (peephole.py):
        # inside a kernel that walks the other side's rows it would be re-evaluated, operands gathered, for every edge
        return True
    if stmt.callback is not None and stmt.op_name.lower() in ("sum", "view"):
        return True
    return not is_vm_supported(stmt)


def assign_slots(prog):
    """Return {stmt: slot}; slot = 3k (torch unit k), 3k+1 (edge phase of kernel k), 3k+2 (post phase)."""
    avail = {}          # Var -> slot where it is produced (inputs absent)
    slots = {}

    def e_avail(v):
        s = avail.get(v, -1)
        if s < 0:
            return 0
        k, ph = divmod(s, 3)
        return k + 1 if ph == 2 else k

    def t_avail(v):
        s = avail.get(v, -1)
        if s < 0:
            return 0
        k, ph = divmod(s, 3)
        return k if ph == 0 else k + 1

    for st in prog:
        vargs = st.var_args()
        if _is_torch_only(st):
            if not st.is_nodewise() and vargs:
                raise NotImplementedError(
                    f"edge-wise op '{st.op_name}' is not in the op registry (registry.py) and cannot be fused")
            slot = 3 * max([t_avail(a) for a in vargs] or [0])
        elif st.is_agg():
            slot = 3 * max(e_avail(a) for a in vargs) + 2
        else:
            k_edge = max([e_avail(a) for a in vargs] or [0])
            slot = 3 * k_edge + 1
            if st.is_nodewise() and k_edge > 0:
                k = k_edge - 1
                node_t = ValType.SRC if st.is_src() else ValType.DEST
                uses_post = any(avail.get(a, -1) == 3 * k + 2 for a in vargs)
                ok = uses_post
                for a in vargs:
                    s = avail.get(a, -1)
                    if a.is_param():
                        continue
                    if a.val_type != node_t or s == 3 * k + 1 or s > 3 * k + 2:
                        ok = False
                if ok:
                    slot = 3 * k + 2
        slots[st] = slot
        st.phase = "TEP"[slot % 3]      # torch unit / edge phase / post-loop phase (read by lowering)
        avail[st.ret] = slot
    return slots


def fuse(prog, outputs):
    """Partition ``prog`` into ordered ExecutionUnits; ``outputs`` are the Vars that must be materialised."""
    slots = assign_slots(prog)
    groups = {}
    for st in prog:
        k, ph = divmod(slots[st], 3)
        key = (k, 0 if ph == 0 else 1)
        groups.setdefault(key, []).append(st)
    units = []
    for key in sorted(groups):
        stmts = groups[key]
        compiled = key[1] == 1 and any(not s.is_nodewise() for s in stmts)
        units.append(ExecutionUnit(stmts, compiled))
    # connect: anything a later unit reads, or the caller asked for, is a ret of its producer
    for i, u in enumerate(units):
        for later in units[i + 1:]:
            for a in later._args:
                u.add_ret_val(a)
        for v in outputs:
            u.add_ret_val(v)
    return units
