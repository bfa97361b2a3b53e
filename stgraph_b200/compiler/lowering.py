"""Lower compiled execution units to launches of the pre-compiled sm_100a kernels.

This replaces the reference's code generator (``stgraph/compiler/code_gen/code_gen.py:39-116``,
``kernel_context.py:126-204``, Jinja templates, run-time ``nvcc``): nothing is generated or
compiled at run time.  A unit becomes one launch per *side* that owns an aggregation
(destination-parallel over the in-edge CSR, source-parallel over the out-edge CSR), so a
reduction onto the "other" side is a plain row reduction on the transposed structure instead
of the reference's ``atomicAdd`` inside the edge loop (``registry.py:67-102``).

Each launch is first matched against the hand-written fast path
(``stg_agg_scaled_sum_f32``: ``out = rs * sum(ns * es * x)`` -- every GCNConv unit, forward and
backward, weighted or not) and otherwise lowered to a register program for the fused VM
kernel (``stg_vm_run_f32``), which covers the whole op registry.
"""
from __future__ import annotations

import ctypes

from .. import _lib
from . import registry as R
from .utils import ValType, is_const_scalar, numel

#: STG_SHARE_EDGE=0: every launch of a unit recomputes the edge values it aggregates (A/B)
SHARE_EDGE_VALUES = __import__("os").environ.get("STG_SHARE_EDGE", "1") != "0"
PRE, LOOP, POST = 0, 1, 2
SIDE_CENTER, SIDE_NBR, SIDE_EDGE, SIDE_PARAM = 0, 1, 2, 3


class Launch:
    """One kernel launch: ``kind`` in {"scaled_sum", "vm"}; ``center`` = ValType.DEST (in-edge CSR) or SRC."""

    def __init__(self, kind, center):
        self.kind = kind
        self.center = center
        self.writes = []          # Vars this launch fully or partially writes
        self.needs_zero = set()   # Vars that must be zero-filled first (atomic / partial writes)
        # scaled_sum
        self.x = self.ns = self.es = self.rs = self.out = None
        # vm
        self.program = None
        self.tensor_vars = []

    def __repr__(self):
        if self.kind == "scaled_sum":
            return (f"Launch(scaled_sum, center={self.center.name}, x={self.x}, ns={self.ns}, es={self.es}, "
                    f"rs={self.rs}, out={self.out})")
        return f"Launch(vm, center={self.center.name}, tensors={self.tensor_vars}, n_instr={self.program.n_instr})"


def _shape2(shape):
    s = list(shape)
    if len(s) > 2:
        raise NotImplementedError("per-element shapes with more than two dims are not supported")
    return [1] * (2 - len(s)) + s


def _slice(unit, targets, stop=()):
    """Statements of the unit needed to compute ``targets`` (in program order); ``stop``: Vars another launch has
    already stored (read from memory instead of being recomputed), unless they are targets themselves."""
    produced = {s.ret: s for s in unit.program}
    need, stack = set(), list(targets)
    while stack:
        v = stack.pop()
        st = produced.get(v)
        if st is None or st in need or (v in stop and v not in targets):
            continue
        need.add(st)
        stack.extend(st.var_args())
    return [s for s in unit.program if s in need]


def _side_of(var, center):
    if var.is_param():
        return SIDE_PARAM
    if var.is_edgevar():
        return SIDE_EDGE
    return SIDE_CENTER if var.val_type == center else SIDE_NBR


def plan_unit(unit):
    """Split a compiled unit into per-side target sets: [(center ValType, [target Vars])]."""
    rets = set(unit._rets)
    aggs = [s for s in unit.program if s.is_agg()]
    sides = []
    for vt in (ValType.DEST, ValType.SRC):
        tg = [s.ret for s in aggs if s.ret.val_type == vt]
        tg += [s.ret for s in unit.program if getattr(s, "phase", "E") == "P" and s.ret in rets
               and s.ret.val_type == vt and s.ret not in tg]
        if tg:
            sides.append((vt, tg))
    # materialised edge-phase values (e.g. GAT's [E,H,1] scores) ride on the first launch
    loop_rets = [s.ret for s in unit.program if not s.is_agg() and getattr(s, "phase", "E") == "E" and s.ret in rets]
    if not sides:
        sides.append((ValType.DEST, []))
    sides[0] = (sides[0][0], sides[0][1] + loop_rets)
    return sides


# ------------------------------------------------------------------ fast path
def _match_scaled_sum(unit, center, targets, stmts):
    rets = set(unit._rets)
    aggs = [s for s in stmts if s.is_agg()]
    if len(aggs) != 1 or aggs[0].op_name.lower() != "aggsum":
        return None
    agg = aggs[0]
    dims = _shape2(unit.max_dims())
    lanes = numel(dims)
    if numel(agg.ret.var_shape) != lanes or numel(agg.args[0].var_shape) != lanes:
        return None
    produced = {s.ret: s for s in stmts}
    if any(s.op_name.lower() != "mul" for s in stmts if not s.is_agg()):
        return None
    written = [s.ret for s in stmts if s.ret in rets]

    def leaves(v, out):
        st = produced.get(v)
        if st is None or st.is_agg():
            out.append(v)
            return True
        for a in st.args:
            if is_const_scalar(a):
                return False
            if not leaves(a, out):
                return False
        return True

    inner = []
    if not leaves(agg.args[0], inner):
        return None
    # post chain: the single written var, reached from the agg result through Mul statements only
    if len(written) != 1:
        return None
    out_var = written[0]
    post = []
    if out_var != agg.ret:
        if not leaves(out_var, post) or post.count(agg.ret) != 1:
            return None
        post.remove(agg.ret)
    if numel(out_var.var_shape) != lanes:
        return None
    x = ns = es = rs = None
    for v in inner:
        side = _side_of(v, center)
        n = numel(v.var_shape)
        if side == SIDE_NBR and n == lanes and x is None:
            x = v
        elif side == SIDE_NBR and n == 1 and ns is None:
            ns = v
        elif side == SIDE_EDGE and n == 1 and es is None:
            es = v
        elif side == SIDE_CENTER and n == 1 and rs is None:
            rs = v
        else:
            return None
    for v in post:
        if _side_of(v, center) == SIDE_CENTER and numel(v.var_shape) == 1 and rs is None:
            rs = v
        else:
            return None
    if x is None:
        return None
    la = Launch("scaled_sum", center)
    la.x, la.ns, la.es, la.rs, la.out = x, ns, es, rs, out_var
    la.writes = [out_var]
    return la


# ------------------------------------------------------------------ VM path
class _VmBuilder:
    def __init__(self, unit, center):
        self.unit = unit
        self.center = center
        self.dims = _shape2(unit.max_dims())
        self.instr = {PRE: [], LOOP: [], POST: []}
        self.tensors = []          # Vars
        self.tensor_meta = []      # (side, bc0, bc1)
        self.reg_of = {}
        self.n_regs = 0
        self.acc_init = []
        self.acc_kind = []

    def new_reg(self):
        r = self.n_regs
        self.n_regs += 1
        if self.n_regs > _lib.VM_MAX_REGS:
            raise NotImplementedError("vertex program needs more registers than the VM kernel provides")
        return r

    def tensor_index(self, var):
        if var in self.tensors:
            return self.tensors.index(var)
        a, b = self.dims
        s0, s1 = _shape2(var.var_shape)
        if s0 not in (1, a) or s1 not in (1, b):
            raise NotImplementedError(f"{var} does not broadcast against the unit shape {self.dims}")
        self.tensors.append(var)
        self.tensor_meta.append((_side_of(var, self.center), 1 if s0 == a else 0, 1 if s1 == b else 0))
        if len(self.tensors) > _lib.VM_MAX_TENSORS:
            raise NotImplementedError("vertex program touches more tensors than the VM kernel provides")
        return len(self.tensors) - 1

    def emit(self, phase, op, dst=0, a=0, b=0, imm=0.0):
        self.instr[phase].append((op, phase, dst, a, b, float(imm)))

    def operand(self, arg, phase):
        """Register holding ``arg`` (loading tensors / constants on first use)."""
        if is_const_scalar(arg):
            r = self.new_reg()
            self.emit(phase, R.OP_CONST, dst=r, imm=float(arg))
            return r
        if arg in self.reg_of:
            return self.reg_of[arg]
        side = _side_of(arg, self.center)
        load_phase = PRE if side in (SIDE_CENTER, SIDE_PARAM) else LOOP
        if load_phase == LOOP and phase != LOOP:
            raise NotImplementedError(f"{arg} (neighbour/edge data) is read outside the edge loop")
        r = self.new_reg()
        self.emit(load_phase, R.OP_LOAD, dst=r, a=self.tensor_index(arg))
        self.reg_of[arg] = r
        return r

    def _reuse_registers(self, seq):
        """Linear-scan renaming of the virtual registers (they live in shared memory: fewer = more rows per SM).

        A value defined before the edge loop and read inside it stays live until the loop ends; a value
        defined inside the loop is dead after its last use in the same iteration.
        """
        n_pre, n_loop = len(self.instr[PRE]), len(self.instr[LOOP])
        loop_end = n_pre + n_loop

        def reads(ins):
            op, _, dst, a, b, _ = ins
            if op in (R.OP_LOAD, R.OP_CONST, R.OP_ACC_READ):
                return []
            if op in (R.OP_ACC_SUM, R.OP_ACC_MAX, R.OP_ACC_MIN):
                return [a]
            if op == R.OP_STORE:
                return [b]
            if op in (R.OP_ADD, R.OP_SUB, R.OP_MUL, R.OP_DIV, R.OP_RELU_BWD, R.OP_AMAX_BWD):
                return [a, b]
            return [a]          # unary ops, GSUM

        def writes(ins):
            op = ins[0]
            if op in (R.OP_ACC_SUM, R.OP_ACC_MAX, R.OP_ACC_MIN, R.OP_STORE):
                return None
            return ins[2]

        first_def, last_use = {}, {}
        for i, ins in enumerate(seq):
            for r in reads(ins):
                last_use[r] = i
            w = writes(ins)
            if w is not None and w not in first_def:
                first_def[w] = i
        for r, d in first_def.items():
            u = last_use.get(r, d)
            if d < n_pre and n_pre <= u < loop_end:
                u = loop_end                          # read in every iteration: stays live until the loop has ended
            last_use[r] = max(u, d)
        free, mapping, out, peak = [], {}, [], 0
        expire = {}
        for r, u in last_use.items():
            expire.setdefault(u, []).append(r)
        for i, ins in enumerate(seq):
            op, ph, dst, a, b, imm = ins
            # which OPERAND SLOTS hold registers (STORE's `a` and LOAD's `a` are tensor indices, ACC_*'s `dst` / ACC_READ's
            # `a` accumulator indices): a tensor index that happens to equal a live register number must not be renamed
            if op in (R.OP_LOAD, R.OP_CONST, R.OP_ACC_READ):
                a_is_reg = b_is_reg = False
            elif op == R.OP_STORE:
                a_is_reg, b_is_reg = False, True
            elif op in (R.OP_ADD, R.OP_SUB, R.OP_MUL, R.OP_DIV, R.OP_RELU_BWD, R.OP_AMAX_BWD):
                a_is_reg = b_is_reg = True
            else:                                     # ACC_*, unary ops, GSUM: `a` only
                a_is_reg, b_is_reg = True, False
            na = mapping[a] if a_is_reg and a in mapping else a
            nb = mapping[b] if b_is_reg and b in mapping else b
            for r in expire.get(i, []):               # operands read here die here: their slot can hold the result
                if r in mapping and first_def.get(r, -1) < i:
                    free.append(mapping[r])
            w = writes(ins)
            nd = dst
            if w is not None:
                if w not in mapping:
                    mapping[w] = free.pop() if free else peak
                    if mapping[w] == peak:
                        peak += 1
                nd = mapping[w]
                if last_use.get(w, i) == i and first_def.get(w) == i:
                    free.append(mapping[w])           # never read
            out.append((op, ph, nd, na, nb, imm))
        return out, max(peak, 1)

    def finish(self):
        prog = _lib.StgVmProgram()
        prog.dim0, prog.dim1 = self.dims
        seq = self.instr[PRE] + self.instr[LOOP] + self.instr[POST]
        seq, self.n_regs = self._reuse_registers(seq)
        if len(seq) > _lib.VM_MAX_INSTR:
            raise NotImplementedError("vertex program is longer than the VM kernel's instruction buffer")
        prog.n_tensors = len(self.tensors)
        prog.n_instr = len(seq)
        prog.n_regs = max(self.n_regs, 1)
        prog.n_acc = len(self.acc_init)
        prog.n_pre = len(self.instr[PRE])
        prog.n_loop = len(self.instr[LOOP])
        for i, v in enumerate(self.acc_init):
            prog.acc_init[i] = v
            prog.acc_kind[i] = self.acc_kind[i]
        for i, (side, bc0, bc1) in enumerate(self.tensor_meta):
            prog.tensors[i].side, prog.tensors[i].bc0, prog.tensors[i].bc1 = side, bc0, bc1
        for i, (op, ph, dst, a, b, imm) in enumerate(seq):
            ins = prog.instr[i]
            ins.op, ins.phase, ins.dst, ins.a, ins.b, ins.imm = op, ph, dst, a, b, imm
        return prog


def _lower_vm(unit, center, targets, stmts):
    rets = set(unit._rets)
    b = _VmBuilder(unit, center)
    la = Launch("vm", center)
    center_op = "D" if center == ValType.DEST else "S"
    dims = b.dims
    acc_of = {}

    def can_gsum():
        return dims[1] <= 32          # segments of dim1 consecutive lanes inside one warp chunk

    for st in stmts:
        name = st.op_name.lower()
        opdef = R.look_up_registry(name)
        if opdef is None or opdef.vm_op is None:
            raise NotImplementedError(f"op {st.op_name} cannot run inside a fused kernel")
        if st.is_agg():
            if st.ret.val_type != center:
                continue                       # the other side's launch computes it
            k = len(b.acc_init)
            if k >= _lib.VM_MAX_ACC:
                raise NotImplementedError("too many aggregations in one unit")
            b.acc_init.append(opdef.acc_init)
            b.acc_kind.append({"aggmax": 1, "aggmin": 2}.get(name, 0))
            src = b.operand(st.args[0], LOOP)
            b.emit(LOOP, opdef.vm_op, dst=k, a=src)
            acc_of[st.ret] = k
            r = b.new_reg()
            b.emit(POST, R.OP_ACC_READ, dst=r, a=k, b=1 if name == "aggmean" else 0)
            shrink = numel(st.ret.var_shape) != numel(st.args[0].var_shape)
            consumers = [s for s in stmts if st.ret in s.var_args()]
            if shrink and name != "aggsum":
                raise NotImplementedError("only AggSum may reduce across feature lanes")
            if shrink and consumers:
                if not (can_gsum() and _shape2(st.ret.var_shape)[1] == 1 and _shape2(st.ret.var_shape)[0] == dims[0]):
                    raise NotImplementedError("lane reduction feeding further ops needs a last dim <= 32")
                r2 = b.new_reg()
                b.emit(POST, R.OP_GSUM, dst=r2, a=r, b=1)
                r, shrink = r2, False
            b.reg_of[st.ret] = r
            if st.ret in rets:
                b.emit(POST, R.OP_STORE, a=b.tensor_index(st.ret), b=r, imm=1.0 if shrink else 0.0)
                la.writes.append(st.ret)
                if shrink and not (can_gsum() and _shape2(st.ret.var_shape)[1] == 1):
                    la.needs_zero.add(st.ret)
            continue
        phase = POST if getattr(st, "phase", "E") == "P" else LOOP
        if phase == LOOP and st.is_nodewise() and st.op_type.name == center_op:
            phase = PRE                         # centre-side prologue: loop invariant
        if phase == POST and not (st.is_nodewise() and st.op_type.name == center_op):
            continue                            # epilogue of the other side
        if name == "sum":
            dim = st.op_schema.params.get("dim")
            rank = len(st.args[0].var_shape)
            if dim is None or (dim % rank) != rank - 1 or not can_gsum() or len(st.args[0].var_shape) < 2:
                raise NotImplementedError("in-kernel Sum is supported over the last dim (<= 32) only")
            src = b.operand(st.args[0], phase)
            r = b.new_reg()
            b.emit(phase, R.OP_GSUM, dst=r, a=src, b=1)
        else:
            regs = [b.operand(a, phase) for a in st.args]
            r = b.new_reg()
            imm = float(st.op_schema.params.get("negative_slope", 0.0)) if "leakyrelu" in name else 0.0
            b.emit(phase, opdef.vm_op, dst=r, a=regs[0], b=regs[1] if len(regs) > 1 else 0, imm=imm)
        b.reg_of[st.ret] = r
        if st.ret in rets and st.ret in targets:
            side = _side_of(st.ret, center)
            sphase = phase
            if side == SIDE_CENTER and phase == PRE:
                sphase = POST
            b.emit(sphase, R.OP_STORE, a=b.tensor_index(st.ret), b=r, imm=0.0)
            la.writes.append(st.ret)
            if side != SIDE_CENTER:
                la.needs_zero.add(st.ret)      # entries no row touches stay zero, like the reference's new_zeros
    la.program = b.finish()
    la.tensor_vars = list(b.tensors)
    return la


def lower_unit(unit):
    """Return the launches (in execution order) that evaluate a compiled unit."""
    launches = []
    sides = plan_unit(unit)
    stop = set()
    if len(sides) == 2 and SHARE_EDGE_VALUES:
        # An edge value that feeds aggregations on BOTH sides (GAT backward: the score gradient goes to d_el on the source
        # side and to d_er on the destination side) is computed and stored by the first launch and loaded by the
        # second, instead of re-evaluating its whole dependency chain (here a [H,D]-wide product + lane sum) per edge
        # in both launches.
        first = {st.ret for st in _slice(unit, sides[0][1])}
        second_targets = set(sides[1][1])
        shared = []
        for st in unit.program:
            if st.is_agg() and st.ret in second_targets:
                p = st.args[0]
                if not is_const_scalar(p) and p.is_edgevar() and p in first and p not in shared:
                    shared.append(p)
        for p in shared:
            unit.add_ret_val(p)
        if shared:
            sides[0] = (sides[0][0], sides[0][1] + [p for p in shared if p not in sides[0][1]])
            stop = set(shared)
    for i, (center, targets) in enumerate(sides):
        stmts = _slice(unit, targets, stop if i == 1 else ())
        la = _match_scaled_sum(unit, center, targets, stmts)
        if la is None:
            la = _lower_vm(unit, center, targets, stmts)
        launches.append(la)
    written = [v for la in launches for v in la.writes]
    missing = [v for v in unit._rets if v not in written]
    if missing:
        raise NotImplementedError(f"unit {unit.kernel_name}: no launch produces {missing}")
    unit.launches = launches
    return launches
