"""Execution units: the fused groups of IR statements that become kernel launches.

Mirror of ``stgraph/compiler/execution_unit.py:15-90,170-240`` (args / rets / tmps / program /
parallel mode / ``compiled`` flag, ``kernel_args() = unit_args() + unit_rets()`` sorted by id,
``max_dims()`` = widest per-element shape).  The launch side (``Kernel.run`` packing ctypes
pointers for ``cuLaunchKernel``, ``execution_unit.py:317-415``) is replaced by
``lowering.py``: a unit is lowered once to C-ABI launches of pre-compiled sm_100a kernels.
The reference's geometry rule ``calculate_kernel_params_fa`` (``execution_unit.py:92-116``:
feat >= 64 -> one node per block; feat < 64 -> 64 threads, group = largest power of two <= feat,
which drops columns for non-power-of-two widths, trap T1) is replaced by lane groups of the
smallest power of two >= the (vectorised) width, so every column is computed.
"""
from __future__ import annotations

import itertools

from .utils import ParallelMode, is_const_scalar, numel

_unit_counter = itertools.count()


class ExecutionUnit:
    def __init__(self, stmts, compiled: bool):
        self._prog = list(stmts)
        self._compiled = compiled
        self._rets = []
        self._kernel_name = "K" + str(next(_unit_counter))
        self._parallel_mode = None
        self.launches = None       # filled by lowering for compiled units
        produced = {s.ret for s in self._prog}
        self._tmps = produced
        args = []
        for s in self._prog:
            for a in s.var_args():
                if a not in produced and a not in args:
                    args.append(a)
        self._args = args
        dst = sum(1 for s in self._prog if s.is_agg() and s.ret.is_dstvar())
        src = sum(1 for s in self._prog if s.is_agg() and s.ret.is_srcvar())
        if dst or src:
            self._parallel_mode = ParallelMode.DstParallel if dst >= src else ParallelMode.SrcParallel

    # -- reference-compatible accessors ------------------------------------------
    @property
    def program(self):
        return self._prog

    @property
    def tmps(self):
        return self._tmps

    @property
    def compiled(self):
        return self._compiled

    @property
    def kernel_name(self):
        return self._kernel_name

    def parallel_mode(self):
        return self._parallel_mode

    def unit_args(self):
        return sorted(self._args, key=lambda v: v.id)

    def unit_rets(self):
        return sorted(self._rets, key=lambda v: v.id)

    def kernel_args(self):
        return self.unit_args() + self.unit_rets()

    def add_ret_val(self, var):
        if var in self._tmps and var not in self._rets:
            self._rets.append(var)

    def all_rets(self):
        return {s.ret for s in self._prog}

    def get_all_vars(self):
        out = []
        for s in self._prog:
            for v in s.var_args() + [s.ret]:
                if v not in out:
                    out.append(v)
        return out

    def max_dims(self):
        """Widest per-element shape among the unit's variables (at most two dims, like the reference)."""
        best = None
        for v in self.get_all_vars():
            shp = list(v.var_shape)
            if len(shp) > 2:
                raise NotImplementedError("per-element feature shapes with more than 2 dims are not supported")
            if best is None or len(shp) > len(best):
                best = shp if best is None else [max(a, b) for a, b in zip(([1] * (len(shp) - len(best)) + best), shp)]
            else:
                shp = [1] * (len(best) - len(shp)) + shp
                best = [max(a, b) for a, b in zip(best, shp)]
        return best or [1]

    def feature_size(self):
        return numel(self.max_dims())

    def __str__(self):
        body = "\n  ".join(str(s) for s in self._prog)
        return (f"{self._kernel_name}[{'compiled' if self._compiled else 'torch'}, {self._parallel_mode}] "
                f"args={self.unit_args()} rets={self.unit_rets()}\n  {body}")

    __repr__ = __str__
