"""Dead-code elimination (``stgraph/compiler/passes/dce.py:1-8``): drop statements nobody reads."""
from ..utils import is_const_scalar


def DCE(prog, output_vars):
    live = set(output_vars)
    keep = []
    for s in reversed(prog):
        if s.ret in live:
            keep.append(s)
            for a in s.args:
                if not is_const_scalar(a):
                    live.add(a)
    keep.reverse()
    prog.stmts = keep
    return prog
