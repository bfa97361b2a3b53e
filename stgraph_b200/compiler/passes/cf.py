"""Constant folding: ``x = Mul(y, 1)`` -> uses of ``x`` read ``y`` (``stgraph/compiler/passes/cf.py:22-30``)."""
from ..utils import is_const_scalar


def CF(prog):
    replaced = {}
    for s in list(prog):
        if s.op_name.lower() == "mul" and len(s.args) == 2:
            ones = [i for i, a in enumerate(s.args) if is_const_scalar(a) and a == 1]
            if ones:
                other = s.args[1 - ones[0]]
                if not is_const_scalar(other):
                    prog.replace_uses(s.ret, other)
                    replaced[s.ret] = other
                    prog.remove(s)
    return replaced
