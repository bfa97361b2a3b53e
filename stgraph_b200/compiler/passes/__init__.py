"""IR passes (mirror of ``stgraph/compiler/passes``): CF, CSE, DCE and the fusion partitioner."""
from .cf import CF
from .cse import CSE
from .dce import DCE
from .fusion import fuse


def optimize(prog):
    """``passes/__init__.py:7-9``: constant folding then common-subexpression elimination."""
    replaced = dict(CF(prog))
    replaced.update(CSE(prog))
    return replaced


def resolve(var, replaced):
    """Follow a replacement map produced by ``optimize`` to the surviving variable."""
    while var in replaced:
        var = replaced[var]
    return var


__all__ = ["CF", "CSE", "DCE", "fuse", "optimize", "resolve"]
