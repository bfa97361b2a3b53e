"""Common-subexpression elimination keyed by ``Stmt.stmt_info()`` (``stgraph/compiler/passes/cse.py:1-14``)."""


def CSE(prog):
    seen = {}
    replaced = {}
    for s in list(prog):
        key = s.stmt_info()
        if key in seen:
            prog.replace_uses(s.ret, seen[key])
            replaced[s.ret] = seen[key]
            prog.remove(s)
        else:
            seen[key] = s.ret
    return replaced
