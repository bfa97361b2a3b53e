"""The vertex-program decorator: trace once, differentiate, fuse, lower, execute.

API kept from ``stgraph/compiler/stgraph.py:24-226``::

    self.stgraph = STGraph(STGraphBackendTorch())

    @self.stgraph.compile(gnn_module=self)
    def nb_compute(v):
        return sum([nb.h * nb.norm for nb in v.innbs]) * v.norm

    out = nb_compute(g=graph, n_feats={"norm": norm, "h": h}, e_feats={...})

First call: the function runs ONCE on a symbolic ``CentralNode`` whose features are
``TorchVal``s; while it runs, every function of the ``torch`` namespace and the module's
parameters / buffers / sub-modules are replaced by symbolic stand-ins (restored in a
``finally`` block -- the reference leaves torch patched if tracing raises).  The traced
``Program`` is optimised, split into execution units, differentiated and lowered to
pre-compiled kernels (no ``nvcc`` at run time, unlike ``code_gen/compiler.py:14-44``).

Deliberate fixes: contexts are cached by (function name, feature
names), not by name alone (trap T5: a GCNConv first called without edge weights kept its
unweighted kernel forever); executors are cached per feature-shape signature.
"""
from __future__ import annotations

import functools
import types

import torch

from .autodiff import diff
from .backend.callback import STGraphBackend
from .executor import Executor
from .node import CentralNode
from .op.pytorch.torch_op import TorchOp
from .passes import fuse, optimize, resolve
from .program import Program, VarIds
from .utils import ValType, cen_attr_postfix, inb_attr_postfix, var_prefix
from .val.pytorch.torch_val import TorchVal


class _Trace:
    """State of one tracing run (program being built + id allocator)."""

    def __init__(self, device):
        self.fprog = Program()
        self.ids = VarIds()
        self.device = device


class Context:
    def __init__(self, func, nspace, run_cb):
        functools.update_wrapper(self, func)
        self._f = func
        self._nspace = nspace
        self._run_cb = run_cb
        self._entry_count = 0
        self._executors = {}          # shape signature -> (Executor, param input map)
        self._executor_cache = None
        self.last_program = None      # traced + optimised forward Program (introspection / tests)

    # ------------------------------------------------------------------ call
    def __call__(self, **kwargs):
        executor = self._setup_executor(**kwargs)
        ret = self._run_cb(executor)
        if len(ret) == 1:
            return ret[0]
        return ret

    def _setup_executor(self, **kwargs):
        graph = kwargs.get("g", None)
        node_feats = kwargs.get("n_feats", {}) or {}
        edge_feats = kwargs.get("e_feats", {}) or {}
        if not graph:
            raise NameError("Need to provide the graph as one of keyward arguments")
        sig = tuple((k, tuple(v.shape[1:]), v.dtype, v.requires_grad) for k, v in sorted(node_feats.items())) + \
            tuple(("e:" + k, tuple(v.shape[1:]), v.dtype, v.requires_grad) for k, v in sorted(edge_feats.items()))
        if sig not in self._executors:
            param_map = {}
            trace, rets = self._trace(node_feats, edge_feats, param_map)
            self._executors[sig] = (self._diff_then_compile(trace, rets, graph), param_map)
        executor, param_map = self._executors[sig]
        self._executor_cache = executor
        input_map = {}
        for k, v in node_feats.items():
            input_map[var_prefix + k + cen_attr_postfix] = v
            input_map[var_prefix + k + inb_attr_postfix] = v
        for k, v in edge_feats.items():
            input_map[var_prefix + k] = v
        for key, owner_and_name in param_map.items():
            owner, name = owner_and_name
            input_map[key] = owner[name]
        executor.restart(input_map, graph)
        self._entry_count += 1
        return executor

    # ----------------------------------------------------------------- trace
    def _trace(self, nfeats, efeats, param_map):
        any_t = next(iter(list(nfeats.values()) + list(efeats.values())))
        trace = _Trace(any_t.device)
        cen = CentralNode()
        for k, v in nfeats.items():
            setattr(cen, k, TorchVal(trace, v, ValType.DEST, vid=k + cen_attr_postfix, reduce_dim=True))
            for nb in cen.innbs:
                setattr(nb, k, TorchVal(trace, v, ValType.SRC, vid=k + inb_attr_postfix, reduce_dim=True))
        for k, v in efeats.items():
            if v.dim() < 2:
                raise ValueError(f"edge feature '{k}' must be [E, ...] with at least one trailing dim (got {tuple(v.shape)})")
            for e in cen.inedges:
                setattr(e, k, TorchVal(trace, v, ValType.EDGE, vid=k, reduce_dim=True))
        undo = []
        try:
            self._patch_namespaces(trace, param_map, undo)
            ret = self._f(cen)
        finally:
            for owner, key, old in reversed(undo):
                owner[key] = old
        if ret is None:
            raise NameError("Ret is none. Execution is aborted")
        vals = list(ret) if isinstance(ret, (tuple, list)) else [ret]
        return trace, [v.var for v in vals]

    def _patch_namespaces(self, trace, param_map, undo):
        """Symbolise torch functions and the module's parameters / buffers / sub-modules (``stgraph.py:126-173``)."""
        for nspace in self._nspace:
            d = nspace.__dict__
            if "__name__" in d:           # a module namespace such as ``torch``
                for key in list(d.keys()):
                    m = d[key]
                    if isinstance(m, (types.FunctionType, types.BuiltinFunctionType)) and not key.startswith("_"):
                        undo.append((d, key, m))
                        d[key] = TorchOp(m, trace)
            else:                         # the nn.Module that owns the vertex program
                for group in ("_parameters", "_buffers"):
                    table = d.get(group, {})
                    for name in list(table.keys()):
                        t = table[name]
                        if t is None:
                            continue
                        undo.append((table, name, t))
                        param_map[var_prefix + name] = (table, name)
                        table[name] = TorchVal(trace, t, ValType.PARAM, vid=name, reduce_dim=False)
                mods = d.get("_modules", {})
                for name in list(mods.keys()):
                    sub = mods[name]
                    if sub is None:
                        continue
                    undo.append((mods, name, sub))
                    mods[name] = TorchOp(sub, trace)

    # --------------------------------------------------------------- compile
    def _diff_then_compile(self, trace, out_vars, graph):
        fprog = trace.fprog
        replaced = optimize(fprog)
        out_vars = [resolve(v, replaced) for v in out_vars]
        self.last_program = fprog
        forward_units = fuse(fprog, out_vars)
        backward_units, grad_in, grad_out = diff(trace.ids, forward_units, out_vars)
        self.forward_units, self.backward_units = forward_units, backward_units
        return Executor(graph, forward_units, backward_units, grad_in, grad_out, out_vars)


def _rebuild_stgraph(backend_cls):
    return STGraph(backend_cls())


class STGraph:
    def __init__(self, backend_framework: STGraphBackend):
        self._ctx_map = {}
        self._backend_framework = backend_framework
        self._run_cb = backend_framework.backend_cb

    # Traced programs and executors are caches of THIS instance, and the backend object holds the ``torch`` module: a
    # deep copy / pickle of the layer that owns an STGraph (``copy.deepcopy(model)``, ``torch.save(model)``) gets a fresh,
    # empty one.  (The reference's layers cannot be copied or pickled at all for that reason.)
    def __deepcopy__(self, memo):
        return _rebuild_stgraph(type(self._backend_framework))

    def __reduce__(self):
        return _rebuild_stgraph, (type(self._backend_framework),)

    def compile(self, gnn_module, hetero_graph=False):
        namespace = [gnn_module, self._backend_framework.backend_module]

        def wrapper(func):
            if hetero_graph:
                raise NotImplementedError("Heterogeneous graph is not supported yet")
            # keyed by name AND code object: both GCNConv vertex functions are called nb_compute (trap T5)
            key = (func.__name__, id(getattr(func, "__code__", func)))
            ctx = self._ctx_map.get(key)
            if ctx is None:
                ctx = Context(func, namespace, self._run_cb)
                self._ctx_map[key] = ctx
            else:
                ctx._f = func     # rebind the closure of this call (same code object)
            return ctx

        return wrapper
