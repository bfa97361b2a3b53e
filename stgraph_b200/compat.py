"""``import stgraph_b200.compat`` makes ``import stgraph...`` resolve to this package.

Reference user code (``from stgraph.nn.pytorch.static.gcn_conv import GCNConv``,
``from stgraph.graph.static.static_graph import StaticGraph``,
``from stgraph.compiler.backend.pytorch.torch_callback import STGraphBackendTorch``) then runs on the
B200 backend unchanged.  Implemented as a meta-path finder that aliases ``stgraph[.x]`` to
``stgraph_b200[.x]`` (same module objects, so class identities are shared).
"""
import importlib
import importlib.abc
import importlib.util
import sys


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        if fullname == "stgraph" or fullname.startswith("stgraph."):
            real = "stgraph_b200" + fullname[len("stgraph"):]
            try:
                mod = importlib.import_module(real)
            except ImportError:
                return None
            spec = importlib.util.spec_from_loader(fullname, _AliasLoader(real), is_package=hasattr(mod, "__path__"))
            return spec
        return None


def install():
    if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _AliasFinder())


install()
