"""Vertex-centric compiler (mirror of ``stgraph.compiler``): tracer, IR, passes, autodiff, executor."""
from .node import CentralNode
from .program import Program, Stmt, Var
from .stgraph import Context, STGraph

__all__ = ["STGraph", "Context", "CentralNode", "Program", "Stmt", "Var"]
