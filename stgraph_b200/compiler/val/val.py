"""Abstract symbolic value seen by a vertex program while it is being traced (``stgraph/compiler/val/val.py``)."""
import abc


class Val(abc.ABC):
    """``_t``: the real tensor, ``_v``: a per-element stand-in used for shape inference, ``var``: the IR variable."""

    def __init__(self, tensor, vid, fprog):
        self._t = tensor
        self._id = vid
        self._v = None
        self.var = None
        self.fprog = fprog

    @property
    def v(self):
        return self._v

    @property
    def id(self):
        return self._id

    def __str__(self):
        return str(self.var)

    __repr__ = __str__
