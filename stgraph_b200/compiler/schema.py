"""Op schema: a name plus keyword parameters (``stgraph/compiler/schema.py:1-10``)."""


class Schema:
    def __init__(self, op_name, **kargs):
        self._op_name = op_name
        self._params = kargs

    @property
    def op_name(self):
        return self._op_name

    @property
    def params(self):
        return self._params

    def __eq__(self, other):
        return isinstance(other, Schema) and self._op_name == other._op_name and self._params == other._params

    def __hash__(self):
        return hash(self._op_name)

    def __str__(self):
        return f"{self._op_name}({self._params})"

    __repr__ = __str__
