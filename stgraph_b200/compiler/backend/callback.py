"""Backend plug-in interface -- the drop-in boundary on the Python side.

Kept verbatim in shape from ``stgraph/compiler/backend/callback.py:3-21``: a backend provides
``new_zeros_call_back(size, dtype, device, requires_grad)``, ``tensor_raw_ptr(tensor)`` and the
attributes ``backend_name / backend_module / kernel_wrapper``; ``backend_cb(executor)`` wires the
callbacks into the executor and runs the forward pass.  ``new_empty_call_back`` is an optional
addition (defaults to zeros) used for outputs a kernel overwrites completely.
"""
from abc import ABC, abstractmethod


class STGraphBackend(ABC):
    def __init__(self):
        self.backend_name = None
        self.backend_module = None
        self.kernel_wrapper = None

    @abstractmethod
    def new_zeros_call_back(self, size, dtype, device, requires_grad=True):
        pass

    @abstractmethod
    def tensor_raw_ptr(self, tensor):
        pass

    def new_empty_call_back(self, size, dtype, device, requires_grad=True):
        return self.new_zeros_call_back(size, dtype, device, requires_grad)

    def backend_cb(self, executor):
        executor.set_new_zeros_cb(self.new_zeros_call_back)
        executor.set_new_empty_cb(self.new_empty_call_back)
        executor.set_raw_ptr_cb(self.tensor_raw_ptr)
        return executor.execute(self.kernel_wrapper)
