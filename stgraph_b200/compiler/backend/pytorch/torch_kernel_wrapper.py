"""``torch.autograd.Function`` flavour of the kernel wrapper (``backend/pytorch/torch_kernel_wrapper.py:1-5``)."""
import torch

from ..kernel_wrapper import KernelWrapper


class KernelWrapperTorch(KernelWrapper, torch.autograd.Function):
    pass
