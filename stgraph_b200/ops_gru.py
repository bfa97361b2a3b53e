"""Element-wise pieces of the TGCN (GRU) cell as fused torch autograd ops (``csrc/gates.cu``).

``stgraph/nn/pytorch/temporal/tgcn.py:21-47`` spends about sixteen element-wise kernels per step forward and thirty
backward on ``[N, H]`` tensors; these three ops do the same arithmetic in three passes each way (the ``Linear``
layers between them remain cuBLAS GEMMs driven by torch autograd):

* :func:`bias_clamp`   ``clamp(a + bias, lo, hi)`` (in place on ``a``: GCNConv's bias add + the cell's ``clamp(+-1e6)``)
* :func:`gru_reset`    ``h * sigmoid(pr)``
* :func:`gru_update`   ``z * h + (1 - z) * tanh(ph)`` with ``z = sigmoid(pz)``

Backward recomputes the activations from the saved pre-activations.  Sync-free and allocation-free on the C side, so a
whole BPTT window stays capturable in a CUDA graph.
"""
from __future__ import annotations

import torch

from . import _lib, kernels


def _c(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("TGCN gate tensors must live on a CUDA device (stgraph_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError(f"TGCN gate tensors must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


class _BiasClamp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, bias, lo, hi):
        a = _c(a)
        rows = a.shape[0]
        cols = a.numel() // max(rows, 1)
        b = _c(bias) if bias is not None else None
        _lib.call("stg_bias_clamp_f32", a.data_ptr(), _lib.ptr(b), rows, cols, float(lo), float(hi), _lib.current_stream_ptr())
        kernels.launch_count += 1
        ctx.mark_dirty(a)
        ctx.bounds, ctx.has_bias = (float(lo), float(hi)), bias is not None
        ctx.save_for_backward(a)
        return a

    @staticmethod
    def backward(ctx, d_y):
        (y,) = ctx.saved_tensors
        d_y = _c(d_y)
        d_a = torch.empty_like(d_y)
        _lib.call("stg_clamp_bwd_f32", y.data_ptr(), d_y.data_ptr(), d_a.data_ptr(), d_y.numel(), ctx.bounds[0], ctx.bounds[1],
                  _lib.current_stream_ptr())
        kernels.launch_count += 1
        d_b = d_a.reshape(d_a.shape[0], -1).sum(0) if ctx.has_bias else None
        return d_a, d_b, None, None


class _GruReset(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pr, h):
        pr, h = _c(pr), _c(h)
        hr = torch.empty_like(h)
        _lib.call("stg_gru_reset_fwd_f32", pr.data_ptr(), h.data_ptr(), hr.data_ptr(), h.numel(), _lib.current_stream_ptr())
        kernels.launch_count += 1
        ctx.save_for_backward(pr, h)
        return hr

    @staticmethod
    def backward(ctx, d_hr):
        pr, h = ctx.saved_tensors
        d_hr = _c(d_hr)
        d_pr, d_h = torch.empty_like(pr), torch.empty_like(h)
        _lib.call("stg_gru_reset_bwd_f32", pr.data_ptr(), h.data_ptr(), d_hr.data_ptr(), d_pr.data_ptr(), d_h.data_ptr(),
                  h.numel(), _lib.current_stream_ptr())
        kernels.launch_count += 1
        return d_pr, d_h


class _GruUpdate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pz, ph, h):
        pz, ph, h = _c(pz), _c(ph), _c(h)
        out = torch.empty_like(h)
        _lib.call("stg_gru_update_fwd_f32", pz.data_ptr(), ph.data_ptr(), h.data_ptr(), out.data_ptr(), h.numel(),
                  _lib.current_stream_ptr())
        kernels.launch_count += 1
        ctx.save_for_backward(pz, ph, h)
        return out

    @staticmethod
    def backward(ctx, d_out):
        pz, ph, h = ctx.saved_tensors
        d_out = _c(d_out)
        d_pz, d_ph, d_h = torch.empty_like(pz), torch.empty_like(ph), torch.empty_like(h)
        _lib.call("stg_gru_update_bwd_f32", pz.data_ptr(), ph.data_ptr(), h.data_ptr(), d_out.data_ptr(), d_pz.data_ptr(),
                  d_ph.data_ptr(), d_h.data_ptr(), h.numel(), _lib.current_stream_ptr())
        kernels.launch_count += 1
        return d_pz, d_ph, d_h


def bias_clamp(a: torch.Tensor, bias, lo: float, hi: float) -> torch.Tensor:
    """``clamp(a + bias, lo, hi)`` written into ``a`` (which must be a fresh intermediate: it is modified in place)."""
    return _BiasClamp.apply(a, bias, lo, hi)


def gru_reset(pr: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """``h * sigmoid(pr)`` -- the reset gate applied to the hidden state (``tgcn.py:33-41``)."""
    return _GruReset.apply(pr, h)


def gru_update(pz: torch.Tensor, ph: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """``z * h + (1 - z) * tanh(ph)``, ``z = sigmoid(pz)`` -- update gate + candidate state (``tgcn.py:43-47``)."""
    return _GruUpdate.apply(pz, ph, h)
