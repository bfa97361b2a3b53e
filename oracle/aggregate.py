"""Aggregation oracle (torch CPU ``index_add_``) for the vertex programs.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference evaluates a vertex program with one thread per (row, feature
lane) and a sequential loop over the row's edges
(``/root/reference/stgraph/compiler/code_gen/templates/fa/tpl_fa_csr_unsorted.jinja:20-44``).
The oracle states the same sums with ``index_add_`` over the edge list.  By
default it accumulates in fp64 and rounds once to fp32 (the check tolerance is
rel 1e-5, SURVEY.md section 8 trap T9); ``dtype=torch.float32`` gives the
"torch-CPU index_add implementation" that ``BASELINE.json`` names as the CPU
baseline.
"""
from __future__ import annotations

import numpy as np
import torch


def _t(x, dtype=None):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        x = x.to(dtype)
    return x


def csr_to_coo(row_offset, column_indices):
    """Expand CSR rows to a per-edge row index (int64)."""
    row_offset = _t(row_offset, torch.int64)
    column_indices = _t(column_indices, torch.int64)
    n = row_offset.shape[0] - 1
    rows = torch.repeat_interleave(torch.arange(n, dtype=torch.int64), row_offset[1:] - row_offset[:-1])
    return rows, column_indices


def scaled_sum(row_offset, column_indices, eids, x, nbr_scale=None, edge_scale=None,
               row_scale=None, eid_base: int = 0, dtype=torch.float64):
    """``out[r,:] = row_scale[r] * sum_{e in row r} nbr_scale[c_e] * edge_scale[eid_e] * x[c_e,:]``.

    With the in-edge CSR this is the GCN forward kernel ``K0``
    (``gcn_conv.py:162-166`` / ``169-182``); with
    the out-edge CSR and ``x = grad_out`` it is the backward kernel ``K1``
    (a gather as well, no atomics).  ``eid_base`` is 1 for PCSR/GPMA views
    (``tpl_fa_pcsr.jinja:32-34``).
    """
    rows, cols = csr_to_coo(row_offset, column_indices)
    x = _t(x)
    out_dtype = x.dtype
    n = _t(row_offset).shape[0] - 1
    feat_shape = x.shape[1:]
    xf = x.reshape(x.shape[0], -1).to(dtype)
    msg = xf[cols]
    if nbr_scale is not None:
        msg = msg * _t(nbr_scale).reshape(-1, 1).to(dtype)[cols]
    if edge_scale is not None:
        e = _t(eids, torch.int64) - eid_base
        msg = msg * _t(edge_scale).reshape(-1, 1).to(dtype)[e]
    out = torch.zeros((n, xf.shape[1]), dtype=dtype)
    out.index_add_(0, rows, msg)
    if row_scale is not None:
        out = out * _t(row_scale).reshape(-1, 1).to(dtype)
    return out.to(out_dtype).reshape((n,) + tuple(feat_shape))


def gcn_forward(fwd_csr, h, norm, edge_weight=None, **kw):
    """``out[v] = norm[v] * sum_{u in in(v)} norm[u] * w[eid] * h[u]`` (``gcn_conv.py:162-182``)."""
    return scaled_sum(fwd_csr.row_offset, fwd_csr.column_indices, fwd_csr.eids, h,
                      nbr_scale=norm, edge_scale=edge_weight, row_scale=norm, **kw)


def gcn_backward(bwd_csr, grad_out, norm, edge_weight=None, **kw):
    """``dh[u] = norm[u] * sum_{v in out(u)} norm[v] * w[eid] * dout[v]`` (SURVEY.md A.1/A.2 ``K1``)."""
    return scaled_sum(bwd_csr.row_offset, bwd_csr.column_indices, bwd_csr.eids, grad_out,
                      nbr_scale=norm, edge_scale=edge_weight, row_scale=norm, **kw)


# --------------------------------------------------------------------------
# GAT
# --------------------------------------------------------------------------
def _leaky(x, slope):
    return torch.where(x > 0, x, x * slope)


def gat_softmax_forward(fwd_csr, el, er, feat, slope=0.2, dtype=torch.float64):
    """Genuine edge-softmax GAT: ``alpha = softmax_e(lrelu(el[u]+er[v]))``, ``out[v] = sum alpha * feat[u]``.

    Closed-form oracle for the fused online-softmax kernel (not the stock
    ``GATConv`` trace, which degenerates to a mean - trap T2).
    ``el, er``: ``[N,H,1]`` (or ``[N,H]``), ``feat``: ``[N,H,D]``.
    Returns ``(out [N,H,D], row_max [N,H], row_sum [N,H])`` where ``row_sum``
    is the sum of ``exp(score - row_max)``.
    """
    rows, cols = csr_to_coo(fwd_csr.row_offset, fwd_csr.column_indices)
    feat = _t(feat)
    n, h, d = feat.shape
    elf = _t(el).reshape(n, h).to(dtype)
    erf = _t(er).reshape(n, h).to(dtype)
    score = _leaky(elf[cols] + erf[rows], slope)                     # [E,H]
    m = torch.full((n, h), -float("inf"), dtype=dtype)
    m.index_reduce_(0, rows, score, "amax", include_self=True)
    p = torch.exp(score - m[rows])
    s = torch.zeros((n, h), dtype=dtype).index_add_(0, rows, p)
    alpha = p / s[rows]
    out = torch.zeros((n, h, d), dtype=dtype)
    out.index_add_(0, rows, alpha.unsqueeze(-1) * feat.to(dtype)[cols])
    m_out = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    return out.to(feat.dtype), m_out.to(feat.dtype), s.to(feat.dtype)


def gat_softmax_backward(fwd_csr, el, er, feat, grad_out, slope=0.2, dtype=torch.float64):
    """Gradients of :func:`gat_softmax_forward` w.r.t. ``feat, el, er`` via torch autograd (fp64)."""
    feat64 = _t(feat).to(dtype).clone().requires_grad_(True)
    n, h, d = feat64.shape
    el64 = _t(el).reshape(n, h).to(dtype).clone().requires_grad_(True)
    er64 = _t(er).reshape(n, h).to(dtype).clone().requires_grad_(True)
    rows, cols = csr_to_coo(fwd_csr.row_offset, fwd_csr.column_indices)
    score = torch.nn.functional.leaky_relu(el64[cols] + er64[rows], slope)
    m = torch.full((n, h), -float("inf"), dtype=dtype)
    m = m.index_reduce(0, rows, score.detach(), "amax", include_self=True)
    p = torch.exp(score - m[rows])
    s = torch.zeros((n, h), dtype=dtype).index_add(0, rows, p)
    alpha = p / s[rows]
    out = torch.zeros((n, h, d), dtype=dtype).index_add(0, rows, alpha.unsqueeze(-1) * feat64[cols])
    out.backward(_t(grad_out).to(dtype))
    f32 = _t(feat).dtype
    return feat64.grad.to(f32), el64.grad.to(f32), er64.grad.to(f32)


def gat_stock_forward(fwd_csr, el, er, feat, slope=0.2, dtype=torch.float64):
    """Stock ``GATConv`` exactly as the reference traces it (trap T2).

    ``gat_conv.py:48-56``: Python's ``max`` on a one-element list returns the
    element, so the trace is ``V1 = Sub(V0, V0)``; ``V3 = exp(lrelu(V1))``
    (= 1 for finite inputs), ``V4 = AggSum(V3)``, ``V7 = AggSum(V3/V4 * feat)``
    (SURVEY.md appendix A.3).  Returns ``(out, V3 [E,H,1], V4 [N,H,1])``.
    """
    rows, cols = csr_to_coo(fwd_csr.row_offset, fwd_csr.column_indices)
    feat = _t(feat)
    n, h, d = feat.shape
    elf = _t(el).reshape(n, h, 1).to(dtype)
    erf = _t(er).reshape(n, h, 1).to(dtype)
    eids = _t(fwd_csr.eids, torch.int64)
    v0 = elf[cols] + erf[rows]
    v1 = v0 - v0
    v3_csr = torch.exp(_leaky(v1, slope))
    v4 = torch.zeros((n, h, 1), dtype=dtype).index_add_(0, rows, v3_csr)
    v5 = v3_csr / v4[rows]
    out = torch.zeros((n, h, d), dtype=dtype).index_add_(0, rows, v5 * feat.to(dtype)[cols])
    v3 = torch.zeros_like(v3_csr)
    v3[eids] = v3_csr
    return out.to(feat.dtype), v3.to(feat.dtype), v4.to(feat.dtype)


def gat_stock_backward(fwd_csr, el, er, feat, grad_out, slope=0.2, dtype=torch.float64, return_mag=False):
    """Backward of the stock trace with the *reference's* gradient rules.

    ``registry.py:210-213`` gives ``Sub`` a +1 gradient for both operands, so
    although the forward does not depend on ``el``/``er`` the reference still
    emits ``d_el[u] = sum_e V25``, ``d_er[v] = sum_e V25`` with
    ``V25 = ((dout*feat)/V4 - (dout/V4)*out) * V3 * lrelu'(V1)`` summed over D
    (``V24 = 0.2`` since ``V1 = 0``).  ``V0`` feeds both
    operands of ``Sub`` but ``Stmt.grad`` collects per-operand gradients in a
    ``dict`` keyed by the variable (``program.py:328-329``), so only ONE of the two
    +1 contributions survives: ``dV0 = dV1``, not ``2*dV1``.
    Returns ``(d_feat [N,H,D], d_el [N,H,1], d_er [N,H,1])``.
    """
    rows, cols = csr_to_coo(fwd_csr.row_offset, fwd_csr.column_indices)
    feat = _t(feat)
    n, h, d = feat.shape
    g = _t(grad_out).to(dtype)
    f = feat.to(dtype)
    out, v3e, v4 = gat_stock_forward(fwd_csr, el, er, feat, slope, dtype)
    out = out.to(dtype)
    v4 = v4.to(dtype)
    elf = _t(el).reshape(n, h, 1).to(dtype)
    erf = _t(er).reshape(n, h, 1).to(dtype)
    v0 = elf[cols] + erf[rows]
    v1 = v0 - v0
    v3 = torch.exp(_leaky(v1, slope))
    v5 = v3 / v4[rows]
    d_feat = torch.zeros((n, h, d), dtype=dtype).index_add_(0, cols, g[rows] * v5)
    # d(out)/d(V3) path: V5 = V3 / V4 ; V4 = AggSum(V3)
    v15 = (g[rows] * f[cols]) * (1.0 / v4[rows])            # dV5 * 1/V4
    v17 = (g[rows] / v4[rows]) * out[rows]                  # through V4 = sum V3
    v22 = v15 - v17
    v23 = v22 * v3                                          # exp'
    v24 = torch.where(v1 > 0, torch.ones_like(v1), torch.full_like(v1, slope))
    v25 = (v23 * v24).sum(dim=-1, keepdim=True)             # [E,H,1]
    # Sub(V0, V0): +1 for both operands, but the dict keeps one contribution; Add passes it to el and er
    d_v0 = v25
    d_el = torch.zeros((n, h, 1), dtype=dtype).index_add_(0, cols, d_v0)
    d_er = torch.zeros((n, h, 1), dtype=dtype).index_add_(0, rows, d_v0)
    t = feat.dtype
    if return_mag:
        # sum of |terms| of each reduction: d_er is an exact zero in real arithmetic (the softmax
        # weights sum to one), so only this magnitude gives a meaningful tolerance
        a25 = ((v15.abs() + v17.abs()) * v3 * v24.abs()).sum(dim=-1, keepdim=True)
        m_feat = torch.zeros((n, h, d), dtype=dtype).index_add_(0, cols, (g[rows] * v5).abs())
        m_el = torch.zeros((n, h, 1), dtype=dtype).index_add_(0, cols, a25)
        m_er = torch.zeros((n, h, 1), dtype=dtype).index_add_(0, rows, a25)
        return d_feat.to(t), d_el.to(t), d_er.to(t), m_feat, m_el, m_er
    return d_feat.to(t), d_el.to(t), d_er.to(t)


# --------------------------------------------------------------------------
# tolerance helper (trap T9)
# --------------------------------------------------------------------------
def assert_close_rel(actual, expected, rel=1e-5, abs_terms=None, what=""):
    """``|x - ref| <= rel * max(|ref|, scale)`` where ``scale`` bounds the summed magnitudes.

    ``abs_terms`` (same shape as ``expected``) is ``sum |terms|`` of the
    reduction that produced each element; cancellation can make ``|ref|`` tiny
    while the fp32 rounding error stays proportional to ``sum |terms|``.
    """
    a = _t(actual).double().reshape(-1)
    e = _t(expected).double().reshape(-1)
    assert a.shape == e.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(e.shape)}"
    scale = e.abs()
    if abs_terms is not None:
        scale = torch.maximum(scale, _t(abs_terms).double().reshape(-1))
    scale = torch.clamp(scale, min=1e-30)
    err = (a - e).abs() / scale
    bad = torch.nonzero(~(err <= rel) , as_tuple=False)
    if bad.numel():
        i = int(bad[0])
        raise AssertionError(
            f"{what}: {bad.shape[0]} / {a.numel()} elements exceed rel {rel}: first idx {i} "
            f"actual {a[i].item():.9g} expected {e[i].item():.9g} scale {scale[i].item():.3g} "
            f"(max rel err {err.max().item():.3g})")
