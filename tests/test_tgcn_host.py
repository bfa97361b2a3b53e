"""Host-side logic of the TGCN module that needs no GPU: parameter packing (ops_tgcn.pack_parameters) against the
module's own parameters, and that the fused cell's parameter pack never travels with a copy / pickle of the module."""
import copy
import io

import torch

from stgraph_b200.nn.pytorch import TGCN
from stgraph_b200.ops_tgcn import pack_parameters


def _mods(cell):
    return cell.conv_z, cell.conv_r, cell.conv_h, cell.linear_z, cell.linear_r, cell.linear_h


def test_pack_parameters_layout_and_gradients():
    """W3 / b3 / La / Lc_zr / Lc_h / lb hold the reference cell's parameters (tgcn.py:16-47) in the block layout, and a
    gradient on the pack reaches the module's own parameters."""
    torch.manual_seed(0)
    cell = TGCN(5, 4)
    for c in (cell.conv_z, cell.conv_r, cell.conv_h):
        torch.nn.init.normal_(c.bias)
    W3, b3, La, Lc_zr, Lc_h, lb = pack_parameters(*_mods(cell))
    hid = 4
    assert W3.shape == (5, 12) and b3.shape == (12,) and La.shape == (3, 4, 4) and Lc_zr.shape == (4, 8) and Lc_h.shape == (4, 4)
    x, h = torch.randn(7, 5), torch.randn(7, hid)
    conv = [x @ c.weight + c.bias for c in (cell.conv_z, cell.conv_r, cell.conv_h)]          # no graph: identity aggregation
    hb = x @ W3 + b3
    for g in range(3):
        torch.testing.assert_close(hb[:, g * hid:(g + 1) * hid], conv[g])
    # linear_g(cat(a, c)) == a @ La[g] + c @ Lc_g + lb_g
    for g, lin in enumerate((cell.linear_z, cell.linear_r, cell.linear_h)):
        ref = lin(torch.cat((conv[g], h), 1))
        Lc = Lc_zr[:, g * hid:(g + 1) * hid] if g < 2 else Lc_h
        torch.testing.assert_close(conv[g] @ La[g] + h @ Lc + lb[g * hid:(g + 1) * hid], ref, rtol=1e-5, atol=1e-6)
    (W3.sum() + 2 * b3.sum() + 3 * La.sum() + 4 * Lc_zr.sum() + 5 * Lc_h.sum() + 6 * lb.sum()).backward()
    assert torch.all(cell.conv_r.weight.grad == 1) and torch.all(cell.conv_h.bias.grad == 2)
    assert torch.all(cell.linear_z.weight.grad[:, :hid] == 3) and torch.all(cell.linear_r.weight.grad[:, hid:] == 4)
    assert torch.all(cell.linear_h.weight.grad[:, hid:] == 5) and torch.all(cell.linear_z.bias.grad == 6)


def test_parameter_pack_cache_reuse_and_invalidation():
    torch.manual_seed(1)
    cell = TGCN(3, 4)
    a = cell._packed_parameters()
    assert cell._packed_parameters()[0] is a[0]                      # reused inside a window
    with torch.no_grad():
        cell.conv_z.weight.add_(1.0)                                 # an optimizer step: in-place update
    b = cell._packed_parameters()
    assert b[0] is not a[0] and torch.equal(b[0][:, :4], cell.conv_z.weight)
    b[0].sum().backward()                                            # a backward pass through the pack marks it stale
    assert cell._packed_parameters()[0] is not b[0]
    with torch.no_grad():
        c = cell._packed_parameters()
    assert not c[0].requires_grad
    assert cell._packed_parameters()[0].requires_grad                # grad mode is part of the key


def test_parameter_pack_does_not_travel_with_copies():
    torch.manual_seed(2)
    cell = TGCN(3, 4)
    cell._packed_parameters()
    assert "_pack_cache" in cell.__dict__
    twin = copy.deepcopy(cell)                                       # non-leaf tensors cannot be deep-copied
    assert "_pack_cache" not in twin.__dict__
    buf = io.BytesIO()
    torch.save(cell, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    assert "_pack_cache" not in back.__dict__
    for (k, p), (_, q) in zip(cell.state_dict().items(), back.state_dict().items()):
        assert torch.equal(p, q), k
    assert set(cell.state_dict()) == {f"{m}.{t}" for m in ("conv_z", "conv_r", "conv_h", "linear_z", "linear_r", "linear_h")
                                      for t in ("weight", "bias")}


def test_layers_can_be_deep_copied_and_pickled():
    """GCNConv / GATConv own an STGraph (trace + executor caches, a backend holding the torch module): a copy of the layer
    starts with an empty one and the same parameters."""
    from stgraph_b200.nn.pytorch import GATConv, GCNConv

    for layer in (GCNConv(6, 3), GATConv(6, 4, 2)):
        twin = copy.deepcopy(layer)
        assert twin.stgraph is not layer.stgraph and twin.stgraph._ctx_map == {}
        buf = io.BytesIO()
        torch.save(layer, buf)
        buf.seek(0)
        back = torch.load(buf, weights_only=False)
        for (k, p), (_, q), (_, r) in zip(layer.state_dict().items(), twin.state_dict().items(), back.state_dict().items()):
            assert torch.equal(p, q) and torch.equal(p, r), k


def test_parameter_pack_goes_stale_with_partly_frozen_parameters():
    """Frozen convolutions: the packed W3 / b3 get no gradient, the hook on the linear halves still marks the pack stale."""
    torch.manual_seed(3)
    cell = TGCN(3, 4)
    for c in (cell.conv_z, cell.conv_r, cell.conv_h):
        c.weight.requires_grad_(False)
        c.bias.requires_grad_(False)
    a = cell._packed_parameters()
    assert not a[0].requires_grad and a[2].requires_grad
    (a[2].sum() + a[5].sum()).backward()
    assert cell._packed_parameters()[2] is not a[2]
    assert cell.linear_z.weight.grad is not None and cell.conv_z.weight.grad is None
