"""CPU test of the host-side cache that decides when a static CSR runs the packed kernel (``CSR.packed_meta``,
``graph/static/csr.py``): keying by (address, version), hits on fresh views of the same storage, repack after an
in-place update, two live entries, no packing during CUDA-graph capture, and the fallback for scales that change
every call.  The pack kernel itself is replaced by a counter here; its arithmetic is covered by the GPU tests."""
import pytest
import torch

from stgraph_b200 import kernels
from stgraph_b200.graph.static import csr as csr_mod


@pytest.fixture
def csr(monkeypatch):
    packs = []

    def fake_pack(view, nbr_scale=None, edge_scale=None, out=None, device=None):
        packs.append((None if nbr_scale is None else nbr_scale.data_ptr(), None if edge_scale is None else edge_scale.data_ptr()))
        return torch.zeros(4, 2, dtype=torch.int32)

    monkeypatch.setattr(kernels, "pack_edge_meta", fake_pack)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    ro = torch.tensor([0, 2, 3, 4], dtype=torch.int32)
    col = torch.tensor([1, 2, 0, 1], dtype=torch.int32)
    c = csr_mod.CSR(ro, col, torch.arange(4, dtype=torch.int32), torch.arange(3, dtype=torch.int32),
                    torch.tensor([2, 1, 1], dtype=torch.int32), torch.tensor([1, 2, 1], dtype=torch.int32), eids_identity=True)
    c.pack_enabled = True
    c.packs = packs
    return c


def test_hit_on_a_fresh_view_and_repack_after_inplace_update(csr):
    norm = torch.rand(3, 1)
    m1 = csr.packed_meta(norm.reshape(-1), None)
    m2 = csr.packed_meta(norm.detach().reshape(-1), None)          # new tensor objects, same storage and version
    assert m1 is m2 and len(csr.packs) == 1
    norm.mul_(2.0)                                                 # in-place: version counter moves -> repack
    m3 = csr.packed_meta(norm.reshape(-1), None)
    assert m3 is not m1 and len(csr.packs) == 2


def test_two_entries_and_keepalive(csr):
    norm = torch.rand(3)
    w = torch.rand(4)
    a = csr.packed_meta(norm, None)
    b = csr.packed_meta(norm, w)                                   # weighted and unweighted layers share a graph
    assert csr.packed_meta(norm, None) is a and csr.packed_meta(norm, w) is b and len(csr.packs) == 2
    other = torch.rand(3)
    csr.packed_meta(other, None)                                   # third key evicts the oldest entry only
    assert len(csr._meta_cache) == 2 and len(csr.packs) == 3
    # an entry holds its scale tensors, so their storage cannot be recycled under the key
    assert any(k[0] is other for _, k, _ in csr._meta_cache)


def test_nothing_to_pack_and_disabled(csr, monkeypatch):
    assert csr.packed_meta(None, None) is None and not csr.packs
    monkeypatch.setattr(csr_mod, "PACK_META", False)
    assert csr.packed_meta(torch.rand(3), None) is None and not csr.packs


def test_no_packing_during_graph_capture(csr, monkeypatch):
    norm = torch.rand(3)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: True)
    assert csr.packed_meta(norm, None) is None and not csr.packs   # a miss while capturing: plain kernel
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    m = csr.packed_meta(norm, None)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: True)
    assert csr.packed_meta(norm, None) is m                        # a hit while capturing is fine


def test_scales_that_change_every_call_stop_packing(csr):
    for _ in range(csr_mod.MAX_META_REPACKS):
        assert csr.packed_meta(torch.rand(3), None) is not None
    assert csr.packed_meta(torch.rand(3), None) is None
    assert csr.pack_enabled is False and csr._meta_cache == []


def test_invalidate_after_a_write_torch_does_not_see(csr):
    """A raw write (storage changed without a version bump) needs an explicit invalidate: the key is (address, version)."""
    norm = torch.rand(3)
    a = csr.packed_meta(norm, None)
    norm.untyped_storage().copy_(torch.rand(3).untyped_storage())      # same address, same version counter
    assert csr.packed_meta(norm, None) is a and len(csr.packs) == 1    # the documented caveat
    csr.invalidate_packed_meta()
    b = csr.packed_meta(norm, None)
    assert b is not a and len(csr.packs) == 2 and csr.pack_enabled
