"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports what the header declares."""
import ctypes
import os
import re

from stgraph_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "stgraph_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stg_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"


def test_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/stgraph_b200.h but not exported"


def test_python_binding_covers_header():
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols()


def test_load_and_version():
    lib = _lib.load()
    assert lib.stg_abi_version() == _lib.ABI_VERSION
    assert lib.stg_last_error() is not None


def test_struct_layouts_match_header():
    # sizes implied by the C declarations (LP64)
    assert ctypes.sizeof(_lib.StgCsrView) == 4 * 8 + 4 * 4 + 2 * 8 + 2 * 4 + 8   # + work_queue (since ABI 2)
    assert ctypes.sizeof(_lib.StgVmInstr) == 16
    assert ctypes.sizeof(_lib.StgVmTensor) == 16
    assert ctypes.sizeof(_lib.StgVmProgram) == 8 * 4 + 8 * _lib.VM_MAX_ACC + 16 * _lib.VM_MAX_TENSORS + 16 * _lib.VM_MAX_INSTR


def test_argument_validation_without_gpu():
    """Entry points reject bad arguments before touching the device (no compute call is made)."""
    lib = _lib.load()
    rc = lib.stg_agg_scaled_sum_f32(None, None, 4, None, None, None, None, None)
    assert rc == -1
    assert b"NULL" in lib.stg_last_error()
    v = _lib.StgCsrView()
    v.num_nodes = 4
    v.num_edges = 0
    v.row_offset = 1  # never dereferenced: feat is rejected first
    rc = lib.stg_agg_scaled_sum_f32(ctypes.byref(v), None, 0, None, None, None, None, None)
    assert rc == -1 and b"feat" in lib.stg_last_error()


def test_gemm_tn_workspace_query_and_argument_checks_without_gpu():
    """stg_gemm_tn_workspace_bytes is a host-side query: slabs x (K*Nc + Nc) floats, none for small M; the entry point
    rejects bad shapes / a missing workspace before any launch."""
    lib = _lib.load()
    for m, k, nc in ((1_000_000, 64, 64), (2_449_029, 100, 47), (1_000_000, 32, 192), (50_000, 8, 48)):
        b = lib.stg_gemm_tn_workspace_bytes(m, k, nc)
        per_slab = (k * nc + nc) * 4
        assert b > 0 and b % per_slab == 0
        slabs = b // per_slab
        assert 2 <= slabs <= 2 * 148 and m // slabs >= 128          # about two CTAs per SM, at least 128 rows per slab
    assert lib.stg_gemm_tn_workspace_bytes(100, 64, 64) == 0           # one slab: the result is written directly
    assert lib.stg_gemm_tn_workspace_bytes(0, 64, 64) == 0
    assert lib.stg_gemm_tn_f32(None, 64, None, 64, 10, 0, 64, None, None, None, 0, None) == -1
    assert b"shape" in lib.stg_last_error()
    assert lib.stg_gemm_tn_f32(1, 8, 1, 64, 10, 64, 64, 1, None, None, 0, None) == -1       # lda < K
    assert b"leading" in lib.stg_last_error()
    assert lib.stg_gemm_tn_f32(1, 64, 1, 64, 1_000_000, 64, 64, 1, None, None, 0, None) == -1   # workspace missing
    assert b"workspace" in lib.stg_last_error()
