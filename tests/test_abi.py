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
