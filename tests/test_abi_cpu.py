"""CPU: the C-ABI shared library loads without a GPU and exports every entry point include/lrce_b200.h declares, and the
ctypes binding (vqa-lrce-kbs-2023_b200/_lib.py) covers exactly that set. No compute call is made here."""
import ctypes
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lrce_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lrce_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import lrce_b200

    names = declared_symbols()
    assert len(names) >= 30
    handle = ctypes.CDLL(lrce_b200._lib.LIB_PATH)
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing
    assert sorted(lrce_b200._lib.exported_symbols()) == names      # the binding and the header agree
    assert lrce_b200._lib.lib().lrce_abi_version() >= 1            # host-only call


def test_compute_entry_points_fail_loudly_without_a_gpu():
    import torch

    import lrce_b200

    if torch.cuda.is_available():
        return
    try:
        lrce_b200.ops.layernorm(torch.zeros(4, 128, dtype=torch.bfloat16), torch.ones(128), torch.zeros(128), 1e-5)
    except lrce_b200.LrceError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("a CPU tensor must be rejected")
