"""The C-ABI library loads on a CPU-only box and exports every symbol include/schnorr_b200.h declares
(no compute calls without a GPU); the product path fails loudly when no CUDA device exists."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "schnorr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(schnorr_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    import schnorr_sig_b200 as s
    names = declared_symbols()
    assert len(names) >= 20
    so = s._lib.library_path()
    assert os.path.dirname(so).endswith(os.path.join("schnorr-sig_b200", "csrc"))    # built in-tree
    lib = ctypes.CDLL(s.build.build())
    for name in names:
        assert hasattr(lib, name), "library does not export %s" % name
    assert sorted(s._lib.EXPORTED_SYMBOLS) == names      # the Python binding covers the whole header


def test_every_entry_point_cites_the_reference():
    src = open(os.path.join(ROOT, "include", "schnorr_b200.h")).read()
    for ref in ("src/signature.rs:181-205", "src/signature.rs:274-306", "src/batch.rs:31-50", "src/public.rs:26-32",
                "src/signature.rs:114-129", "src/public.rs:54-56", "src/error.rs:13-18"):
        assert ref in src


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import schnorr_sig_b200 as s
    with pytest.raises(s.EngineError, match="no CPU fallback"):
        s.Engine(0)
    h = ctypes.c_void_p()
    assert s._lib.lib().schnorr_b200_create(0, ctypes.byref(h)) == -3 and not h.value


def test_product_does_not_import_the_oracle():
    """Nothing under schnorr-sig_b200/ may reference oracle/ (checked textually)."""
    pkg = os.path.join(ROOT, "schnorr-sig_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import cref" not in text and "import pyref" not in text and "libcref" not in text
                assert "oracle/cref" not in text, f
