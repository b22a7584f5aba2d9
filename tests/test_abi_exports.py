"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/gtb200.h declares, and refuses to work without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gtb200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gtb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import gtb200
    if not os.path.exists(gtb200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return gtb200.load_library()


def test_header_declares_the_path():
    names = declared_functions()
    for must in ("gtb_overlap_count", "gtb_overlap_coverage", "gtb_index_create", "gtb_index_add_queries",
                 "gtb_index_finish", "gtb_scan_create", "gtb_scan_add_reads", "gtb_scan_finish", "gtb_scan_fetch"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    import gtb200
    names = declared_functions()
    assert sorted(gtb200.EXPORTS) == names, "python binding and header disagree"
    for name in names:
        assert hasattr(lib, name), name
    assert lib.gtb_abi_version() == 1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    h = ctypes.c_void_p()
    assert lib.gtb_ctx_create(0, ctypes.byref(h)) == 100      # GTB_ERR_NO_DEVICE
    assert not h.value


def test_product_does_not_reference_oracle():
    """The oracle is test infrastructure: nothing under the product tree may mention it."""
    pkg = os.path.join(ROOT, "ibm-cbc-genomic-tools_b200")
    bad = []
    for d, _dirs, files in os.walk(pkg):
        if os.sep + "build" in d or os.sep + "lib" in d or os.sep + "bin" in d or "__pycache__" in d:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py", "Makefile")):
                if re.search(r"oracle|liboracle|orc_", open(os.path.join(d, f), errors="ignore").read()):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
