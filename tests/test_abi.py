"""The product shared library (built by __graft_entry__.build()) loads without a GPU and exports every symbol that
include/xlprop.h declares; host-only queries work.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from xlumina_b200 import _lib
    return _lib.lib()


def test_header_and_binding_agree(lib):
    from xlumina_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "xlprop.h")).read()
    declared = set(re.findall(r"\b(xl_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)


def test_host_queries(lib):
    hdr = open(os.path.join(ROOT, "include", "xlprop.h")).read()
    assert lib.xl_version() == int(re.search(r"#define XLPROP_VERSION (\d+)", hdr).group(1))
    assert lib.xl_rs_padded_length(2048) == 4096 and lib.xl_rs_padded_length(1024) == 2048
    assert lib.xl_rs_padded_length(1000) == 2048 and lib.xl_rs_padded_length(3000) == 0
    assert lib.xl_czt_padded_length(2048, 2048) == 4096 and lib.xl_czt_padded_length(1024, 400) == 2048
    assert lib.xl_rs_transfer_bytes(2048) == 4096 * 4096 * 8
    assert lib.xl_rs_workspace_bytes(2048, 1, 0) >= 4096 * 2048 * 8
    assert lib.xl_czt_workspace_bytes(2048, 2048, 2048, 0) > 0


def test_product_has_no_cpu_path():
    import torch
    import xlumina_b200 as xb
    from xlumina_b200._lib import XlpropError
    with pytest.raises(XlpropError):
        xb.ops.rs_propagation(torch.zeros(8, 8, dtype=torch.complex64), 1.0, 1.0, 1.0, 1.0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "xlumina_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/ is", ""), f
