"""CPU tier: the product library loads without a GPU, exports every symbol include/edsparser_b200.h
declares, and fails loudly (EDS_ERR_CUDA) instead of falling back when there is no device."""
import ctypes
import os
import re
import subprocess

import pytest

import edsparser_b200
from edsparser_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "edsparser_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(eds_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def product():
    if not os.path.exists(capi.PRODUCT_SO):
        subprocess.check_call(["make", "-C", ROOT, "lib"], stdout=subprocess.DEVNULL)
    return edsparser_b200.load()


def test_header_symbols_are_exported(product):
    names = _declared()
    assert len(names) >= 15
    raw = ctypes.CDLL(capi.PRODUCT_SO)
    for n in names:
        assert hasattr(raw, n), n
    assert sorted(capi.EXPORTS) == names


def test_version(product):
    assert product.version().endswith("sm_100a")


def test_msa_index_is_host_only(product):
    idx = capi.MsaIndex()
    text = b">a\nACGT\nAC\n>b desc\nAC-T\nAC\n"
    rc = product.L.eds_msa_index_host(text, len(text), ctypes.byref(idx))
    assert rc == 0
    assert (idx.n_rows, idx.n_cols, idx.line_width, idx.row_bytes) == (2, 6, 4, 7)
    assert [idx.row_start[i] for i in range(2)] == [3, 19]
    product.L.eds_msa_index_free(ctypes.byref(idx))


def test_no_cpu_fallback(product):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.EdsError) as ei:
        product.context(0)
    assert ei.value.status == capi.EDS_ERR_CUDA
    assert "no CPU fallback" in ei.value.message
