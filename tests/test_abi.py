"""CPU tests of the C-ABI boundary: the library builds, loads and exports every
symbol that include/lightglue_b200.h declares.  No compute call is made."""
import ctypes
import re
from pathlib import Path

import pytest

from helpers import ROOT

HEADER = ROOT / "include" / "lightglue_b200.h"


@pytest.fixture(scope="module")
def libpath():
    from glue_factory_colon_b200 import build

    return build.build()


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(lgb200_\w+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "lgb200_linear" in syms and "lgb200_filter_matches" in syms and len(syms) >= 13


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(str(libpath))
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_binding_table_matches_header(libpath):
    from glue_factory_colon_b200 import _abi

    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    for name, argtypes in _abi.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", text, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(argtypes), f"{name}: header has {len(params)} params, binding {len(argtypes)}"
    lib = _abi.load()
    assert lib.lgb200_abi_version() == _abi.ABI_VERSION
    assert lib.lgb200_error_string(-1).decode().startswith("unsupported")


def test_argument_validation_without_gpu(libpath):
    """Bad arguments are rejected before any CUDA call."""
    from glue_factory_colon_b200 import _abi

    lib = _abi.load()
    assert lib.lgb200_attention(0, None, None, None, 2, 128, None, 0, None, None) == -2
    assert lib.lgb200_pack_rows(1, 1, 10, 255, 0, 128, 1, None, None) == -1
    assert lib.lgb200_filter_matches(None, 1, 5, 5, None, 0.0, None, None, 0, 4, 4, None, None, None, None, None, 0, None) == -2
    # training-side helpers (ABI v6): workspace size of the attention backward, plane split / merge
    n_ws = ctypes.c_longlong(0)
    assert lib.lgb200_attention_bwd_workspace(4, 256, ctypes.byref(n_ws)) == 0
    assert n_ws.value == 2 * 4 * 4 * 256 + 16 + 4 * 4 * 256 * 256  # statistics + scale slot + four fp16 plane pairs
    assert lib.lgb200_attention_bwd_workspace(4, 200, ctypes.byref(n_ws)) == -1 and lib.lgb200_attention_bwd_workspace(4, 256, None) == -2
    assert lib.lgb200_attention_bwd(None, None, None, None, None, 4, 256, None, 0, None, None, None, None, None) == -2
    assert lib.lgb200_split_dynamic(None, 1024, None, None, 0, None, 0, None) == -2
    assert lib.lgb200_merge_rows(None, 1024, 1.0, None, None) == -2


def test_product_path_has_no_cpu_fallback():
    import torch

    from glue_factory_colon_b200 import LightGlue, _abi
    from glue_factory_colon_b200.synthetic import make_pairs

    model = LightGlue({"n_layers": 1}).eval()
    with pytest.raises(_abi.LightGlueB200Error):
        model(make_pairs(1, 16, 16))
    # and nothing under the package imports the oracle
    for f in (ROOT / "glue_factory_colon_b200").glob("*.py"):
        assert "oracle" not in f.read_text().replace("oracle/", "").split("\"\"\"")[-1] or f.name == "lightglue.py"
