"""The C-ABI library loads, exports every symbol include/coffe_b200.h declares, and fails loudly
(no CPU fallback) when there is no GPU."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "coffe_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(coffe_b200_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_entry_points():
    names = declared_symbols()
    for must in ("coffe_b200_create", "coffe_b200_destroy", "coffe_b200_upload_tables", "coffe_b200_multipoles",
                 "coffe_b200_corrfunc", "coffe_b200_average_multipoles", "coffe_b200_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from coffe_b200 import capi
    lib = capi.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/coffe_b200.h but not exported"


def test_struct_layouts_match_header():
    """ctypes mirrors have the field order of the C structs"""
    from coffe_b200 import capi
    text = (ROOT / "include" / "coffe_b200.h").read_text()
    body = re.search(r"typedef struct coffe_b200_quadrature \{(.*?)\} coffe_b200_quadrature;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1]
        fields += [n.strip() for n in names.split(",")]
    assert fields == [n for n, _ in capi.Quadrature._fields_]
    body = re.search(r"typedef struct coffe_b200_options \{(.*?)\} coffe_b200_options;", text, re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            fields += [n.strip() for n in decl.split(None, 1)[1].split(",")]
    assert fields == [n for n, _ in capi.Options._fields_]


def test_no_gpu_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import coffe_b200
    with pytest.raises(coffe_b200.CoffeB200Error) as e:
        coffe_b200.Engine(0)
    assert "no CUDA device" in str(e.value) or "ECUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    for f in (ROOT / "coffe_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h", ".cpp") and f.is_file():
            text = f.read_text()
            assert "oracle" not in text.replace("# oracle-free", ""), f"{f} mentions the oracle"
            assert "hostsim" not in text or f.name in ("hd.h", "assemble.h", "item.cuh"), f"{f} mentions hostsim"
