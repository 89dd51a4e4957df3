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
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
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


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/coffe_b200.h compiles as C99 (and as C++), and a C program linked against the library gets a clean
    error code -- not a crash, not a fallback -- from coffe_b200_create when there is no GPU"""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    from coffe_b200 import capi
    inc = ROOT / "include"
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include "coffe_b200.h"
int main(void)
{
    coffe_b200_ctx *ctx = NULL;
    coffe_b200_quadrature q = {0};
    coffe_b200_integrals_params ip = {0};
    int n_r = 0, n_r40 = 0, n_ren = 0;
    ip.bessel_bins = 10000;
    if (coffe_b200_integrals_sizes(&ip, &n_r, &n_r40, &n_ren) != COFFE_B200_OK) return 3;
    printf("sizes %d %d %d\n", n_r, n_r40, n_ren);
    int rc = coffe_b200_create(0, &ctx);
    printf("create rc=%d msg=%s\n", rc, coffe_b200_last_error(NULL));
    if (rc == COFFE_B200_OK) { rc = coffe_b200_get_quadrature(ctx, &q); coffe_b200_destroy(ctx); return rc ? 4 : 0; }
    return ctx == NULL ? 0 : 5;
}
''')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", f"-I{inc}", str(src)], check=True)
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", f"-I{inc}", str(src)], check=True)
    exe = tmp_path / "t"
    lib = capi.LIB_PATH
    subprocess.run(["gcc", "-std=c99", f"-I{inc}", str(src), "-o", str(exe), str(lib), f"-Wl,-rpath,{lib.parent}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sizes 10043 297 201" in r.stdout
