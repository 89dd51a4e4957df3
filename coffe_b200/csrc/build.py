"""Build libcoffe_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
SOURCES = ["engine.cu", "covariance.cu", "integrals.cu", "growth_host.cpp"]
HEADERS = ["hd.h", "tables.cuh", "physics.cuh", "quadrature.cuh", "item.cuh", "di_fast.cuh", "assemble.h", "compact.cuh", "compact_build.h", "table_build.h", "rules_build.h", "ctx.h", "special.cuh", "fft.cuh", "group.inc",
           "../../include/coffe_b200.h"]
OUT = HERE / "libcoffe_b200.so"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def up_to_date():
    if not OUT.exists():
        return False
    t = OUT.stat().st_mtime
    return all((HERE / f).stat().st_mtime <= t for f in SOURCES + HEADERS + ["build.py"])


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = ["nvcc", *NVCC_FLAGS, "-o", str(OUT), *[str(HERE / s) for s in SOURCES], "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    # registers / spills / shared memory per kernel, kept in the tree as evidence; compile times dropped so the file is stable
    (HERE / "ptxas_info.txt").write_text("".join(l for l in r.stderr.splitlines(keepends=True) if "Compile time" not in l))
    if verbose or r.returncode:
        sys.stderr.write(r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building libcoffe_b200.so")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(OUT)
