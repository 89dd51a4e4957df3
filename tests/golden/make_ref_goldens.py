"""Condense the reference's own golden vectors (/root/reference/tests/benchmarks/*.dat, the files its
pytest and CTest suites compare against, tests/test_coffe.py:130-430) into one compressed fixture,
tests/golden/reference_benchmarks.npz, so the parity tests can run where /root/reference is absent.

Small files are stored in full; the 10^4-row tables are subsampled (every 20th row for the I^n_l
tables and the background, every 4th row/column of the 201x201 renormalisation grid).
Run here (container with /root/reference):  python tests/golden/make_ref_goldens.py
"""
import re
import sys
from pathlib import Path

import numpy as np

SRC = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/tests/benchmarks")
OUT = Path(__file__).resolve().parent / "reference_benchmarks.npz"

store = {}
for f in sorted(SRC.glob("benchmark_*.dat")):
    name = f.stem[len("benchmark_"):]
    data = np.loadtxt(f)
    if re.fullmatch(r"integral\d", name) and data.shape[0] > 2000:
        data = data[::20]
    elif name == "background":
        data = data[::20]
    elif name == "integral8_renormalization":
        # rows are (chi1, chi2, value) on a 201 x 201 grid
        n = int(round(np.sqrt(data.shape[0])))
        data = data.reshape(n, n, -1)[::4, ::4].reshape(-1, data.shape[1])
    store[name] = data
np.savez_compressed(OUT, **store)
print(f"wrote {OUT} with {len(store)} arrays, {OUT.stat().st_size / 1024:.0f} KiB")
