#!/usr/bin/env python
"""How far the rules in force sit from the 1e-5 gate on the dense subsample of cfg5 that
tests/test_gpu_parity.py::test_cfg5_dense_self_convergence_default_vs_doubled_rules checks: prints max / 99.9th percentile / median of
|default - fine| relative to the l = 0..4 scale of the point (GPU)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, coffe_b200
import test_gpu_parity as T
from coffe_b200.default_tables import load_default_tableset
eng = coffe_b200.Engine(0); eng.upload_tables(load_default_tableset()); eng.set_options(T.ALL)
z, sep, ls = T.cfg5_grid()
pick = set(range(0, 1000, 2)) | {999}
switches = [16.0, 32.0, 64.0, 75.0, 96.0, 128.0, 150.0, 160.0, 192.0, 224.0, 225.0, 240.0, 256.0, 300.0, 375.0, 384.0] + [240.0 + 12.0 * k for k in range(1, 18)]
for s in switches:
    j = int(np.searchsorted(sep, s)); pick |= {k for k in range(j - 2, j + 2) if 0 <= k < 1000}
idx = np.array(sorted(pick))
base = eng.multipoles(z, sep[idx], ls)
eng.set_quadrature(**T.FINE_RULES)
fine = np.concatenate([eng.multipoles(z[i:i + 4], sep[idx], ls) for i in range(0, len(z), 4)])
scale = np.max(np.abs(fine), axis=1, keepdims=True)
err = np.abs(base - fine) / scale
w = np.unravel_index(np.argmax(err), err.shape)
print("points", len(idx) * len(z), "max", err.max(), "at z", z[w[0]], "l", ls[w[1]], "r Mpc/h", sep[idx][w[2]] * T.H, "p99.9", np.quantile(err, 0.999), "median", np.median(err))
