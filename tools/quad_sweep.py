#!/usr/bin/env python
"""Error and cost of candidate fixed-node rules on a dense subsample of the cfg5 grid (GPU).

Truth = the kernels with every node count doubled and every panel narrowed (tests/test_gpu_parity.py::FINE_RULES), which the
converged-oracle fixture pins independently.  For each candidate: max / 99.9th percentile of |candidate - fine| relative to the
l = 0..4 scale of the point, per piece, and the integrand evaluations + device time of a full cfg5 step."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import coffe_b200  # noqa: E402
from coffe_b200.default_tables import load_default_tableset  # noqa: E402

H = 0.67
import os
ALL = os.environ["COFFE_SWEEP_CONTRIB"].split(",") if os.environ.get("COFFE_SWEEP_CONTRIB") else list(coffe_b200.capi.CONTRIBUTIONS)
FINE = dict(n_mu_panel=12, n_mu_panel_di=10, mu_panel_width=16.0, n_outer=96, n_panel_si=12, n_panel_di=12, panel_width=8.0,
            panel_uniform_extent=640.0, panel_width_di=16.0, panel_uniform_extent_di=640.0, mu_panel_width_di=16.0,
            tail_ratio=1.3, tail_ratio_di=1.3, geometric_shift_di=1, geometric_growth=1.0, geometric_growth_di=1.0, n_renorm=64, tail_map=1, tail_map_di=1, mu_geometric_growth=1.0, mu_geometric_growth_di=1.0, mu_switch_break_di=1, mu_tail_ratio=1.3, n_mu_ni=64)
z, sep, ls = np.linspace(0.5, 2.0, 32), np.linspace(10.0, 300.0, 1000) / H, [0, 2, 4]
idx = np.arange(0, 1000, 8)
zi = np.arange(0, 32, 2)                      # 16 z x 125 r = 2000 points
eng = coffe_b200.Engine(0)
eng.upload_tables(load_default_tableset())
eng.set_options(ALL)
eng.set_quadrature(**FINE)
fine = np.concatenate([eng.multipoles_pieces(z[zi][i:i + 4], sep[idx], ls) for i in range(0, len(zi), 4)], axis=1)     # [3, nz, nl, ns]
scale = np.max(np.abs(fine.sum(0)), axis=1, keepdims=True)

cands = json.loads(sys.argv[1]) if len(sys.argv) > 1 else {"default": {}}
for name, kw in cands.items():
    eng.set_quadrature(**kw)
    got = eng.multipoles_pieces(z[zi], sep[idx], ls)              # pieces: always at piece accuracy unless the target is pinned
    err = np.abs(got - fine) / scale[None]
    tot = np.abs(eng.multipoles(z[zi], sep[idx], ls) - fine.sum(0)) / scale          # the total under the rules in force
    plan = eng.plan_multipoles(z, sep, ls)
    plan.execute(); ms = plan.execute()
    si_ms, di_ms = eng.last_kernel_times()
    _, _, n_si, n_di = plan.count()
    plan.close()
    print(f"{name:28s} total max {tot.max():.2e} p99.9 {np.quantile(tot, 0.999):.2e} | pieces (piece rules) SI max {err[1].max():.2e} DI max {err[2].max():.2e} "
          f"| step {ms:7.2f} ms (SI {si_ms:6.2f} DI {di_ms:7.2f}) evals SI {n_si:.3e} DI {n_di:.3e}  {kw}", flush=True)
