"""Generate the converged-oracle fixtures the dense parity tests and bench.py compare the kernels with:

  tests/golden/converged_cfg5.npz   128 stratified (z, r) points of the benchmarked grid (32 z_mean x 1000 separations, all effects):
                                    multipoles l = 0, 2, 4 of the three pieces
  tests/golden/converged_cfg2.npz   32 (mu, r) points of cfg2's 50 x 100 grid (full-sky 2PCF, all effects, z_mean = 1)
  tests/golden/converged_cfg3.npz   redshift-averaged multipoles of cfg3 (den + rsd + len): 2 bins x 3 separations

Every number is an integral of the REFERENCE's own integrand functions (oracle/_ref/libcoffe_ref.so = /root/reference/src compiled in
place); nothing of the product is involved.  Integrators, per piece:
  local x local            reference integrand under a 64-node Gauss-Legendre rule in mu (multipoles) / evaluated directly (2PCF)
  local x integrated       multipoles: the vendored Cuhre called directly on the reference integrand in (mu, x), epsrel 2e-7 -- adaptive and
                           panel-free, i.e. independent of the product's panel scheme; cross-checked against oracle/converged.py
                           2PCF: the reference's own QAG path (coffe_integrate, src/signal.c:584-606) at integration_1d_prec = 1e-9
  integrated x integrated  2PCF: Cuhre directly in (x1, x2), epsrel 1e-7 (scheme-independent)
                           multipoles: oracle/converged.py at two rule sizes (the 3-D integral is beyond Cuhre's reach);
                           the difference of the two sizes is stored as the oracle's own convergence error
Run where /root/reference exists and the oracle is built:   python tests/golden/make_converged_fixtures.py [cfg5] [cfg2] [cfg3]
(about 1.5 h on 8 cores for all three; results are cached per point under /tmp/coffe_fixture_cache so the script can be resumed).
"""
import os
import pickle
import sys
from multiprocessing import Pool
from pathlib import Path

os.environ["OMP_NUM_THREADS"] = "1"          # one process per core, each single-threaded (Cuhre is serial anyway)
import numpy as np  # noqa: E402

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
H = 0.67
ALL = ("den", "rsd", "len", "d1", "d2", "g1", "g2", "g3", "g4", "g5")
LS = [0, 2, 4]
CACHE = Path(os.environ.get("COFFE_FIXTURE_CACHE", "/tmp/coffe_fixture_cache"))
RULE_TRUTH = dict(n_outer=96, n_panel=10, mu_kw=dict(n=8, W=16.0))
RULE_CHECK = dict(n_outer=64, n_panel=8, mu_kw=dict(n=6, W=24.0))

_ref = {}


def reference(contributions=ALL):
    """same set-up as tests/golden/make_default_tables.py: the tables bench.py and the cfg5 tests upload"""
    key = tuple(contributions)
    if key not in _ref:
        from oracle import coffe_oracle as O
        r = O.RefCoffe()
        r.set_power_spectrum(O.load_pk())
        r.set_z_mean([2.0]); r.set_deltaz([0.3]); r.set_l(LS)
        r.set_contributions(list(contributions))
        r.lib.ref_integrals(r.h_)
        _ref[key] = r
    return _ref[key]


def cached(name, fn):
    CACHE.mkdir(parents=True, exist_ok=True)
    f = CACHE / (name + ".pkl")
    if f.exists():
        return pickle.loads(f.read_bytes())
    v = fn()
    f.write_bytes(pickle.dumps(v))
    return v


def job_multipoles(args):
    """one (z, sep): dict of the pieces under each integrator"""
    tag, contributions, z, sep = args[:4]
    with_check = args[4] if len(args) > 4 else True
    from oracle.converged import ConvergedOracle

    def run():
        ref = reference(contributions)
        truth = ConvergedOracle(ref, **RULE_TRUTH)
        check = ConvergedOracle(ref, **RULE_CHECK)
        out = {"z": z, "sep": sep}
        out["ni"] = truth.multipoles_pieces(z, sep, LS, want=(True, False, False))[0]
        out["si_cuhre"], out["si_cuhre_err"], out["si_cuhre_fail"] = truth.si_multipoles_cuhre(z, sep, LS, epsrel=2e-7, maxeval=6_000_000)
        out["si_rule"] = truth.multipoles_pieces(z, sep, LS, want=(False, True, False))[1]
        out["di"] = truth.multipoles_pieces(z, sep, LS, want=(False, False, True))[2]
        # the second rule size of the 3-D piece (the oracle's own convergence) on a quarter of the points: it is a third of the cost
        out["di_check"] = check.multipoles_pieces(z, sep, LS, want=(False, False, True))[2] if with_check else np.full(len(LS), np.nan)
        return out
    return cached(f"{tag}_z{z:.10f}_s{sep:.10f}", run)


def job_corrfunc(args):
    tag, z, sep, mu = args
    from oracle.converged import ConvergedOracle

    def run():
        ref = reference(ALL)
        truth = ConvergedOracle(ref, **RULE_TRUTH)
        out = {"z": z, "sep": sep, "mu": mu}
        out["ni"] = truth.ni(z, sep, mu)
        ref.set_integration(prec=1e-9)
        out["si_qag"] = ref.integrate(z, sep, mu, 0, 1, 1)              # SINGLE_INTEGRATED, CORRFUNC: the reference's own 1-D QAG
        ref.set_integration(prec=1e-5)
        out["si_rule"] = truth.si(z, sep, mu)
        out["di_cuhre"], out["di_cuhre_err"], out["di_cuhre_fail"] = truth.di_corrfunc_cuhre(z, sep, mu, epsrel=1e-7)
        out["di_rule"] = truth.di(z, sep, mu)
        return out
    return cached(f"{tag}_z{z:.6f}_s{sep:.10f}_m{mu:.10f}", run)


def cfg5_points():
    z, sep = np.linspace(0.5, 2.0, 32), np.linspace(10.0, 300.0, 1000) / H
    iz = np.arange(128) % 32
    ir = np.round(((np.arange(128) * 0.6180339887498949) % 1.0) * 999).astype(int)
    # corners and the neighbourhood of the rule switches of the product (they are the places a fixed-node rule would fail first)
    forced = {0: (0, 0), 1: (31, 0), 2: (0, 999), 3: (31, 999), 4: (0, int(np.searchsorted(sep, 240.0))), 5: (31, int(np.searchsorted(sep, 240.0)) - 1),
              6: (16, int(np.searchsorted(sep, 256.0))), 7: (8, int(np.searchsorted(sep, 64.0))), 8: (24, int(np.searchsorted(sep, 32.0)) - 1)}
    for k, (a, b) in forced.items():
        iz[k], ir[k] = a, b
    return z, sep, iz, ir


def make_cfg5(pool):
    z, sep, iz, ir = cfg5_points()
    res = pool.map(job_multipoles, [("cfg5", ALL, float(z[a]), float(sep[b]), k % 4 == 0) for k, (a, b) in enumerate(zip(iz, ir))], chunksize=1)
    pieces = np.array([[r["ni"], r["si_cuhre"], r["di"]] for r in res])             # [k, piece, l]
    scale = np.max(np.abs(pieces.sum(1)), axis=1, keepdims=True)
    si_scale = np.max(np.abs(pieces[:, 1]), axis=1, keepdims=True)
    di_scale = np.max(np.abs(pieces[:, 2]), axis=1, keepdims=True)
    si_rule = np.array([r["si_rule"] for r in res])
    di_check = np.array([r["di_check"] for r in res])
    checked = np.isfinite(di_check[:, 0])
    di_dev = np.where(checked[:, None], np.abs(pieces[:, 2] - np.nan_to_num(di_check)), 0.0)
    self_err = np.maximum(np.abs(pieces[:, 1] - si_rule), di_dev) / scale
    self_err_pieces = np.stack([np.abs(pieces[:, 1] - si_rule) / si_scale, di_dev / di_scale], axis=1)
    cuhre_err = np.array([r["si_cuhre_err"] for r in res])
    print(f"cfg5: {len(res)} points ({int(checked.sum())} with the second rule size of the 3-D piece); oracle self error (of the point scale) max {self_err.max():.2e}; per piece (SI Cuhre vs rule, DI two rule sizes) "
          f"{self_err_pieces[:, 0].max():.2e} {self_err_pieces[:, 1].max():.2e}; Cuhre's own estimate max {cuhre_err.max():.2e}, "
          f"budget exhausted on {int(sum(np.any(r['si_cuhre_fail']) for r in res))} points")
    np.savez_compressed(ROOT / "tests" / "golden" / "converged_cfg5.npz", iz=iz, ir=ir, z=z[iz], sep=sep[ir], l=np.array(LS), pieces=pieces,
                        si_rule=si_rule, di_check=di_check, oracle_self_error=self_err, oracle_self_error_pieces=self_err_pieces,
                        si_cuhre_rel_error_estimate=cuhre_err)


def make_cfg2(pool):
    sep, mu = np.linspace(10.0, 300.0, 100) / H, np.linspace(-0.98, 0.98, 50)
    k = np.arange(32)
    imu = np.round(((k * 0.6180339887498949) % 1.0) * 49).astype(int)
    ir = np.round(((k * 0.7548776662466927) % 1.0) * 99).astype(int)
    imu[:4], ir[:4] = [0, 49, 0, 49], [0, 0, 99, 99]
    res = pool.map(job_corrfunc, [("cfg2", 1.0, float(sep[b]), float(mu[a])) for a, b in zip(imu, ir)], chunksize=1)
    pieces = np.array([[r["ni"], r["si_qag"], r["di_cuhre"]] for r in res])
    rule = np.array([[r["ni"], r["si_rule"], r["di_rule"]] for r in res])
    tot = np.abs(pieces.sum(1))
    self_err = np.abs(pieces - rule).max(1) / tot
    print(f"cfg2: {len(res)} points; adaptive integrators (QAG 1e-9, Cuhre 1e-7) vs the fixed rule of oracle/converged.py: max {self_err.max():.2e} of the 2PCF; "
          f"Cuhre estimate max {max(r['di_cuhre_err'] for r in res):.2e}, budget exhausted on {sum(r['di_cuhre_fail'] != 0 for r in res)} points")
    np.savez_compressed(ROOT / "tests" / "golden" / "converged_cfg2.npz", imu=imu, ir=ir, mu=mu[imu], sep=sep[ir], z=np.array(1.0), pieces=pieces,
                        rule_pieces=rule, oracle_self_error=self_err)


CFG3 = ("den", "rsd", "len")
CFG3_BINS = ([0.5, 1.8125], [0.6875, 2.0])
CFG3_SEP = np.array([10.0, 60.0, 160.0]) / H
CFG3_NZ = 16


def make_cfg3(pool):
    from oracle.converged import ConvergedOracle
    ref = reference(CFG3)
    co = ConvergedOracle(ref, **RULE_TRUTH)
    jobs, index = [], []
    for ib in range(2):
        for ir, s in enumerate(CFG3_SEP):
            zs, ws = co.average_nodes(CFG3_BINS[0][ib], CFG3_BINS[1][ib], s, CFG3_NZ)
            for zk, wk in zip(zs, ws):
                jobs.append(("cfg3", CFG3, float(zk), float(s))); index.append((ib, ir, wk))
    res = pool.map(job_multipoles, jobs, chunksize=1)
    pieces = np.zeros((2, 3, 3, 3))          # [bin, sep, piece, l]
    check = np.zeros((2, 3, 3, 3))
    for (ib, ir, wk), r in zip(index, res):
        pieces[ib, ir] += wk * np.array([r["ni"], r["si_cuhre"], r["di"]])
        check[ib, ir] += wk * np.array([r["ni"], r["si_rule"], r["di_check"]])
    # convergence of the z rule itself, on the cheap local piece: 16 nodes against 32
    zerr = 0.0
    for ib in range(2):
        for ir, s in enumerate(CFG3_SEP):
            zs, ws = co.average_nodes(CFG3_BINS[0][ib], CFG3_BINS[1][ib], s, 32)
            ni32 = sum(wk * co.multipoles_pieces(float(zk), s, LS, want=(True, False, False))[0] for zk, wk in zip(zs, ws))
            zerr = max(zerr, np.max(np.abs(ni32 - pieces[ib, ir, 0]) / np.max(np.abs(ni32))))
    scale = np.max(np.abs(pieces.sum(2)), axis=2, keepdims=True)
    self_err = np.abs(pieces.sum(2) - check.sum(2)) / scale
    print(f"cfg3: 2 bins x 3 separations x {CFG3_NZ} z nodes; z rule 16 vs 32 nodes on the local piece {zerr:.2e}; oracle self error max {self_err.max():.2e}")
    np.savez_compressed(ROOT / "tests" / "golden" / "converged_cfg3.npz", z_min=np.array(CFG3_BINS[0]), z_max=np.array(CFG3_BINS[1]), sep=CFG3_SEP,
                        l=np.array(LS), pieces=pieces, oracle_self_error=self_err, z_rule_error=np.array(zerr), contributions=np.array(CFG3))


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if a.startswith("cfg")] or ["cfg2", "cfg3", "cfg5"]
    with Pool(int(os.environ.get("COFFE_FIXTURE_PROCS", os.cpu_count() or 1))) as pool:
        for w in which:
            {"cfg5": make_cfg5, "cfg2": make_cfg2, "cfg3": make_cfg3}[w](pool)
