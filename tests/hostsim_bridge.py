"""TEST INFRASTRUCTURE: ctypes bridge to tests/hostsim (the hot path's headers compiled for the CPU)."""
import ctypes as C
from pathlib import Path

import numpy as np

from coffe_b200 import capi
from oracle_tables import tableset_from_oracle

LIB = Path(__file__).resolve().parent / "hostsim" / "_build" / "libhostsim.so"


def available():
    return LIB.exists()


class HostSim:
    def __init__(self, ref, contributions, tableset=None, compact=True, **flags):
        self.lib = C.CDLL(str(LIB))
        self.lib.hs_create.restype = C.c_void_p
        for f in ("hs_ni", "hs_si", "hs_di"):
            getattr(self.lib, f).restype = C.c_double
        self.ts = tableset if tableset is not None else tableset_from_oracle(ref)
        o = capi.make_options(contributions, **flags)
        self.h = C.c_void_p(self.lib.hs_create(C.byref(self.ts.struct), C.byref(o)))
        assert self.h.value, "hostsim: table assembly failed"
        self.lib.hs_use_compact.restype = C.c_int
        self.compact_info = None
        if compact:
            self.use_compact()

    def use_compact(self, knot_budget=-1, tol=0.0):
        """select the shared-memory (resampled) I^n_l tables, library defaults unless given; 0 = exact tables.
        Returns (knots, achieved max relative error, blob bytes, subdivisions m per region)"""
        info = np.zeros(130)
        n = self.lib.hs_use_compact(self.h, C.c_int(knot_budget), C.c_double(tol), capi.ptr(info), C.c_int(len(info)))
        assert n >= 0, "hostsim: compact table construction failed"
        self.compact_info = (n, info[0], int(info[1]), info[3:].astype(int))
        self.compact_err_i40 = info[2]
        return self.compact_info

    def tables_at(self, r):
        """the eight I^n_l and I^4_0 at r through the selected tables"""
        out = np.zeros(9)
        self.lib.hs_tables_at(self.h, C.c_double(r), capi.ptr(out))
        return out

    def __del__(self):
        try:
            self.lib.hs_destroy(self.h)
        except Exception:
            pass

    def set_quadrature(self, **kw):
        q = capi.make_quadrature(**kw)
        self.lib.hs_set_quadrature(self.h, C.byref(q))

    def set_di_mode(self, legacy=False, lanes=1):
        """legacy: di_eval on generic ridge panels (cross-check); otherwise di_fast.cuh emulating `lanes` lanes"""
        self.lib.hs_set_di_mode(self.h, C.c_int(int(legacy)), C.c_int(lanes))

    def ni(self, z, mu, sep):
        return self.lib.hs_ni(self.h, C.c_double(z), C.c_double(mu), C.c_double(sep))

    def si(self, z, mu, sep, x, exact=True):
        return self.lib.hs_si(self.h, C.c_double(z), C.c_double(mu), C.c_double(sep), C.c_double(x), C.c_int(int(exact)))

    def di(self, z, mu, sep, x1, x2, exact=True):
        return self.lib.hs_di(self.h, C.c_double(z), C.c_double(mu), C.c_double(sep), C.c_double(x1), C.c_double(x2), C.c_int(int(exact)))

    def item(self, z, sep, mu):
        out = np.zeros(3)
        self.lib.hs_item(self.h, C.c_double(z), C.c_double(sep), C.c_double(mu), capi.ptr(out))
        return out

    def multipoles(self, z, sep, ls):
        ls = capi.iarr(ls)
        out = np.zeros(3 * len(ls))
        self.lib.hs_multipoles(self.h, C.c_double(z), C.c_double(sep), capi.iptr(ls), C.c_int(len(ls)), capi.ptr(out))
        return out.reshape(3, len(ls))

    def n_mu(self, sep):
        """nodes of the mu rule of the local x integrated piece at this separation"""
        self.lib.hs_n_mu.restype = C.c_int
        return self.lib.hs_n_mu(self.h, C.c_double(sep))

    def mu_rule(self, sep, piece, half_only=False):
        """(mu, w) of the mu rule of piece 1 (local x integrated) or 2 (integrated x integrated) at this separation"""
        mu, w = np.zeros(4096), np.zeros(4096)
        self.lib.hs_mu_rule.restype = C.c_int
        n = self.lib.hs_mu_rule(self.h, C.c_double(sep), C.c_int(piece), C.c_int(int(half_only)), C.c_int(4096), capi.ptr(mu), capi.ptr(w))
        return mu[:n].copy(), w[:n].copy()

    def eval_counts(self, z, sep):
        """(SI, DI) integrand evaluations of the fixed-node rules at one (z, sep), summed over the mu items"""
        out = np.zeros(2)
        self.lib.hs_eval_counts(self.h, C.c_double(z), C.c_double(sep), capi.ptr(out))
        return out

    def flatsky(self, z, sep, l, n_x=96):
        """(density-lensing, lensing-lensing) flat-sky closed-form multipoles"""
        out = np.zeros(2)
        self.lib.hs_flatsky(self.h, C.c_double(z), C.c_double(sep), C.c_int(l), C.c_int(n_x), capi.ptr(out))
        return out
