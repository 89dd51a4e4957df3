"""Thin Python handle on libcoffe_b200.so (the C ABI in include/coffe_b200.h).

`Engine` owns one `coffe_b200_ctx` (one GPU).  It does no arithmetic of its own: every number it returns
was computed by the sm_100a kernels behind the C ABI, and it raises if the library or a GPU is missing.
"""
import ctypes as C

import numpy as np

from . import capi


class CoffeB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libcoffe_b200 error {capi.ERRORS.get(code, code)}: {message}")
        self.code = code


class Engine:
    def __init__(self, device=0):
        self.lib = capi.load()
        h = C.c_void_p()
        rc = self.lib.coffe_b200_create(C.c_int(device), C.byref(h))
        if rc:
            raise CoffeB200Error(rc, self.lib.coffe_b200_last_error(None).decode())
        self._h = h
        self._tables = None
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self.lib.coffe_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise CoffeB200Error(rc, self.lib.coffe_b200_last_error(self._h).decode())

    # ---- state --------------------------------------------------------------------------------------
    def upload_tables(self, tableset):
        self._tables = tableset  # keep the host buffers alive during the call
        self._check(self.lib.coffe_b200_upload_tables(self._h, C.byref(tableset.struct)))

    def set_options(self, contributions, **flags):
        o = capi.make_options(contributions, **flags)
        self._check(self.lib.coffe_b200_set_options(self._h, C.byref(o)))

    def set_quadrature(self, **kw):
        q = capi.make_quadrature(**kw)
        self._check(self.lib.coffe_b200_set_quadrature(self._h, C.byref(q)))

    def get_quadrature(self):
        q = capi.Quadrature()
        self._check(self.lib.coffe_b200_get_quadrature(self._h, C.byref(q)))
        return {name: getattr(q, name) for name, _ in q._fields_}

    def set_kernel_variant(self, variant):
        """1 = shared-memory (resampled) I^n_l tables, the default; 0 = exact tables through L1/L2"""
        self._check(self.lib.coffe_b200_set_kernel_variant(self._h, C.c_int(variant)))

    def table_info(self):
        """(variant the last call ran, knots / bytes / achieved max relative error of the shared-memory tables)"""
        v, k, b, err = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        self._check(self.lib.coffe_b200_table_info(self._h, C.byref(v), C.byref(k), C.byref(b), C.byref(err)))
        return v.value, k.value, b.value, err.value

    def last_timing(self):
        ms, n = C.c_double(), C.c_int()
        self._check(self.lib.coffe_b200_last_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def last_kernel_times(self):
        """(local x integrated, integrated x integrated) kernel times of the last multipole computation [ms]"""
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.coffe_b200_last_kernel_times(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- signal -------------------------------------------------------------------------------------
    def multipoles(self, z_mean, sep, l):
        z, s, ls = capi.darr(z_mean), capi.darr(sep), capi.iarr(l)
        out = np.empty((len(z), len(ls), len(s)))
        self._check(self.lib.coffe_b200_multipoles(self._h, capi.ptr(z), C.c_int(len(z)), capi.ptr(s), C.c_int(len(s)),
                                                   capi.iptr(ls), C.c_int(len(ls)), capi.ptr(out)))
        return out

    def multipoles_pieces(self, z_mean, sep, l):
        """[3 pieces (NI, SI, DI), n_z, n_l, n_sep]"""
        z, s, ls = capi.darr(z_mean), capi.darr(sep), capi.iarr(l)
        out = np.empty((3, len(z), len(ls), len(s)))
        self._check(self.lib.coffe_b200_multipoles_pieces(self._h, capi.ptr(z), C.c_int(len(z)), capi.ptr(s), C.c_int(len(s)),
                                                          capi.iptr(ls), C.c_int(len(ls)), capi.ptr(out)))
        return out

    def multipoles_range(self, z_mean, sep, l, begin, end):
        """[n_l, end-begin] for the flattened (iz, ir) items [begin, end)"""
        z, s, ls = capi.darr(z_mean), capi.darr(sep), capi.iarr(l)
        out = np.empty((len(ls), max(end - begin, 0)))
        self._check(self.lib.coffe_b200_multipoles_range(self._h, capi.ptr(z), C.c_int(len(z)), capi.ptr(s), C.c_int(len(s)),
                                                         capi.iptr(ls), C.c_int(len(ls)), C.c_long(begin), C.c_long(end), capi.ptr(out)))
        return out

    def corrfunc(self, z_mean, mu, sep):
        z, m, s = capi.darr(z_mean), capi.darr(mu), capi.darr(sep)
        out = np.empty((len(z), len(m), len(s)))
        self._check(self.lib.coffe_b200_corrfunc(self._h, capi.ptr(z), C.c_int(len(z)), capi.ptr(m), C.c_int(len(m)),
                                                 capi.ptr(s), C.c_int(len(s)), capi.ptr(out)))
        return out

    def corrfunc_pieces(self, z_mean, mu, sep, begin=0, end=None):
        """[3 pieces, end - begin] of the flattened output entries (iz*n_mu + imu)*n_sep + ir in [begin, end)"""
        z, m, s = capi.darr(z_mean), capi.darr(mu), capi.darr(sep)
        if end is None:
            end = len(z) * len(m) * len(s)
        out = np.empty((3, max(end - begin, 0)))
        self._check(self.lib.coffe_b200_corrfunc_range_pieces(self._h, capi.ptr(z), C.c_int(len(z)), capi.ptr(m), C.c_int(len(m)),
                                                              capi.ptr(s), C.c_int(len(s)), C.c_long(begin), C.c_long(end), capi.ptr(out)))
        return out

    def average_multipoles_pieces(self, z_min, z_max, sep, l, begin=0, end=None):
        """[3 pieces, n_l, end - begin] of the flattened (ibin*n_sep + ir) groups in [begin, end)"""
        a, b, s, ls = capi.darr(z_min), capi.darr(z_max), capi.darr(sep), capi.iarr(l)
        if end is None:
            end = len(a) * len(s)
        out = np.empty((3, len(ls), max(end - begin, 0)))
        self._check(self.lib.coffe_b200_average_multipoles_range_pieces(self._h, capi.ptr(a), capi.ptr(b), C.c_int(len(a)), capi.ptr(s), C.c_int(len(s)),
                                                                        capi.iptr(ls), C.c_int(len(ls)), C.c_long(begin), C.c_long(end), capi.ptr(out)))
        return out

    def multipoles_cost(self, z_mean, sep, l):
        """[n_z, n_sep] estimated cost (integrand evaluations) of every item, counted on the device"""
        z, s, ls = capi.darr(z_mean), capi.darr(sep), capi.iarr(l)
        out = np.empty((len(z), len(s)))
        self._check(self.lib.coffe_b200_multipoles_cost(self._h, capi.ptr(z), C.c_int(len(z)), capi.ptr(s), C.c_int(len(s)),
                                                        capi.iptr(ls), C.c_int(len(ls)), capi.ptr(out)))
        return out

    # ---- one process per GPU: NCCL inside the library ------------------------------------------------------
    def comm_init(self, rank, world, unique_id):
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        self._check(self.lib.coffe_b200_comm_init(self._h, C.c_int(rank), C.c_int(world), buf))
        self.rank, self.world = rank, world

    def comm_destroy(self):
        self.lib.coffe_b200_comm_destroy(self._h)

    def multipoles_sharded(self, z_mean, sep, l, bounds=None, want_result=True):
        """every rank computes its block, the blocks are gathered device to device (NCCL) and copied to the host of the ranks that
        ask for it.  Returns ([n_z, n_l, n_sep] or None, kernel ms, gather ms)"""
        z, s, ls = capi.darr(z_mean), capi.darr(sep), capi.iarr(l)
        out = np.empty((len(z), len(ls), len(s))) if want_result else None
        bd = (C.c_long * len(bounds))(*bounds) if bounds is not None else None
        kms, gms = C.c_double(), C.c_double()
        self._check(self.lib.coffe_b200_multipoles_sharded(self._h, capi.ptr(z), C.c_int(len(z)), capi.ptr(s), C.c_int(len(s)), capi.iptr(ls), C.c_int(len(ls)),
                                                           bd, capi.ptr(out) if want_result else None, C.byref(kms), C.byref(gms)))
        return out, kms.value, gms.value

    def average_multipoles(self, z_min, z_max, sep, l):
        a, b, s, ls = capi.darr(z_min), capi.darr(z_max), capi.darr(sep), capi.iarr(l)
        assert len(a) == len(b)
        out = np.empty((len(a), len(ls), len(s)))
        self._check(self.lib.coffe_b200_average_multipoles(self._h, capi.ptr(a), capi.ptr(b), C.c_int(len(a)), capi.ptr(s), C.c_int(len(s)),
                                                           capi.iptr(ls), C.c_int(len(ls)), capi.ptr(out)))
        return out

    def _covariance_params(self, pk, sep, l, fsky, density1, density2, pixelsize, z_mean=None, deltaz=None, z_min=None, z_max=None,
                           pops=(1, 1, 1, 1), cosmic=True, mixed=True, poisson=True, window=False, k_min=None, k_max=None,
                           method="standard", bins=8000):
        average = z_mean is None
        k, p = capi.darr(pk[0]), capi.darr(pk[1])
        s, ls = capi.darr(sep), capi.iarr(l)
        n_bins = len(z_min) if average else len(z_mean)
        arrs = {name: capi.darr(v) for name, v in (("fsky", fsky), ("density1", density1), ("density2", density2), ("pixelsize", pixelsize))}
        for name, v in arrs.items():
            assert len(v) == n_bins, name
        cp = capi.CovarianceParams()
        cp.pk_k, cp.pk_p, cp.n_pk = capi.ptr(k), capi.ptr(p), len(k)
        cp.k_min = float(k[0] if k_min is None else k_min)
        cp.k_max = float(k[-1] if k_max is None else k_max)
        cp.n_bins = n_bins
        keep = []
        for name, v in (("z_mean", z_mean), ("deltaz", deltaz), ("z_min", z_min), ("z_max", z_max)):
            if v is not None:
                a = capi.darr(v)
                keep.append(a)
                setattr(cp, name, capi.ptr(a))
        for name, a in arrs.items():
            setattr(cp, name, capi.ptr(a))
        cp.sep, cp.n_sep, cp.l, cp.n_l = capi.ptr(s), len(s), capi.iptr(ls), len(ls)
        cp.pop = (C.c_int * 4)(*pops)
        cp.cosmic, cp.mixed, cp.poisson, cp.window, cp.average = int(cosmic), int(mixed), int(poisson), int(window), int(average)
        cp.integration_method, cp.integration_bins = {"standard": 1, "fftlog": 2}[method], int(bins)
        self._cov_keep = (k, p, s, ls, keep, arrs)
        return cp, (n_bins, len(ls), len(s), len(ls), len(s))

    def covariance(self, *args, rows=None, **kw):
        """[n_bins, n_l, n_sep, n_l, n_sep]; pk = (k, P) knots of the power-spectrum spline.  rows = (begin, end): only the rows
        il1*n_sep + ir1 in that range of every bin are computed (the rest of the array stays NaN): the unit of multi-GPU sharding"""
        cp, shape = self._covariance_params(*args, **kw)
        if rows is None:
            out = np.empty(shape)
            self._check(self.lib.coffe_b200_covariance(self._h, C.byref(cp), capi.ptr(out)))
        else:
            out = np.full(shape, np.nan)
            self._check(self.lib.coffe_b200_covariance_rows(self._h, C.byref(cp), C.c_int(rows[0]), C.c_int(rows[1]), capi.ptr(out)))
        return out

    # ---- I^n_l tables ---------------------------------------------------------------------------------
    def integrals(self, pk, k_min, k_max, tables, bessel_bins=10000, divergent=False, chi_max=0.0, interp_method=5):
        """what coffe_integrals_init builds (src/integrals.c:580-1423), on the device.
        pk = (k, P) knots; tables = list of (n, l) or (n, l, half).  Returns a dict: r [n_r], tables [n_tables, n_r] and,
        with divergent, r40, i40, renorm0 (on the knots r40 with r40[0] -> 0), renorm_chi [201], renorm_z [201, 201]
        with renorm_z[k, j] <-> (chi_j, chi_k); else those are None."""
        k, p = capi.darr(pk[0]), capi.darr(pk[1])
        tb = [tuple(t) + (0,) * (3 - len(t)) for t in tables]
        n_, l_, h_ = capi.iarr([t[0] for t in tb]), capi.iarr([t[1] for t in tb]), capi.iarr([t[2] for t in tb])
        ip_ = capi.IntegralsParams()
        ip_.pk_k, ip_.pk_p, ip_.n_pk, ip_.interp_method = capi.ptr(k), capi.ptr(p), len(k), int(interp_method)
        ip_.k_min, ip_.k_max, ip_.bessel_bins = float(k_min), float(k_max), int(bessel_bins)
        ip_.n, ip_.l, ip_.half, ip_.n_tables = capi.iptr(n_), capi.iptr(l_), capi.iptr(h_), len(tb)
        ip_.divergent, ip_.chi_max = int(bool(divergent)), float(chi_max)
        n_r, n_r40, n_ren = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.coffe_b200_integrals_sizes(C.byref(ip_), C.byref(n_r), C.byref(n_r40), C.byref(n_ren)))
        r, out = np.zeros(n_r.value), np.zeros((max(len(tb), 1), n_r.value))
        div = bool(divergent)
        r40, i40, ren0 = np.zeros(n_r40.value), np.zeros(n_r40.value), np.zeros(n_r40.value)
        chi, rz = np.zeros(n_ren.value), np.zeros((n_ren.value, n_ren.value))
        null = C.POINTER(C.c_double)()
        self._check(self.lib.coffe_b200_integrals(self._h, C.byref(ip_), capi.ptr(r), capi.ptr(out),
                                                  capi.ptr(r40) if div else null, capi.ptr(i40) if div else null, capi.ptr(ren0) if div else null,
                                                  capi.ptr(chi) if div else null, capi.ptr(rz) if div else null))
        return {"r": r, "tables": out[:len(tb)], "r40": r40 if div else None, "i40": i40 if div else None, "renorm0": ren0 if div else None,
                "renorm_chi": chi if div else None, "renorm_z": rz if div else None}

    def integrals2d(self, k, pk2d, k_min, k_max, tables, bessel_bins=10000, interp_method=5):
        """I^n_l(z_i, r) of every table from the rows pk2d[i] = P(k, z_i): what coffe_integrals_init adds per table with CLASS and the
        midpoint approximation (src/integrals.c:663-767).  Returns (r [n_r], values [n_tables, n_r, n_z])."""
        k, pk2d = capi.darr(k), capi.darr(pk2d)
        assert pk2d.ndim == 2 and pk2d.shape[1] == len(k)
        tb = [tuple(t) + (0,) * (3 - len(t)) for t in tables]
        n_, l_, h_ = capi.iarr([t[0] for t in tb]), capi.iarr([t[1] for t in tb]), capi.iarr([t[2] for t in tb])
        ip_ = capi.IntegralsParams()
        ip_.pk_k, ip_.pk_p, ip_.n_pk, ip_.interp_method = capi.ptr(k), capi.ptr(pk2d), len(k), int(interp_method)
        ip_.k_min, ip_.k_max, ip_.bessel_bins = float(k_min), float(k_max), int(bessel_bins)
        ip_.n, ip_.l, ip_.half, ip_.n_tables = capi.iptr(n_), capi.iptr(l_), capi.iptr(h_), len(tb)
        n_r, n_r40, n_ren = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.coffe_b200_integrals_sizes(C.byref(ip_), C.byref(n_r), C.byref(n_r40), C.byref(n_ren)))
        out = np.zeros((len(tb), n_r.value, pk2d.shape[0]))
        self._check(self.lib.coffe_b200_integrals2d(self._h, C.byref(ip_), capi.ptr(pk2d), C.c_int(pk2d.shape[0]), capi.ptr(out)))
        r = self.integrals((k, pk2d[0]), k_min, k_max, tables[:1], bessel_bins=bessel_bins, interp_method=interp_method)["r"]
        return r, out

    # ---- prepared computations (inputs resident in HBM) -----------------------------------------------
    def plan_multipoles(self, z_mean, sep, l, begin=0, end=None):
        z, s, ls = capi.darr(z_mean), capi.darr(sep), capi.iarr(l)
        if end is None:
            end = len(z) * len(s)
        p = C.c_void_p()
        self._check(self.lib.coffe_b200_plan_multipoles(self._h, capi.ptr(z), C.c_int(len(z)), capi.ptr(s), C.c_int(len(s)),
                                                        capi.iptr(ls), C.c_int(len(ls)), C.c_long(begin), C.c_long(end), C.byref(p)))
        return Plan(self, p, len(ls), end - begin)

    def fp64_peak_tflops(self):
        v = C.c_double()
        self._check(self.lib.coffe_b200_fp64_peak(self._h, C.byref(v)))
        return v.value


class Plan:
    def __init__(self, engine, handle, n_l, n_items):
        self.engine, self._p, self.n_l, self.n = engine, handle, n_l, n_items

    def execute(self):
        """launch the kernels on the resident inputs; returns their device time in ms"""
        ms = C.c_double()
        self.engine._check(self.engine.lib.coffe_b200_plan_execute(self._p, C.byref(ms)))
        return ms.value

    def fetch(self):
        out = np.empty((self.n_l, self.n))
        self.engine._check(self.engine.lib.coffe_b200_plan_fetch(self._p, capi.ptr(out)))
        return out

    def execute_gather(self, bounds):
        """kernels of this rank's block + the NCCL gather of all blocks, both on the compute stream: (kernel ms, gather ms)"""
        bd = (C.c_long * len(bounds))(*bounds)
        self._n_total = int(bounds[-1])
        kms, gms = C.c_double(), C.c_double()
        self.engine._check(self.engine.lib.coffe_b200_plan_execute_gather(self._p, bd, C.c_long(self._n_total), C.byref(kms), C.byref(gms)))
        return kms.value, gms.value

    def fetch_gathered(self):
        out = np.empty((self.n_l, self._n_total))
        self.engine._check(self.engine.lib.coffe_b200_plan_fetch_gathered(self._p, capi.ptr(out)))
        return out

    def count(self):
        """(work items, NI evaluations, SI evaluations, DI evaluations)"""
        v = [C.c_double() for _ in range(4)]
        self.engine._check(self.engine.lib.coffe_b200_plan_count(self._p, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def close(self):
        if self._p:
            self.engine.lib.coffe_b200_plan_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id():
    """128 opaque bytes from rank 0 (ncclGetUniqueId); hand them to every rank's Engine.comm_init"""
    lib = capi.load()
    buf = (C.c_ubyte * 128)()
    rc = lib.coffe_b200_comm_unique_id(buf)
    if rc:
        raise CoffeB200Error(rc, lib.coffe_b200_last_error(None).decode())
    return bytes(buf)


def balance_bounds(cost, n_parts):
    """n_parts + 1 boundaries of contiguous blocks of (nearly) equal total cost (host arithmetic of the library, no GPU needed)"""
    lib = capi.load()
    c = capi.darr(np.asarray(cost).ravel())
    out = (C.c_long * (n_parts + 1))()
    rc = lib.coffe_b200_balance_bounds(capi.ptr(c), C.c_long(c.size), C.c_int(n_parts), out)
    if rc:
        raise ValueError("cost must be finite, non-negative and not all zero; n_parts >= 1")
    return [int(v) for v in out]


def rebalance_cost(cost, bounds, part_ms):
    """fold the measured device time of every block into the cost of its items (coffe_b200_rebalance_cost); returns the new cost"""
    lib = capi.load()
    c = capi.darr(np.asarray(cost).ravel()).copy()
    bd = (C.c_long * len(bounds))(*bounds)
    ms = capi.darr(part_ms)
    rc = lib.coffe_b200_rebalance_cost(capi.ptr(c), C.c_long(c.size), C.c_int(len(bounds) - 1), bd, capi.ptr(ms))
    if rc:
        raise ValueError("bounds must run from 0 to len(cost); times must be non-negative")
    return c


class Group:
    """All GPUs of the box behind one call (coffe_b200_group_*): the counterpart of the reference's OpenMP loop over the output grid."""

    def __init__(self, n_devices=-1):
        self.lib = capi.load()
        self.lib.coffe_b200_group_last_error.restype = C.c_char_p
        self.lib.coffe_b200_group_last_error.argtypes = [C.c_void_p]
        h = C.c_void_p()
        rc = self.lib.coffe_b200_group_create(C.c_int(n_devices), C.byref(h))
        if rc:
            raise CoffeB200Error(rc, self.lib.coffe_b200_group_last_error(None).decode())
        self._h = h
        self.size = self.lib.coffe_b200_group_size(h)
        self._tables = None

    def close(self):
        if getattr(self, "_h", None):
            self.lib.coffe_b200_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise CoffeB200Error(rc, self.lib.coffe_b200_group_last_error(self._h).decode())

    def upload_tables(self, tableset):
        self._tables = tableset
        self._check(self.lib.coffe_b200_group_upload_tables(self._h, C.byref(tableset.struct)))

    def set_options(self, contributions, **flags):
        o = capi.make_options(contributions, **flags)
        self._check(self.lib.coffe_b200_group_set_options(self._h, C.byref(o)))

    def set_quadrature(self, **kw):
        q = capi.make_quadrature(**kw)
        self._check(self.lib.coffe_b200_group_set_quadrature(self._h, C.byref(q)))

    def multipoles(self, z_mean, sep, l):
        z, s, ls = capi.darr(z_mean), capi.darr(sep), capi.iarr(l)
        out = np.empty((len(z), len(ls), len(s)))
        self._check(self.lib.coffe_b200_group_multipoles(self._h, capi.ptr(z), C.c_int(len(z)), capi.ptr(s), C.c_int(len(s)),
                                                         capi.iptr(ls), C.c_int(len(ls)), capi.ptr(out)))
        return out

    def corrfunc(self, z_mean, mu, sep):
        z, m, s = capi.darr(z_mean), capi.darr(mu), capi.darr(sep)
        out = np.empty((len(z), len(m), len(s)))
        self._check(self.lib.coffe_b200_group_corrfunc(self._h, capi.ptr(z), C.c_int(len(z)), capi.ptr(m), C.c_int(len(m)),
                                                       capi.ptr(s), C.c_int(len(s)), capi.ptr(out)))
        return out

    def average_multipoles(self, z_min, z_max, sep, l):
        a, b, s, ls = capi.darr(z_min), capi.darr(z_max), capi.darr(sep), capi.iarr(l)
        out = np.empty((len(a), len(ls), len(s)))
        self._check(self.lib.coffe_b200_group_average_multipoles(self._h, capi.ptr(a), capi.ptr(b), C.c_int(len(a)), capi.ptr(s), C.c_int(len(s)),
                                                                 capi.iptr(ls), C.c_int(len(ls)), capi.ptr(out)))
        return out

    def covariance(self, *args, **kw):
        cp, shape = Engine._covariance_params(self, *args, **kw)
        out = np.empty(shape)
        self._check(self.lib.coffe_b200_group_covariance(self._h, C.byref(cp), capi.ptr(out)))
        return out

    def last_blocks(self):
        """(block boundaries of the last call, device time of every member [ms])"""
        bd, ms = (C.c_long * (self.size + 1))(), (C.c_double * self.size)()
        self._check(self.lib.coffe_b200_group_last_blocks(self._h, bd, ms))
        return [int(v) for v in bd], [float(v) for v in ms]
