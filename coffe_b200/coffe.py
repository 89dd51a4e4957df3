"""`coffe_b200.Coffe`: the host-side mirror of the reference's Cython class `coffe.Coffe`
(python/coffe/coffe.pyx:237-3600) for the signal-evaluation path.

Same property names, same validation errors (ValueError / TypeError raised before any compute call, python/coffe/
coffe.pyx:86-118, 1057-1110), same lazy "background -> integrals -> result" chain with cache invalidation on every
setter, same `compute_*` methods returning the records of `coffe_b200.representation` in the reference's loop order.
Every number a `compute_*` method returns is produced by the sm_100a kernels behind the C ABI
(include/coffe_b200.h): the I^n_l tables by `coffe_b200_integrals`, the signal by `coffe_b200_multipoles` /
`_corrfunc` / `_average_multipoles`, the covariance by `coffe_b200_covariance`.  Only the background (a one-off
O(10^4) host computation, coffe_b200/background.py) and argument checking run on the host; there is no CPU fallback:
without the built library or a GPU the first compute call raises.

`from_file` / `to_file` read and write COFFE settings files (coffe_b200/settings.py).  CLASS itself is not part of this package:
`pk_type` stays "linear", and the reference's CLASS-midpoint mode is reached by handing over what CLASS would have produced
(`set_power_spectrum_2d` + `midpoint_approximation`).  Not mirrored: `cross_spectrum`, the CLASS precision knobs.
"""
import numpy as np

from . import capi
from .background import COFFE_H0, akima_eval, compute_background
from .representation import AverageCovariance, AverageMultipoles, Corrfunc, Covariance, Multipoles

FAMILY = [(0, 0), (0, 2), (0, 4), (1, 1), (1, 3), (2, 0), (2, 2), (3, 1)]
_CONTRIB = {"has_density": "den", "has_rsd": "rsd", "has_lensing": "len", "has_d1": "d1", "has_d2": "d2",
            "has_g1": "g1", "has_g2": "g2", "has_g3": "g3", "has_g4": "g4", "has_g5": "g5"}
_BIASES = ("galaxy_bias1", "galaxy_bias2", "galaxy_bias3", "galaxy_bias4", "magnification_bias1", "magnification_bias2",
           "evolution_bias1", "evolution_bias2")


def _check_parameter(name, value, kind, xmin=None, xmax=None):
    """python/coffe/coffe.pyx:86-109"""
    if isinstance(value, bool) or not isinstance(value, kind):
        if not isinstance(value, kind):
            raise TypeError(f"Expected {kind}, got {type(value)}")
    if xmin is not None and value < xmin:
        raise ValueError(f"Value {value} smaller than xmin ({xmin})")
    if xmax is not None and value > xmax:
        raise ValueError(f"Value {value} larger than xmax ({xmax})")


def _float_list(value, lo, hi, lo_open=True, hi_open=True):
    try:
        iter(value)
    except TypeError as err:
        raise TypeError(f"The value {value} is not iterable") from err
    try:
        temp = [float(v) for v in value]
    except TypeError as err:
        raise TypeError("Cannot convert all values to floats") from err
    for v in temp:
        if not ((v > lo if lo_open else v >= lo) and (v < hi if hi_open else v <= hi)):
            raise ValueError
    return np.array(temp, dtype=float)


class _Scalar:
    """cosmological scalar with range check; changing it invalidates everything but the parameters"""

    def __init__(self, name, default, lo, hi, kind=(int, float)):
        self.name, self.default, self.lo, self.hi, self.kind = name, default, lo, hi, kind

    def __set_name__(self, owner, name):
        self.attr = "_p_" + name

    def __get__(self, obj, owner=None):
        if obj is None:
            return self
        return getattr(obj, self.attr, self.default)

    def __set__(self, obj, value):
        _check_parameter(self.name, value, self.kind, self.lo, self.hi)
        if not np.allclose(value, self.__get__(obj)):
            setattr(obj, self.attr, value)
            obj._free_except_parameters()


class _Flag:
    def __init__(self, default, frees):
        self.default, self.frees = default, frees

    def __set_name__(self, owner, name):
        self.attr = "_p_" + name

    def __get__(self, obj, owner=None):
        if obj is None:
            return self
        return bool(getattr(obj, self.attr, self.default))

    def __set__(self, obj, value):
        setattr(obj, self.attr, bool(value))
        for f in self.frees:
            getattr(obj, f)()


class _Array:
    """list-valued coordinate / survey parameter (python/coffe/coffe.pyx:1551-2440)"""

    def __init__(self, default, lo, hi, frees, lo_open=True, hi_open=True):
        self.default, self.lo, self.hi, self.frees, self.lo_open, self.hi_open = default, lo, hi, frees, lo_open, hi_open

    def __set_name__(self, owner, name):
        self.attr = "_p_" + name

    def __get__(self, obj, owner=None):
        if obj is None:
            return self
        return np.array(getattr(obj, self.attr, self.default), dtype=float)

    def __set__(self, obj, value):
        setattr(obj, self.attr, _float_list(value, self.lo, self.hi, self.lo_open, self.hi_open))
        for f in self.frees:
            getattr(obj, f)()


_SIGNAL = ("_free_corrfunc", "_free_multipoles", "_free_average_multipoles")
_COV = ("_free_covariance_multipoles", "_free_average_covariance_multipoles")


class Coffe:
    # ---- cosmology (defaults: coffe_parse_default_parameters, src/parser.c:743-762) ----
    h = _Scalar("h", 0.67, 0, 100)
    w0 = _Scalar("w0", -1.0, -100, 100)
    wa = _Scalar("wa", 0.0, -100, 100)
    n_s = _Scalar("n_s", 0.96, 0, 2)
    sigma8 = _Scalar("sigma8", 0.8156, 0, 2)
    T_cmb = _Scalar("T_cmb", 2.726, 0, 10)
    N_ur = _Scalar("N_ur", 2.0328, 0, 10)
    # ---- coordinates ----
    sep = _Array(np.arange(5.0, 305.0, 3.0), 0, 10000, ("_free_corrfunc", "_free_multipoles"))     # coffe_generate_range(5, 305, 100)
    mu = _Array([0.5], -1, 1, ("_free_corrfunc",))
    z_mean = _Array([1.0], 0, 10, ("_free_corrfunc", "_free_multipoles", "_free_covariance_multipoles"), lo_open=False)
    deltaz = _Array([0.2], 0, 10, ("_free_integrals",) + _SIGNAL + _COV, lo_open=False)
    z_min = _Array([], 0, 10, ("_free_average_multipoles", "_free_average_covariance_multipoles"), lo_open=False)
    z_max = _Array([], 0, 10, ("_free_average_multipoles", "_free_average_covariance_multipoles"), lo_open=False)
    number_density1 = _Array([1e-3], 0, 1e30, _COV, lo_open=False)
    number_density2 = _Array([1e-3], 0, 1e30, _COV, lo_open=False)
    pixelsize = _Array([10.0], 0, 1e4, _COV)
    fsky = _Array([1.0], 0, 1, _COV, hi_open=False)
    # ---- switches ----
    has_density = _Flag(True, _SIGNAL + _COV)
    has_rsd = _Flag(False, _SIGNAL + _COV)
    has_lensing = _Flag(False, _SIGNAL)
    has_d1 = _Flag(False, _SIGNAL)
    has_d2 = _Flag(False, _SIGNAL + ("_free_integrals",))
    has_g1 = _Flag(False, _SIGNAL + ("_free_integrals",))
    has_g2 = _Flag(False, _SIGNAL + ("_free_integrals",))
    has_g3 = _Flag(False, _SIGNAL + ("_free_integrals",))
    has_g4 = _Flag(False, _SIGNAL + ("_free_integrals",))
    has_g5 = _Flag(False, _SIGNAL + ("_free_integrals",))
    has_only_cross_correlations = _Flag(False, _SIGNAL)
    has_flatsky_local = _Flag(False, _SIGNAL)
    has_flatsky_local_nonlocal = _Flag(False, _SIGNAL + ("_free_integrals",))
    has_flatsky_nonlocal = _Flag(False, _SIGNAL + ("_free_integrals",))
    has_binned_covariance = _Flag(False, _COV)
    covariance_cosmic = _Flag(True, _COV)
    covariance_mixed = _Flag(True, _COV)
    covariance_poisson = _Flag(True, _COV)
    verbose = _Flag(False, ())

    def __init__(self, device=0, **kwargs):
        self._device = device
        self._engine = None
        self._omega = {"m": 0.3, "baryon": 0.05, "gamma": 9e-5, "nu": 0.0}
        self._l = np.array([0, 2, 4], dtype=int)
        self._populations = [1, 1, 1, 1]
        self._bias = {name: (np.arange(0.0, 11.0), np.full(11, 1.0 if name.startswith("galaxy") else 0.0)) for name in _BIASES}
        self._pk = None
        self._pk2d = None                 # (z [n_z], P(k, z_i) [n_z, n_k]) for the midpoint approximation
        self._midpoint = False
        self._k_min, self._k_max = 1e-5, 300.0
        self._bessel_bins, self._background_bins = 10000, 10000
        self._cov_method, self._cov_bins = "standard", 8000
        self._background = None
        self._integrals = None
        self._tables_uploaded = False
        self._cache = {}
        if kwargs:
            self.set_parameters(**kwargs)

    @staticmethod
    def from_file(filename):
        """python/coffe/coffe.pyx:537-793: a Coffe object with the parameters of a COFFE settings file"""
        from .settings import apply_settings, read_settings
        return apply_settings(Coffe(), read_settings(filename))

    def to_file(self, filename):
        """python/coffe/coffe.pyx:794-866"""
        from .settings import write_settings
        write_settings(self, filename)

    # ---- cache invalidation, named as in python/coffe/coffe.pyx:948-993 -------------------------------
    def _free_background(self):
        self._background = None
        self._tables_uploaded = False

    def _free_power_spectrum(self):
        self._pk = None

    def _free_integrals(self):
        self._integrals = None
        self._tables_uploaded = False

    def _free_corrfunc(self): self._cache.pop("corrfunc", None)
    def _free_multipoles(self): self._cache.pop("multipoles", None)
    def _free_average_multipoles(self): self._cache.pop("average_multipoles", None)
    def _free_covariance_multipoles(self): self._cache.pop("covariance", None)
    def _free_average_covariance_multipoles(self): self._cache.pop("average_covariance", None)

    def _free_except_parameters(self):
        self._free_background()
        self._free_integrals()
        self._cache.clear()

    # ---- Omegas (python/coffe/coffe.pyx:1172-1266) ---------------------------------------------------
    @staticmethod
    def _check_omegas(*omegas):
        if np.sum(np.array(omegas)) > 1:
            raise ValueError("Cannot satisfy the condition Omega_de >= 0")

    @property
    def omega_m(self): return self._omega["m"]

    @omega_m.setter
    def omega_m(self, value):
        _check_parameter("omega_m", value, (int, float), -100, 100)
        if not np.allclose(value, self.omega_m):
            self._check_omegas(value, self.omega_gamma)
            self._omega["m"] = value
            self._free_except_parameters()

    @property
    def omega_baryon(self): return self._omega["baryon"]

    @omega_baryon.setter
    def omega_baryon(self, value):
        _check_parameter("omega_baryon", value, (int, float), -100, 100)
        if not np.allclose(value, self.omega_baryon):
            self._check_omegas(value, self.omega_cdm, self.omega_gamma)
            self._omega["baryon"] = value
            self._free_except_parameters()

    @property
    def omega_cdm(self): return self._omega["m"] - self._omega["baryon"] - self._omega["nu"]

    @omega_cdm.setter
    def omega_cdm(self, value):
        _check_parameter("omega_cdm", value, (int, float), -100, 100)
        if not np.allclose(value, self.omega_cdm):
            self._check_omegas(value, self.omega_baryon, self.omega_gamma, self.omega_nu)
            self._omega["m"] = value + self._omega["baryon"]          # Omega0_m = Omega0_cdm + Omega0_baryon (:1185)
            self._free_except_parameters()

    @property
    def omega_gamma(self): return self._omega["gamma"]

    @omega_gamma.setter
    def omega_gamma(self, value):
        _check_parameter("omega_gamma", value, (int, float), -100, 100)
        if not np.allclose(value, self.omega_gamma):
            self._check_omegas(value, self.omega_m)
            self._omega["gamma"] = value
            self._free_except_parameters()

    @property
    def omega_nu(self): return self._omega["nu"]

    @property
    def omega_de(self): return 1 - self._omega["m"] - self._omega["gamma"]

    @property
    def has_cuba(self): return False          # fixed-node GPU quadrature; nothing to link

    @property
    def has_class(self): return False

    @property
    def pk_type(self): return "linear"

    @pk_type.setter
    def pk_type(self, value):
        if value != "linear":
            raise ValueError("no Boltzmann code in coffe_b200: keep pk_type = 'linear' and, for the reference's CLASS-midpoint mode, pass "
                             "P(k, z) with set_power_spectrum_2d and set midpoint_approximation = True")

    # ---- coordinates with integer / special handling ----------------------------------------------------
    @property
    def l(self): return np.array(self._l)

    @l.setter
    def l(self, value):
        try:
            iter(value)
        except TypeError as err:
            raise TypeError(f"The value {value} is not iterable") from err
        try:
            temp = [v.__index__() for v in value]
        except AttributeError as err:
            raise TypeError("Cannot convert all values to int") from err
        if not all(0 <= v < 10 for v in temp):
            raise ValueError
        self._l = np.array(temp, dtype=int)
        self._free_integrals()      # the flat-sky tables depend on l
        for f in ("_free_multipoles", "_free_average_multipoles") + _COV:
            getattr(self, f)()

    @property
    def covariance_populations(self): return np.array(self._populations)

    @covariance_populations.setter
    def covariance_populations(self, value):
        try:
            iter(value)
        except TypeError as err:
            raise TypeError(f"The value {value} is not iterable") from err
        if len(value) != 4:
            raise ValueError("The array for setting `covariance_populations` must have length 4")
        if any(v not in [1, 2, 3, 4] for v in value):
            raise ValueError("The values for `covariance_populations` must be one of: [1, 2, 3, 4]")
        self._populations = [int(v) for v in value]
        for f in _COV:
            getattr(self, f)()

    @property
    def covariance_integration_method(self): return self._cov_method

    @covariance_integration_method.setter
    def covariance_integration_method(self, value):
        """'standard' (k-quadrature) or 'fftlog' (2-D FFTLog), python/coffe/coffe.pyx:2243-2272"""
        if value not in ("standard", "fftlog"):
            raise ValueError(f"The value `covariance_integration_method={value}` is not one of `['standard', 'fftlog']")
        self._cov_method = value
        for f in _COV:
            getattr(self, f)()

    @property
    def covariance_integration_sampling(self): return self._cov_bins

    @covariance_integration_sampling.setter
    def covariance_integration_sampling(self, value):
        """number of points of the 2-D FFTLog (python/coffe/coffe.pyx:2275-2296); default 8000"""
        _check_parameter("covariance_integration_sampling", value, int, 0, 2147483647)
        self._cov_bins = int(value)
        for f in _COV:
            getattr(self, f)()

    @property
    def integrals_fftlog_sampling(self): return self._bessel_bins

    @integrals_fftlog_sampling.setter
    def integrals_fftlog_sampling(self, value):
        _check_parameter("integrals_fftlog_sampling", value, int, 0, 1000000)
        self._bessel_bins = int(value)
        self._free_except_parameters()

    @property
    def background_sampling(self): return self._background_bins

    @background_sampling.setter
    def background_sampling(self, value):
        _check_parameter("background_sampling", value, int, 0, 2147483647)
        self._background_bins = int(value)
        self._free_except_parameters()

    @property
    def k_min(self): return self._k_min

    @k_min.setter
    def k_min(self, value):
        _check_parameter("k_min", value, (int, float), 0, 1e-3)
        self._k_min = value
        self._free_except_parameters()

    @property
    def k_max(self): return self._k_max

    @k_max.setter
    def k_max(self, value):
        _check_parameter("k_max", value, (int, float), 0.1, 1e3)
        self._k_max = value
        self._free_except_parameters()

    # ---- bulk parameter access (python/coffe/coffe.pyx:1137-1169) --------------------------------------
    @property
    def parameters(self):
        out = {}
        for key in dir(type(self)):
            if key == "parameters" or key.startswith("_"):
                continue
            attr = getattr(type(self), key)
            if isinstance(attr, (_Scalar, _Flag, _Array)) or (isinstance(attr, property) and attr.fset is not None):
                out[key] = getattr(self, key)
        return out

    def set_parameters(self, **value):
        for key in value:
            if not hasattr(self, key):
                raise AttributeError(f"The parameter {key} is not a valid parameter in COFFE.")
        for key in value:
            setattr(self, key, value[key])

    def reset_contributions(self):
        for name in _CONTRIB:
            setattr(self, name, False)

    # ---- biases (python/coffe/coffe.pyx:158-193, 1716-1933) ----------------------------------------------
    def _set_bias(self, name, x_sampling, y_sampling):
        if not np.allclose(x_sampling, np.sort(x_sampling)):
            raise ValueError("The input array must be sorted")
        if len(x_sampling) != len(y_sampling):
            raise ValueError(f"Mismatching lengths for x ({len(x_sampling)}) and y ({len(y_sampling)})")
        _check_parameter("xmin", float(x_sampling[0]), (int, float), xmin=0)
        _check_parameter("xmax", float(x_sampling[-1]), (int, float), xmax=15)
        self._bias[name] = (np.array(x_sampling, dtype=float), np.array(y_sampling, dtype=float))
        if name.startswith(("magnification", "evolution")):
            self._free_background()      # G1, G2 depend on them (src/background.c:505-533)
        self._tables_uploaded = False
        self._cache.clear()

    def _eval_bias(self, name, z):
        x, y = self._bias[name]
        if not (x[0] <= z <= x[-1]):
            raise ValueError(f"Value {z} outside the sampled range [{x[0]}, {x[-1]}]")
        return float(akima_eval(x, y, z)[0])

    # ---- power spectrum ------------------------------------------------------------------------------
    def set_power_spectrum_linear(self, k, pk, z=0):
        """python/coffe/coffe.pyx:2929-2995: k in 1/Mpc, P in Mpc^3; resets k_min / k_max to the ends of the table"""
        if not np.allclose(k, np.sort(k)):
            raise ValueError("The input k array must be sorted")
        if len(k) != len(pk):
            raise ValueError("The input arrays have mismatching lengths")
        self._pk = (np.array(k, dtype=float), np.array(pk, dtype=float))
        self._k_min, self._k_max = float(k[0]), float(k[-1])
        self._free_integrals()
        self._cache.clear()

    def set_power_spectrum_2d(self, k, z, pk2d, k_min=None, k_max=None):
        """P(k, z_i) on the knots k (1/Mpc, Mpc^3), rows pk2d[i] = P(k, z_i): the input of the midpoint approximation.  In the reference
        these rows come out of CLASS inside parse_external_power_spectrum (src/parser.c:655-709, 100 redshifts z_i = 3 i/100) when
        `pk_type` is not linear; here the host hands them over (no Boltzmann code in this package).  Needs the 1-D spectrum too
        (set_power_spectrum_linear: CLASS's linear P(k, 0)), which the potential terms keep using."""
        k, z, pk2d = np.array(k, dtype=float), np.array(z, dtype=float), np.array(pk2d, dtype=float)
        if pk2d.shape != (len(z), len(k)):
            raise ValueError("pk2d must have shape (len(z), len(k))")
        if self._pk is None or len(self._pk[0]) != len(k) or not np.allclose(self._pk[0], k):
            raise ValueError("set_power_spectrum_linear must be called first, with the same k knots")
        self._pk2d = (z, pk2d)
        if k_min is not None:
            self._k_min = float(k_min)
        if k_max is not None:
            self._k_max = float(k_max)
        self._free_integrals()
        self._cache.clear()

    @property
    def midpoint_approximation(self):
        """python/coffe/coffe.pyx `midpoint_approximation` (par->midpoint_approximation, used with pk_type != linear):
        the standard terms read I^n_l((z_a + z_b)/2, r) built from P(k, z) instead of D1 D1 I^n_l(r)"""
        return self._midpoint

    @midpoint_approximation.setter
    def midpoint_approximation(self, value):
        self._midpoint = bool(value)
        self._free_integrals()
        self._cache.clear()

    def _power_spectrum_init(self):
        """default spectrum: the PkL_CLASS.dat table shipped with the package, k h and P / h^3 (src/parser.c:811-829)"""
        from pathlib import Path
        k, pk = np.loadtxt(Path(__file__).resolve().parent / "data" / "PkL_CLASS.dat", unpack=True)
        self._pk = (k * self.h, pk / self.h ** 3)

    def power_spectrum(self, k, z):
        _check_parameter("k", k, (int, float), self.k_min, self.k_max)
        _check_parameter("z", z, (int, float), 0, 15)
        if self._pk is None:
            self._power_spectrum_init()
        return float(akima_eval(self._pk[0], self._pk[1], k)[0]) * self.growth_factor(z) ** 2

    # ---- background accessors (python/coffe/coffe.pyx:2998-3156) -------------------------------------------
    def _background_init(self):
        self._background = compute_background(
            self.omega_cdm + self.omega_baryon, self.omega_gamma, self.omega_de, self.h, self.w0, self.wa, n=self._background_bins,
            magnification_bias1=self._bias["magnification_bias1"], magnification_bias2=self._bias["magnification_bias2"],
            evolution_bias1=self._bias["evolution_bias1"], evolution_bias2=self._bias["evolution_bias2"])
        self._tables_uploaded = False

    def _bg(self, name, z):
        _check_parameter("z", z, (int, float), 0, 15)
        if self._background is None:
            self._background_init()
        x, y = self._background[name]
        return float(akima_eval(x, y, z)[0])

    def comoving_distance(self, z): return self._bg("comoving_distance", z)
    def hubble_rate(self, z): return self._bg("Hz", z)
    def scale_factor(self, z): return self._bg("a", z)
    def hubble_rate_conformal(self, z): return self._bg("conformal_Hz", z)
    def hubble_rate_conformal_derivative(self, z): return self._bg("conformal_Hz_prime", z)
    def growth_factor(self, z): return self._bg("D1", z)
    def growth_rate(self, z): return self._bg("f", z)

    def maximum_separation(self, z_mean, deltaz):
        _check_parameter("z_mean", z_mean, (int, float), 0, 15)
        _check_parameter("deltaz", deltaz, (int, float), 0, z_mean)
        return 2 * (self.comoving_distance(z_mean + deltaz) - self.comoving_distance(z_mean))

    # ---- checks (python/coffe/coffe.pyx:1064-1110) --------------------------------------------------------
    def _check_coords_corrfunc(self):
        if not (len(self.mu) and len(self.sep) and len(self.z_mean)):
            raise ValueError(f"At least one of the following parameters is empty: mu ({self.mu}), sep ({self.sep}), z_mean ({self.z_mean})")

    def _check_coords_multipoles(self):
        if not (len(self.l) and len(self.sep) and len(self.z_mean)):
            raise ValueError(f"At least one of the following parameters is empty: l ({self.l}), sep ({self.sep}), z_mean ({self.z_mean})")

    def _check_coords_average_multipoles(self):
        if not (len(self.l) and len(self.sep) and len(self.z_min) and len(self.z_max)):
            raise ValueError(f"At least one of the following parameters is empty: l ({self.l}), sep ({self.sep}), z_min ({self.z_min}), z_max ({self.z_max})")

    def _contributions(self):
        return [short for name, short in _CONTRIB.items() if getattr(self, name)]

    def _check_contributions(self):
        if not self._contributions():
            raise ValueError("No contributions specified, you need to specify at least one of: 'has_density', 'has_rsd', 'has_lensing, "
                             "'has_d1', 'has_d2', 'has_g1', 'has_g2', 'has_g3', 'has_g4', 'has_g5'")

    @property
    def _divergent(self):
        return any(getattr(self, n) for n in ("has_d2", "has_g1", "has_g2", "has_g3", "has_g4", "has_g5"))

    # ---- device state -----------------------------------------------------------------------------------
    def _get_engine(self):
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(self._device)      # raises without the built library or without a GPU
        return self._engine

    def _table_list(self):
        """the (n, l, half) tables coffe_integrals_init would build for the current switches (src/integrals.c:597-1413)"""
        tables = [(n, l, 0) for n, l in FAMILY]
        if self.has_flatsky_local_nonlocal or self.has_flatsky_nonlocal:
            tables.append((1, -1, 1))
        if self.has_flatsky_nonlocal:
            tables += [(1, int(l), 0) for l in self._l if (1, int(l), 0) not in tables]
        if self.has_flatsky_local_nonlocal:
            for l in self._l:
                for k in range(int(l) // 2 + 1):
                    t = (int(l) - 2 * k + 3, int(l) - 2 * k + 1, 1)
                    if t not in tables:
                        tables.append(t)
        return tables

    def _integrals_init(self):
        if self._background is None:
            self._background_init()
        if self._pk is None:
            self._power_spectrum_init()
        eng = self._get_engine()
        tables = self._table_list()
        chi_max = 0.0
        if self._divergent:
            # the renormalisation grid always ends at chi(max(z_mean + deltaz)) from Python (SURVEY.md 8a note 7)
            if len(self.z_mean) != len(self.deltaz):
                raise ValueError("Mismatching lengths for `z_mean` and `deltaz`")
            chi_max = self.comoving_distance(float(np.max(self.z_mean + self.deltaz)))
        res = eng.integrals(self._pk, self.k_min, self.k_max, tables, bessel_bins=self._bessel_bins, divergent=self._divergent, chi_max=chi_max)
        # the half-integer (1, -1) table is stored rescaled by sqrt(2/pi) 2 pi^2 (src/integrals.c:955-961)
        for i, t in enumerate(tables):
            if t == (1, -1, 1):
                res["tables"][i] = res["tables"][i] * np.sqrt(2.0 / np.pi) * 2 * np.pi ** 2
        res["keys"] = tables
        res["z2d"] = res["tables2d"] = None
        if self._midpoint:
            if self._pk2d is None:
                raise ValueError("midpoint_approximation needs P(k, z): call set_power_spectrum_2d")
            fam = [(n, l) for n, l in FAMILY]
            _, v = eng.integrals2d(self._pk[0], self._pk2d[1], self.k_min, self.k_max, fam, bessel_bins=self._bessel_bins)
            res["z2d"], res["tables2d"] = self._pk2d[0], {key: v[i] for i, key in enumerate(fam)}
        self._integrals = res
        self._tables_uploaded = False

    def _tableset(self):
        """the coffe_b200_tables the drivers hand to the library: background, biases, the I^n_l family and the divergent extras"""
        if self._background is None:
            self._background_init()
        if self._integrals is None or (self._divergent and self._integrals["i40"] is None):
            self._integrals_init()
        res = self._integrals
        integrals = [(n, l, half, res["r"], res["tables"][i]) for i, (n, l, half) in enumerate(res["keys"])]
        renorm = renorm0 = None
        if res["i40"] is not None:
            integrals.append((4, 0, 0, res["r40"], res["i40"]))
            renorm = (res["renorm_chi"], res["renorm_z"])
            x0 = res["r40"].copy()
            x0[0] = 0.0
            renorm0 = (x0, res["renorm0"])
        i2d = (res["z2d"], res["tables2d"]) if res.get("tables2d") is not None else None
        return capi.TableSet(self._background, self._bias, integrals, self.omega_cdm + self.omega_baryon, self.h, interp_method=5,
                             renorm=renorm, renorm0=renorm0, integrals2d=i2d)

    def _upload(self):
        if self._tables_uploaded and self._background is not None and self._integrals is not None and \
                not (self._divergent and self._integrals["i40"] is None):
            return
        ts = self._tableset()
        self._get_engine().upload_tables(ts)
        self._tables_uploaded = True

    def _prepare_signal(self):
        self._upload()
        self._get_engine().set_options(self._contributions(), only_cross=self.has_only_cross_correlations, flatsky_local=self.has_flatsky_local,
                                       flatsky_local_nonlocal=self.has_flatsky_local_nonlocal, flatsky_nonlocal=self.has_flatsky_nonlocal,
                                       midpoint=self._midpoint)

    def integral(self, *, r, n, l):
        """r^(n-l) I^n_l(r) for n > l, I^n_l(r) otherwise (python/coffe/coffe.pyx:1657-1713)"""
        _check_parameter("r", r, (int, float), 0, 25000)
        _check_parameter("n", n, int, -2, 4)
        _check_parameter("l", l, int, 0, 4)
        if self._integrals is None:
            self._integrals_init()
        keys = self._integrals["keys"]
        if (n, l, 0) not in keys:
            raise ValueError(f"Cannot find integral with n = {n}, l = {l}")
        return float(akima_eval(self._integrals["r"], self._integrals["tables"][keys.index((n, l, 0))], r)[0])

    # ---- signal ---------------------------------------------------------------------------------------
    def compute_corrfunc(self, *, z, r, mu, recompute=False):
        self._check_coords_corrfunc()
        self._check_contributions()
        _check_parameter("z", z, (int, float), 0, 15)
        _check_parameter("r", r, (int, float), 0, 1500)
        _check_parameter("mu", mu, (int, float), -1, 1)
        if recompute:
            self._free_except_parameters()
        self._prepare_signal()
        return float(self._get_engine().corrfunc([z], [mu], [r])[0, 0, 0])

    def compute_multipole(self, *, z, r, l, recompute=False):
        self._check_coords_multipoles()
        self._check_contributions()
        _check_parameter("z", z, (int, float), 0, 15)
        _check_parameter("r", r, (int, float), 0, 1500)
        _check_parameter("l", l, (int,), 0, 10)
        if recompute:
            self._free_except_parameters()
        self._prepare_signal()
        return float(self._get_engine().multipoles([z], [r], [l])[0, 0, 0])

    def compute_multipoles_array(self, recompute=False):
        """[n_z, n_l, n_sep] array in the order of coffe_multipoles_init (src/multipoles.c:75-86): the bulk result
        without one Python object per point"""
        self._check_coords_multipoles()
        self._check_contributions()
        if recompute:
            self._free_except_parameters()
        if "multipoles" not in self._cache:
            self._prepare_signal()
            self._cache["multipoles"] = self._get_engine().multipoles(self.z_mean, self.sep, self.l)
        return self._cache["multipoles"]

    def compute_multipoles_bulk(self, recompute=False):
        v = self.compute_multipoles_array(recompute)
        return np.array([Multipoles(z=z, r=r, l=int(l), value=v[iz, il, ir])
                         for iz, z in enumerate(self.z_mean) for il, l in enumerate(self.l) for ir, r in enumerate(self.sep)])

    def compute_corrfunc_array(self, recompute=False):
        """[n_z, n_mu, n_sep] in the order of coffe_corrfunc_init (src/corrfunc.c:81-93)"""
        self._check_coords_corrfunc()
        self._check_contributions()
        if recompute:
            self._free_except_parameters()
        if "corrfunc" not in self._cache:
            self._prepare_signal()
            self._cache["corrfunc"] = self._get_engine().corrfunc(self.z_mean, self.mu, self.sep)
        return self._cache["corrfunc"]

    def compute_corrfunc_bulk(self, recompute=False):
        v = self.compute_corrfunc_array(recompute)
        return np.array([Corrfunc(z=z, r=r, mu=m, value=v[iz, im, ir])
                         for iz, z in enumerate(self.z_mean) for im, m in enumerate(self.mu) for ir, r in enumerate(self.sep)])

    def compute_average_multipoles_array(self, recompute=False):
        """[n_bins, n_l, n_sep] in the order of coffe_average_multipoles_init (src/average_multipoles.c:97-110)"""
        self._check_coords_average_multipoles()
        self._check_contributions()
        if len(self.z_min) != len(self.z_max):
            raise ValueError("Mismatching lengths for `z_min` and `z_max`")
        if recompute:
            self._free_except_parameters()
        if "average_multipoles" not in self._cache:
            self._prepare_signal()
            self._cache["average_multipoles"] = self._get_engine().average_multipoles(self.z_min, self.z_max, self.sep, self.l)
        return self._cache["average_multipoles"]

    def compute_average_multipoles_bulk(self, recompute=False):
        v = self.compute_average_multipoles_array(recompute)
        return np.array([AverageMultipoles(z_min=a, z_max=b, r=r, l=int(l), value=v[ib, il, ir])
                         for ib, (a, b) in enumerate(zip(self.z_min, self.z_max)) for il, l in enumerate(self.l) for ir, r in enumerate(self.sep)])

    # ---- covariance (python/coffe/coffe.pyx:3428-3535) ----------------------------------------------------
    def _covariance(self, average):
        self._check_coords_multipoles()
        self._check_contributions()
        first = self.z_min if average else self.z_mean
        others = (self.deltaz, self.number_density1, self.number_density2, self.pixelsize, self.fsky) + ((self.z_max,) if average else ())
        if not all(len(first) == len(o) for o in others):
            raise ValueError("Mismatching lengths for covariance parameters (`z_mean`, `deltaz`, `number_density1`, `number_density2`, `pixelsize`, `fsky`")
        if self._pk is None:
            self._power_spectrum_init()
        if average:
            [self.maximum_separation((a + b) / 2, (b - a) / 2) for a, b in zip(self.z_min, self.z_max)]
        else:
            [self.maximum_separation(float(z), float(d)) for z, d in zip(self.z_mean, self.deltaz)]
        key = "average_covariance" if average else "covariance"
        if key not in self._cache:
            if self._background is None:
                self._background_init()
            eng = self._get_engine()
            # the covariance reads the background and galaxy biases only: upload them without the I^n_l tables if those are not built
            if self._integrals is not None:
                self._upload()
            else:
                eng.upload_tables(capi.TableSet(self._background, self._bias, [], self.omega_cdm + self.omega_baryon, self.h, interp_method=5))
                self._tables_uploaded = False
            eng.set_options(self._contributions(), only_cross=self.has_only_cross_correlations)
            kw = dict(z_min=self.z_min, z_max=self.z_max) if average else dict(z_mean=self.z_mean, deltaz=self.deltaz)
            self._cache[key] = eng.covariance(self._pk, self.sep, self.l, self.fsky, self.number_density1, self.number_density2, self.pixelsize,
                                              pops=tuple(self._populations), cosmic=self.covariance_cosmic, mixed=self.covariance_mixed,
                                              poisson=self.covariance_poisson, window=self.has_binned_covariance,
                                              k_min=self.k_min, k_max=self.k_max, method=self._cov_method, bins=self._cov_bins, **kw)
        return self._cache[key]

    def compute_covariance_array(self, recompute=False):
        """[n_z, n_l, n_sep, n_l, n_sep] in the order of coffe_covariance_init (src/covariance.c:1046-1059)"""
        if recompute:
            self._free_except_parameters()
        return self._covariance(False)

    def compute_covariance_bulk(self, recompute=False):
        v = self.compute_covariance_array(recompute)
        return np.array([Covariance(z=z, r1=r1, r2=r2, l1=int(l1), l2=int(l2), value=v[iz, i1, j1, i2, j2])
                         for iz, z in enumerate(self.z_mean) for i1, l1 in enumerate(self.l) for j1, r1 in enumerate(self.sep)
                         for i2, l2 in enumerate(self.l) for j2, r2 in enumerate(self.sep)])

    def compute_average_covariance_array(self, recompute=False):
        if recompute:
            self._free_except_parameters()
        return self._covariance(True)

    def compute_average_covariance_bulk(self, recompute=False):
        v = self.compute_average_covariance_array(recompute)
        return np.array([AverageCovariance(z_min=a, z_max=b, r1=r1, r2=r2, l1=int(l1), l2=int(l2), value=v[ib, i1, j1, i2, j2])
                         for ib, (a, b) in enumerate(zip(self.z_min, self.z_max)) for i1, l1 in enumerate(self.l) for j1, r1 in enumerate(self.sep)
                         for i2, l2 in enumerate(self.l) for j2, r2 in enumerate(self.sep)])


def _make_bias_methods():
    for name in _BIASES:
        def setter(self, x_sampling, y_sampling, _n=name):
            self._set_bias(_n, x_sampling, y_sampling)

        def getter(self, z, _n=name):
            return self._eval_bias(_n, z)

        setter.__doc__ = f"Sets the {name.replace('_', ' ')} spline from redshifts and values (python/coffe/coffe.pyx:1728-1933)."
        getter.__doc__ = f"Evaluates the {name.replace('_', ' ')} at redshift z."
        setattr(Coffe, "set_" + name, setter)
        setattr(Coffe, name, getter)


_make_bias_methods()
