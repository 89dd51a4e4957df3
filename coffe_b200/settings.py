"""COFFE settings files (`key = value;` lines in libconfig style, `#` comments): the counterpart of `Coffe.from_file` / `Coffe.to_file`
(python/coffe/coffe.pyx:537-793, 794-866) and, for the subset the signal path reads, of the C parser (src/parser.c:1340-1690).

Table driven: SCHEMA maps every recognised option to the kind of value it holds; RENAMED maps the options of older files to today's
property names (the reference warns and forwards them the same way).  Unknown keys are ignored, like the reference does."""
import re
import warnings

import numpy as np

FLOAT_ARRAYS = ("sep", "z_mean", "deltaz", "z_min", "z_max", "mu", "number_density1", "number_density2", "pixelsize", "fsky")
INT_ARRAYS = ("l",)
FLOATS = ("omega_m", "omega_baryon", "omega_gamma", "w0", "wa", "h", "sigma8", "n_s", "T_cmb", "A_s", "N_ur", "m_ncdm", "k_min", "k_max", "YHe")
INTS = ("has_density", "has_lensing", "has_rsd", "has_d1", "has_d2", "has_g1", "has_g2", "has_g3", "has_g4", "has_g5", "integration_sampling",
        "background_sampling", "has_binned_covariance", "has_flatsky_local", "has_flatsky_nonlocal", "has_flatsky_local_nonlocal",
        "covariance_cosmic", "covariance_mixed", "covariance_poisson", "covariance_integration_sampling", "integrals_fftlog_sampling",
        "has_only_cross_correlations", "N_ncdm")
INT_OR_NAME = ("covariance_integration_method", "pk_type", "covariance_interpolation_method", "integration_1d_type")
SCHEMA = {**{k: "float_array" for k in FLOAT_ARRAYS}, **{k: "int_array" for k in INT_ARRAYS}, **{k: "float" for k in FLOATS},
          **{k: "int" for k in INTS}, **{k: "int_or_name" for k in INT_OR_NAME}}
# option of an older file -> (kind, properties it sets today; empty = dropped)
RENAMED = {
    "only_cross_correlations": ("int", ("has_only_cross_correlations",)), "covariance_z_mean": ("float_array", ("z_mean",)),
    "covariance_deltaz": ("float_array", ("deltaz",)), "covariance_fsky": ("float_array", ("fsky",)),
    "covariance_pixelsize": ("float_array", ("pixelsize",)), "covariance_zmin": ("float_array", ("z_min",)),
    "covariance_zmax": ("float_array", ("z_max",)), "covariance_window": ("int", ("has_binned_covariance",)),
    "covariance_density": ("float_array", ("number_density1", "number_density2")), "multipoles": ("int_array", ("l",)),
    "flatsky_local": ("int", ("has_flatsky_local",)), "flatsky_local_nonlocal": ("int", ("has_flatsky_local_nonlocal",)),
    "flatsky_nonlocal": ("int", ("has_flatsky_nonlocal",)), "covariance_step_size": ("float", ()),
    "covariance_minimum_separation": ("float", ()), "covariance_integration_bins": ("int", ()),
}
CONTRIBUTION_FLAGS = {"den": "has_density", "rsd": "has_rsd", "len": "has_lensing", "d1": "has_d1", "d2": "has_d2", "g1": "has_g1",
                      "g2": "has_g2", "g3": "has_g3", "g4": "has_g4", "g5": "has_g5"}
BIAS_FAMILIES = (("galaxy", (1, 2, 3, 4)), ("magnification", (1, 2)), ("evolution", (1, 2)))
_LINE = re.compile(r"^\s*([A-Za-z_][A-Za-z_0-9]*)\s*[=:]\s*(.*?)\s*;?\s*$")


def read_settings(path):
    """{option: raw text of its value} of a settings file; comments and blank lines dropped, values still unparsed"""
    out = {}
    with open(path, encoding="utf-8") as f:
        for line in f:
            line = line.split("#", 1)[0].strip()
            m = _LINE.match(line) if line else None
            if m:
                out[m.group(1)] = m.group(2)
    return out


def _items(text):
    return [t.strip().strip('"').strip() for t in text.strip().strip("[]()").split(",") if t.strip()]


def convert(kind, text):
    if kind == "float":
        return float(text)
    if kind == "int":
        return int(float(text))
    if kind == "float_array":
        return [float(t) for t in _items(text)]
    if kind == "int_array":
        return [int(float(t)) for t in _items(text)]
    if kind == "str_array":
        return _items(text)
    if kind == "int_or_name":
        try:
            return int(text)
        except ValueError:
            return text.strip().strip('"')
    raise ValueError(kind)


def apply_settings(cosmo, raw):
    """set the properties of a `Coffe` object from the raw options: renamed options first, then today's names (which win)"""
    for old, (kind, targets) in RENAMED.items():
        if old not in raw:
            continue
        if not targets:
            warnings.warn(f"Option '{old}' is no longer supported and will be ignored.", DeprecationWarning)
            continue
        warnings.warn(f"Option '{old}' is no longer supported; it may be replaced by {list(targets)} (if found).", DeprecationWarning)
        for t in targets:
            _assign(cosmo, t, convert(kind, raw[old]))
    if "correlation_contributions" in raw:
        warnings.warn("Option 'correlation_contributions' is no longer supported; use the has_* switches.", DeprecationWarning)
        for name in convert("str_array", raw["correlation_contributions"]):
            if name in CONTRIBUTION_FLAGS:
                setattr(cosmo, CONTRIBUTION_FLAGS[name], True)
    for key, kind in SCHEMA.items():
        if key in raw:
            _assign(cosmo, key, convert(kind, raw[key]))
    for family, populations in BIAS_FAMILIES:
        for p in populations:
            name = f"{family}_bias{p}"
            setter = getattr(cosmo, f"set_{name}")
            if raw.get(f"read_{name}", "0").strip() == "1":
                z, v = np.loadtxt(raw[f"input_{name}"].strip('" '), unpack=True)
                setter(z, v)
            elif f"input_{name}" in raw and name in raw:
                setter(np.linspace(0, 10, 100), [float(raw[name])] * 100)           # a constant bias on the reference's 100 knots
            elif f"{name}_redshifts" in raw and f"{name}_values" in raw:
                setter(convert("float_array", raw[f"{name}_redshifts"]), convert("float_array", raw[f"{name}_values"]))
    return cosmo


# integer codes of the C enums (src/common.h:271-302) behind the options the Python class names
CODES = {"covariance_integration_method": {1: "standard", 2: "fftlog"}, "pk_type": {0: "linear", 1: "linear_class", 2: "halofit", 3: "hmcode"},
         "covariance_interpolation_method": {1: "bilinear", 2: "bicubic"}, "integration_1d_type": {1: "gsl", 2: "double_exponential"}}


def _assign(cosmo, key, value):
    if not hasattr(type(cosmo), key):
        return                                        # an option of the reference this package has no use for (CLASS precision knobs)
    if key in CODES and isinstance(value, int):
        value = CODES[key].get(value, value)
    if isinstance(getattr(cosmo, key), bool):
        value = bool(value)
    setattr(cosmo, key, value)


def write_settings(cosmo, path):
    """every settable parameter as `key = value`, then the eight bias splines as knot lists (the layout `from_file` reads back)"""
    with open(path, "w", encoding="utf-8") as f:
        for key, value in cosmo.parameters.items():
            if isinstance(value, (bool, np.bool_)):
                value = int(value)
            if isinstance(value, (list, tuple, np.ndarray)):
                value = [v.item() if hasattr(v, "item") else v for v in value]
            f.write(f"{key} = {value}\n")
        for family, populations in BIAS_FAMILIES:
            for p in populations:
                z, v = cosmo._bias[f"{family}_bias{p}"]
                f.write(f"{family}_bias{p}_redshifts = {np.asarray(z).tolist()}\n")
                f.write(f"{family}_bias{p}_values = {np.asarray(v).tolist()}\n")
