"""TEST INFRASTRUCTURE: pull the reference's own tables out of the oracle (oracle/_ref/libcoffe_ref.so)
and package them as a `coffe_b200.capi.TableSet`, i.e. exactly what a patched reference driver would
hand to coffe_b200_upload_tables.  Lets the parity tests compare kernels against the
reference's integrands on identical inputs."""
import numpy as np

from coffe_b200 import capi

FAMILY = [(0, 0), (0, 2), (0, 4), (1, 1), (1, 3), (2, 0), (2, 2), (3, 1)]
DEFAULT_BIAS_Z = np.arange(0.0, 11.0)          # src/parser.c:42-48
DEFAULT_BIAS = {"galaxy_bias1": 1.0, "galaxy_bias2": 1.0, "galaxy_bias3": 1.0, "galaxy_bias4": 1.0,
                "magnification_bias1": 0.0, "magnification_bias2": 0.0, "evolution_bias1": 0.0, "evolution_bias2": 0.0}


def tableset_from_oracle(ref, bias=None, omega_matter=0.3, h=0.67, want_flat=True):
    """ref: oracle.coffe_oracle.RefCoffe with power spectrum, contributions and grids already set."""
    bg = {name: ref.background_table(name) for name in capi.BG_NAMES}
    b = {}
    for name in capi.BIAS_NAMES:
        if bias and name in bias:
            b[name] = bias[name]
        else:
            b[name] = (DEFAULT_BIAS_Z, np.full_like(DEFAULT_BIAS_Z, DEFAULT_BIAS[name]))
    integrals = []
    for (n, l) in FAMILY:
        x, y = ref.integral_table(n, l)
        integrals.append((n, l, 0, x, y))
    t40 = ref.integral_table(4, 0)
    renorm = renorm0 = None
    if t40 is not None:
        integrals.append((4, 0, 0, t40[0], t40[1]))
        renorm = ref.renorm_table()
        renorm0 = ref.renorm0_table()
    if want_flat:
        for l in (0, 2, 4):
            if (1, l) in FAMILY:
                continue
            t = ref.integral_table(1, l)
            if t is not None:
                integrals.append((1, l, 0, t[0], t[1]))
        for (n, l) in ((1, -1), (3, 1), (5, 3), (7, 5)):
            t = ref.integral_table(n, l, half=True)
            if t is not None:
                integrals.append((n, l, 1, t[0], t[1]))
    return capi.TableSet(bg, b, integrals, omega_matter, h, interp_method=5, renorm=renorm, renorm0=renorm0)
