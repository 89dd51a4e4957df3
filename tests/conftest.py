import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
os.environ.setdefault("CUBACORES", "0")

H = 0.67
ALL = ("den", "rsd", "len", "d1", "d2", "g1", "g2", "g3", "g4", "g5")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_present():
    """True when libcoffe_b200.so loads and coffe_b200_create(0) finds a device.  Only used to SKIP `gpu`-marked tests on a
    GPU-less machine (so `pytest tests` stays readable there); on a GPU box nothing is skipped and a missing library fails loudly."""
    import ctypes as C
    try:
        from coffe_b200 import capi
        lib = capi.load()
    except Exception:
        # no library: on a machine with a GPU this must fail, not skip
        try:
            C.CDLL("libcuda.so.1").cuInit(0)
            return True
        except OSError:
            return False
    h = C.c_void_p()
    rc = lib.coffe_b200_create(C.c_int(0), C.byref(h))
    if rc == 0:
        lib.coffe_b200_destroy(h)
    return rc != 2          # COFFE_B200_ECUDA: no device


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: coffe_b200 has no CPU fallback (gpu-marked tests run on the B200 box)")
    for it in gpu_items:
        it.add_marker(skip)


def _oracle_available():
    from oracle import coffe_oracle as O
    return O.available()


@pytest.fixture(scope="session")
def golden():
    return np.load(ROOT / "tests" / "golden" / "reference_benchmarks.npz")


@pytest.fixture(scope="session")
def ref():
    """The reference's own code (oracle/_ref/libcoffe_ref.so) at the default cosmology with PkL_CLASS.dat;
    tables cover z_mean in [0.5, 2.0] (the renormalisation grid extends to chi(z = 2.3))."""
    from oracle import coffe_oracle as O
    if not O.available():
        pytest.skip("oracle/_ref/libcoffe_ref.so not built (run __graft_entry__.build() where /root/reference exists)")
    r = O.RefCoffe()
    r.set_power_spectrum(O.load_pk())
    r.set_sep(np.array([10.0, 20.0, 40.0, 100.0, 150.0]) / H)
    r.set_z_mean([0.5, 1.0, 1.5, 2.0])
    r.set_deltaz([0.3] * 4)
    r.set_l([0, 2, 4])
    r.set_contributions(ALL)
    return r


@pytest.fixture(scope="session")
def tableset(ref):
    from oracle_tables import tableset_from_oracle
    ref.set_contributions(ALL)
    return tableset_from_oracle(ref)


@pytest.fixture(scope="session")
def engine(tableset):
    import coffe_b200
    eng = coffe_b200.Engine(0)
    eng.upload_tables(tableset)
    return eng
