import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*_site*.npz")))   # one RenormaliseFrom record each


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_present():
    """True when the product library can bind CUDA device 0 (a plain `pytest tests/` on a CPU box must skip, not error)."""
    try:
        from block_b200 import _lib
        import ctypes
        lib = _lib.load()
        ctx = ctypes.c_void_p()
        if lib.b2d_create(0, ctypes.byref(ctx)) != 0:
            return False
        lib.b2d_destroy(ctx)
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: GPU parity tests run on the B200 box (pytest -m gpu)")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session", params=GOLDEN, ids=[os.path.basename(g)[:-4] for g in GOLDEN])
def golden(request):
    """One RenormaliseFrom call dumped from the real reference (tests/golden/make_golden.py)."""
    from oracle import dumpio
    rec = dumpio.read_records(request.param)
    return rec, dumpio.big_from(rec)
