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


@pytest.fixture(scope="session", params=GOLDEN, ids=[os.path.basename(g)[:-4] for g in GOLDEN])
def golden(request):
    """One RenormaliseFrom call dumped from the real reference (tests/golden/make_golden.py)."""
    from oracle import dumpio
    rec = dumpio.read_records(request.param)
    return rec, dumpio.big_from(rec)
