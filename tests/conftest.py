import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # NVRTC results of the test run go to a scratch directory, not to the user's ~/.cache
    import tempfile
    os.environ.setdefault("SDB_CACHE_DIR", os.path.join(tempfile.gettempdir(), "sdeint_b200_jit_tests"))


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def rel_err(a, b):
    """max over (path, time) of ||a-b||_2 / ||b||_2 along the state axis
    (the parity measure of SURVEY.md 8d)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    num = np.linalg.norm(a - b, axis=-1)
    den = np.linalg.norm(b, axis=-1)
    den = np.where(den == 0.0, 1.0, den)
    return float(np.max(num / den))


SYSTEMS = {
    "kp4446": lambda s: __import__("sdeint_b200").ops.sys_kp4446(s),
    "kp4459": lambda s: __import__("sdeint_b200").ops.sys_kp4459(s),
    "kps445": lambda s: __import__("sdeint_b200").ops.sys_kps445(s),
    "r74": lambda s: __import__("sdeint_b200").ops.sys_r74(s),
    "lin2": lambda s: __import__("sdeint_b200").ops.sys_lin2(),
    "lin10": lambda s: __import__("sdeint_b200").ops.sys_lin(10, 10),
    "ou": lambda s: __import__("sdeint_b200").ops.sys_ou(),
}


@pytest.fixture(scope="session")
def reference():
    """The live reference, when /root/reference exists (build container only)."""
    if not os.path.isdir("/root/reference/sdeint"):
        pytest.skip("/root/reference not present (GPU box): golden fixtures cover it")
    sys.path.insert(0, "/root/reference")
    import sdeint
    return sdeint
