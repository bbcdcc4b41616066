import hashlib
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)

import parity as P  # noqa: E402
from forkerrenderer_b200 import binding as B  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: minutes of CPU time (set FGL_SLOW=1)")


def _ensure_oracle():
    if not (os.path.exists(P.ORACLE_LIB) and os.path.exists(P.ORACLE_HOST_LIB)):
        import __graft_entry__ as g
        g.build_oracle()


@pytest.fixture(scope="session")
def golden():
    return json.load(open(os.path.join(HERE, "golden", "reference_hashes.json")))


@pytest.fixture(scope="session")
def oracle_host():
    _ensure_oracle()
    if not os.path.isdir(os.path.join(P.ASSETS, "obj")):
        pytest.skip("reference assets are not staged (oracle/_ref/assets)")
    return B.Host(P.ORACLE_HOST_LIB)


@pytest.fixture()
def oracle_fgl():
    _ensure_oracle()
    f = B.Fgl(P.ORACLE_LIB)
    yield f
    f.close()


@pytest.fixture(scope="session")
def gpu_host():
    return B.product_host()


@pytest.fixture()
def gpu_fgl():
    f = B.product_fgl(0)
    yield f
    f.close()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# planes the north star requires bit-exact (coverage, depth winners, integer buffers) plus every fp32 G-buffer plane,
# which this build also reproduces bit for bit (no transcendental is involved before lighting)
EXACT_PLANES = ["depth", "shadow", "ids_camera", "ids_light", "normal", "worldpos", "lightndc", "albedo", "emissive", "param",
                "shadingtype", "ao"]
# stated tolerances (DESIGN.md "tolerances"): lighting goes through powf, CUDA's differs from glibc's by <= 2 ulp
FRAME_F32_MAX_ABS = 2.5e-7
COLOUR_MAX_LSB = 1
COLOUR_MIN_FRACTION_WITHIN_1LSB = 0.999
