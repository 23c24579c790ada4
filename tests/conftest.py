import os
import sys

import numpy as np
import pytest

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO_ROOT not in sys.path:
    sys.path.insert(0, REPO_ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref():
    """The oracle: the unmodified reference behind the API-only harness (oracle/_ref, built by `make -C oracle ref`)."""
    from cpupathtrace_b200 import pth

    if not os.path.exists(pth.REF_PARITY):
        pytest.skip("oracle/_ref/libpth_ref.so not built (needs /root/reference once; the binary travels to the GPU box)")
    return pth.load_reference()


@pytest.fixture(scope="session")
def b200():
    """The product seen through the reference's own C++ API (libpth_b200.so -> libPathTrace.so -> libptb.so)."""
    from cpupathtrace_b200 import pth

    return pth.load_b200()


@pytest.fixture(scope="session")
def ctx():
    """A C-ABI device context; fails (not skips) when the extension or the device is missing."""
    from cpupathtrace_b200 import capi

    context = capi.Context(-1)
    yield context
    context.close()


def ulp_distance(a, b):
    """Distance in units in the last place between two float32 arrays (sign-magnitude ordering)."""
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


def random_rays(n, seed, box=1.2, aim=None):
    """Uniform origins in [-box, box]^3 with uniform directions; with `aim` (points, k) the first k rays are aimed at
    the given points (mesh vertices / edge midpoints: the tie-break stress cases)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    o = rng.uniform(-box, box, size=(n, 3))
    d = rng.normal(size=(n, 3))
    if aim is not None:
        pts = np.asarray(aim, np.float64)
        k = min(len(pts), n)
        d[:k] = pts[:k] - o[:k]
    d32 = d.astype(np.float32)
    # normalise exactly like rt_vector::normalize: multiply by the reciprocal of the fp32 length
    l2 = (d32[:, 0] * d32[:, 0] + d32[:, 1] * d32[:, 1]) + d32[:, 2] * d32[:, 2]
    inv = (np.float32(1.0) / np.sqrt(l2)).astype(np.float32)
    d32 = d32 * inv[:, None]
    return np.concatenate([o.astype(np.float32), d32], axis=1)
