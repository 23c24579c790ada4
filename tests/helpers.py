"""Shared helpers of the test-suite: golden fixtures, POD camera construction, oracle access."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_SCENES = ["cornell_mesh", "mixed", "advanced"]


def load_golden(kind, name):
    return np.load(os.path.join(GOLDEN_DIR, f"{kind}_{name}.npz"))


def camera_kwargs(arr):
    """Inverse of make_golden.camera_array."""
    a = [float(v) for v in arr]
    return dict(origin=tuple(a[0:3]), look_at=tuple(a[3:6]), up=tuple(a[6:9]), focal_length=a[9], height=a[10], aspect_ratio=a[11],
                aperture_width=a[12], aperture_height=a[13], sampler=int(a[14]), hex_ratio=a[15], focal_plane_dist=a[16])


def pod_camera(kw):
    """ptb_camera from Camera constructor arguments (host arithmetic of ptb_camera_init; needs no GPU)."""
    from cpupathtrace_b200 import capi

    return capi.camera_init(kw["origin"], kw["look_at"], kw["up"], kw["focal_length"], kw["height"], kw["aspect_ratio"], kw.get("aperture_width", 0.0),
                            kw.get("aperture_height", 0.0), kw.get("sampler", 0), kw.get("hex_ratio", 0.0), kw.get("focal_plane_dist", 0.0))


def counter_key(seed, px, py, sample):
    """numpy restatement of counterKey() in cpupathtrace_b200/csrc/rng.cuh (splitmix64 finaliser chain)."""
    m = np.uint64

    def mix(z):
        z = (z ^ (z >> m(30))) * m(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> m(27))) * m(0x94D049BB133111EB)
        return z ^ (z >> m(31))

    with np.errstate(over="ignore"):
        px = np.asarray(px, np.uint64)
        py = np.asarray(py, np.uint64)
        sample = np.asarray(sample, np.uint64)
        k = mix(m(seed) ^ m(0xA0761D6478BD642F))
        k = mix(k ^ ((py << m(32)) | px))
        k = mix(k ^ (sample * m(0xE7037ED1A0B428DB) + m(1)))
    return k
