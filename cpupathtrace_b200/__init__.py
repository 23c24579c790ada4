"""cpupathtrace_b200 — B200-native render core behind the CPUPathTrace API.

The product is native code:

* ``lib/libptb.so``        the C-ABI render core (``include/ptb.h``): hand-written sm_100a CUDA kernels
  (BVH traversal, wavefront path tracing) plus their host driver;
* ``lib/libPathTrace.so``  the C++ host layer re-providing the reference's public API (``include/PathTrace/**``).

This Python package is only a thin ctypes loader used by the tests, ``bench.py`` and ``__graft_entry__.py``:

* :mod:`cpupathtrace_b200.capi`   ctypes mirror of ``include/ptb.h`` (what a cgo/JNI/N-API binding would bind);
* :mod:`cpupathtrace_b200.pth`    ctypes wrapper of the API-only harness ``harness/pth.cpp`` (the reference's callers),
  loadable against either implementation;
* :mod:`cpupathtrace_b200.scenes` builders for the scenes of BASELINE.json (Cornell demo scene, stand-in mesh, soup).

Nothing here computes on the CPU: every compute entry fails loudly when no CUDA device / extension is present.
"""
import os

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PACKAGE_DIR)
LIB_DIR = os.path.join(PACKAGE_DIR, "lib")


def lib_path(name: str) -> str:
    """Absolute path of a built product library; raises if it has not been built (run __graft_entry__.build())."""
    path = os.path.join(LIB_DIR, name)
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} is missing: the CUDA extension has not been built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C cpupathtrace_b200/csrc`). There is no Python/CPU fallback."
        )
    return path
