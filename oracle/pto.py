"""ctypes wrapper of oracle/libptoracle.so (the plain-C restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libptoracle.so")
_P = C.c_void_p
_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB)
        lib.pto_scene_create.restype = _P
        lib.pto_scene_create.argtypes = [_P]
        lib.pto_scene_destroy.argtypes = [_P]
        lib.pto_scene_depth.argtypes = [_P]
        lib.pto_scene_emissive_count.argtypes = [_P]
        lib.pto_intersect.argtypes = [_P, _P, C.c_uint64, _P, _P]
        lib.pto_aabb_intersect.argtypes = [_P, _P, C.c_uint64, _P, _P]
        lib.pto_intersect_certified.argtypes = [_P, _P, C.c_uint64, C.c_uint64, C.c_int, _P, _P, _P]
        lib.pto_sample_lights.argtypes = [_P, _P, C.c_uint64, C.c_int, _P]
        lib.pto_camera_shoot.argtypes = [_P, C.c_uint64, _P, C.c_float, C.c_float, _P, _P]
        lib.pto_render_samples.argtypes = [_P, _P, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint64, _P, _P, _P, _P]
        lib.pto_process_item.argtypes = [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, _P]
        lib.pto_resolve.argtypes = [C.c_int, C.c_int, C.c_uint32, _P, _P]
        lib.pto_resolve_counts.argtypes = [C.c_int, C.c_int, C.c_uint32, _P, _P, _P]
        lib.pto_set_guard.argtypes = [C.c_int]
        lib.pto_scene_certifiable.argtypes = [_P]
        lib.pto_prim_normal.argtypes = [_P, C.c_uint64, _P, _P]
        lib.pto_prim_sample.argtypes = [_P, C.c_uint64, _P, _P]
        lib.pto_bsdf_propagate.argtypes = [_P, C.c_float, C.c_uint64, _P, _P, _P]
        lib.pto_bsdf_spectrum.argtypes = [_P, C.c_int, C.c_uint64, _P, _P]
        _lib = lib
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleScene:
    def __init__(self, prims, materials, lights):
        from cpupathtrace_b200 import capi

        self.prims = np.ascontiguousarray(prims, dtype=capi.PRIM_DTYPE)
        self.materials = np.ascontiguousarray(materials, dtype=capi.MATERIAL_DTYPE)
        self.lights = np.ascontiguousarray(lights, dtype=capi.LIGHT_DTYPE)
        desc = capi.SceneDesc()
        desc.prims = self.prims.ctypes.data if len(self.prims) else None
        desc.n_prims = len(self.prims)
        desc.materials = self.materials.ctypes.data if len(self.materials) else None
        desc.n_materials = len(self.materials)
        desc.lights = self.lights.ctypes.data if len(self.lights) else None
        desc.n_lights = len(self.lights)
        self.h = load().pto_scene_create(C.byref(desc))

    def close(self):
        if self.h:
            load().pto_scene_destroy(self.h)
            self.h = None

    def depth(self):
        return load().pto_scene_depth(self.h)

    def intersect(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        t = np.empty(len(rays), np.float32)
        prim = np.empty(len(rays), np.int32)
        load().pto_intersect(self.h, _ptr(rays), len(rays), _ptr(t), _ptr(prim))
        return t, prim

    def certifiable(self):
        """Whether the certified walk's guard table covers this scene (ptb_scene_info.certifiable)."""
        return bool(load().pto_scene_certifiable(self.h))

    def intersect_certified(self, rays, tree_seed, shape=0, guarded=False):
        """Certified closest hit (CPU statement of csrc/traverse.cuh) on a random hierarchy: (t, prim, certain).
        guarded: apply the guard table of csrc/cert_guard.h (flagged rays and uncertifiable scenes come back uncertain)."""
        load().pto_set_guard(1 if guarded else 0)
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        t = np.empty(len(rays), np.float32)
        prim = np.empty(len(rays), np.int32)
        certain = np.empty(len(rays), np.uint8)
        load().pto_intersect_certified(self.h, _ptr(rays), len(rays), int(tree_seed), int(shape), _ptr(t), _ptr(prim), _ptr(certain))
        return t, prim, certain.astype(bool)

    def sample_lights(self, pos, seed, max_out=64):
        pos = np.ascontiguousarray(pos, np.float32)
        out = np.zeros((max_out, 8), np.float32)
        n = load().pto_sample_lights(self.h, _ptr(pos), int(seed), max_out, _ptr(out))
        return out[: min(n, max_out)], n

    def render_samples(self, camera, width, height, epsilon, pixels, seeds, max_depth=0):
        pixels = np.ascontiguousarray(pixels, np.int32).reshape(-1, 2)
        seeds = np.ascontiguousarray(seeds, np.uint64)
        out = np.zeros((len(pixels), 4), np.float32)
        counters = np.zeros(4, np.uint64)
        load().pto_render_samples(self.h, C.byref(camera), width, height, epsilon, max_depth, len(pixels), _ptr(pixels), _ptr(seeds), _ptr(out), _ptr(counters))
        return out, dict(zip(["samples", "closest_rays", "shadow_rays", "vertices"], [int(c) for c in counters]))

    def process_item(self, camera, width, height, min_spp, max_spp, epsilon, rect, seed):
        x0, y0, w, h = rect
        out = np.zeros((h, w, 4), np.float32)
        load().pto_process_item(self.h, C.byref(camera), width, height, min_spp, max_spp, epsilon, x0, y0, w, h, int(seed), _ptr(out))
        return out


def aabb_intersect(low, high, rays):
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    lo, hi = np.ascontiguousarray(low, np.float32), np.ascontiguousarray(high, np.float32)
    out = np.empty(len(rays), np.float32)
    load().pto_aabb_intersect(_ptr(lo), _ptr(hi), len(rays), _ptr(rays), _ptr(out))
    return out


def camera_shoot(camera, xy, pw, ph, seeds):
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    seeds = np.ascontiguousarray(seeds, np.uint64)
    out = np.empty((len(xy), 6), np.float32)
    load().pto_camera_shoot(C.byref(camera), len(xy), _ptr(xy), pw, ph, _ptr(seeds), _ptr(out))
    return out


def resolve(min_spp, max_spp, samples):
    """samples: [spp, n_pixels, 4] -> [n_pixels, 4]"""
    samples = np.ascontiguousarray(samples, np.float32)
    n_pixels = samples.shape[1]
    out = np.zeros((n_pixels, 4), np.float32)
    load().pto_resolve(min_spp, max_spp, n_pixels, _ptr(samples), _ptr(out))
    return out


def resolve_counts(min_spp, max_spp, samples):
    """resolve() plus the number of samples each pixel's loop drew before it ended (worker.cpp:236-260)."""
    samples = np.ascontiguousarray(samples, np.float32)
    n_pixels = samples.shape[1]
    out = np.zeros((n_pixels, 4), np.float32)
    consumed = np.zeros(n_pixels, np.int32)
    load().pto_resolve_counts(min_spp, max_spp, n_pixels, _ptr(samples), _ptr(out), _ptr(consumed))
    return out, consumed


def _pod(value, dtype):
    return np.ascontiguousarray(value, dtype=dtype).reshape(1)


def prim_normal(prim, positions):
    """Object::getSurfaceNormal of one primitive (object.cpp:86-88, 126-144)."""
    from cpupathtrace_b200 import capi

    prim = _pod(prim, capi.PRIM_DTYPE)
    positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
    out = np.empty_like(positions)
    load().pto_prim_normal(_ptr(prim), len(positions), _ptr(positions), _ptr(out))
    return out


def prim_sample(prim, states):
    """Object::sampleSurface (object.cpp:101-116, 192-207) from raw engine states: ([n, 5], advanced states)."""
    from cpupathtrace_b200 import capi

    prim = _pod(prim, capi.PRIM_DTYPE)
    states = np.ascontiguousarray(states, np.uint64).copy()
    out = np.empty((len(states), 5), np.float32)
    load().pto_prim_sample(_ptr(prim), len(states), _ptr(states), _ptr(out))
    return out, states


def bsdf_propagate(material, epsilon, inputs, states):
    """BSDF::propagateRay (propagation.cpp:89-99, 120-160, 180-204): ([n, 8], advanced states)."""
    from cpupathtrace_b200 import capi

    material = _pod(material, capi.MATERIAL_DTYPE)
    inputs = np.ascontiguousarray(inputs, np.float32).reshape(-1, 9)
    states = np.ascontiguousarray(states, np.uint64).copy()
    out = np.empty((len(inputs), 8), np.float32)
    load().pto_bsdf_propagate(_ptr(material), epsilon, len(inputs), _ptr(inputs), _ptr(states), _ptr(out))
    return out, states


def bsdf_spectrum(material, synthetic, inputs):
    """BSDF::getSpectrum (propagation.cpp:101-116, 162-176, 206-217): [n, 6]."""
    from cpupathtrace_b200 import capi

    material = _pod(material, capi.MATERIAL_DTYPE)
    inputs = np.ascontiguousarray(inputs, np.float32).reshape(-1, 13)
    out = np.empty((len(inputs), 6), np.float32)
    load().pto_bsdf_spectrum(_ptr(material), 1 if synthetic else 0, len(inputs), _ptr(inputs), _ptr(out))
    return out
