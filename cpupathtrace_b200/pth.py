"""ctypes wrapper of the API-only harness ``harness/pth.cpp``.

The harness is written against the public PathTrace C++ API only and is compiled twice: against the unmodified
reference (``oracle/_ref/libpth_ref*.so`` — the oracle, used by tests / bench baselines only) and against this
repository's headers and host library (``cpupathtrace_b200/lib/libpth_b200.so`` — the product seen through the
reference's own API).  The same Python calls drive both.
"""
import ctypes as C
import os

import numpy as np

from . import REPO_ROOT, lib_path

_P = C.c_void_p
_PROTOTYPES = {
    "pth_impl_name": (C.c_char_p, []),
    "pth_builder_new": (_P, []),
    "pth_builder_free": (None, [_P]),
    "pth_builder_object_count": (C.c_int, [_P]),
    "pth_add_material": (C.c_int, [_P, _P, C.c_float, _P, C.c_int, C.c_int]),
    "pth_add_triangles": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_int]),
    "pth_add_spheres": (C.c_int, [_P, C.c_int, _P, C.c_int]),
    "pth_add_plane": (C.c_int, [_P, _P, _P, C.c_int, C.c_int]),
    "pth_add_box": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int]),
    "pth_add_mesh_obj": (C.c_int, [_P, C.c_char_p, C.c_long, _P, C.c_int, C.c_int, C.c_int]),
    "pth_add_mesh_file": (C.c_int, [_P, C.c_char_p, _P, C.c_int, C.c_int, C.c_int]),
    "pth_add_point_light": (None, [_P, _P, _P]),
    "pth_builder_get_triangles": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "pth_builder_get_object_info": (None, [_P, C.c_int, C.c_int, _P]),
    "pth_scene_new": (_P, [_P]),
    "pth_scene_free": (None, [_P]),
    "pth_scene_intersect": (None, [_P, C.c_long, _P, _P, _P]),
    "pth_scene_intersect_one": (None, [_P, _P, _P, _P]),
    "pth_scene_sample_lights": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_int, _P]),
    "pth_aabb_intersect": (None, [_P, _P, C.c_long, _P, _P]),
    "pth_camera_new": (_P, [_P, _P, _P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float]),
    "pth_camera_free": (None, [_P]),
    "pth_camera_shoot": (None, [_P, C.c_long, _P, C.c_float, C.c_float, _P, _P]),
    "pth_render_samples": (None, [_P, _P, C.c_int, C.c_int, C.c_float, C.c_long, _P, _P, _P]),
    "pth_process_item": (None, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, _P]),
    "pth_process_job": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pth_post_process": (None, [C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "pth_last_job_timeline": (None, [_P]),
    "pth_object_normal": (None, [_P, C.c_int, C.c_long, _P, _P]),
    "pth_object_sample": (None, [_P, C.c_int, C.c_long, _P, _P, _P]),
    "pth_bsdf_propagate": (None, [_P, C.c_int, C.c_float, C.c_long, _P, _P, _P, _P]),
    "pth_bsdf_spectrum": (None, [_P, C.c_int, C.c_int, C.c_long, _P, _P]),
}
_OPTIONAL = {
    "pth_scene_device_handle": (_P, [_P]),
    "pth_png_roundtrip": (C.c_long, [C.c_int, C.c_int, _P, _P]),
    "pth_png_decode_status": (C.c_int, [C.c_char_p, C.c_long]),
    "pth_set_fast_queries": (None, [C.c_int, C.c_int, C.c_int]),
    "pth_set_sharding": (None, [C.c_int, C.c_int, C.c_uint64]),
    "pth_set_render_control": (None, [C.c_int, C.c_int]),
    "pth_set_devices": (C.c_int, [C.c_int]),
}

REF_PARITY = os.path.join(REPO_ROOT, "oracle", "_ref", "libpth_ref.so")
REF_FAST = os.path.join(REPO_ROOT, "oracle", "_ref", "libpth_ref_fast.so")


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(shape) if shape is not None else a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Pth:
    """One loaded harness library (reference or b200 build)."""

    def __init__(self, path):
        self.path = path
        self.lib = C.CDLL(path)
        for name, (restype, argtypes) in _PROTOTYPES.items():
            fn = getattr(self.lib, name)
            fn.restype, fn.argtypes = restype, argtypes
        for name, (restype, argtypes) in _OPTIONAL.items():
            if hasattr(self.lib, name):
                fn = getattr(self.lib, name)
                fn.restype, fn.argtypes = restype, argtypes
        self.name = self.lib.pth_impl_name().decode()

    def set_fast_queries(self, certified_closest, any_hit_shadows, skip_null_shadows):
        """b200 build only: ptb::RenderControl's result-neutral query options for every later call of this process."""
        self.lib.pth_set_fast_queries(int(certified_closest), int(any_hit_shadows), int(skip_null_shadows))

    def set_devices(self, devices):
        """b200 build only: GPUs of this process processJob renders on (ptb::RenderControl::devices); returns how many it will use."""
        return self.lib.pth_set_devices(int(devices))

    def set_render_control(self, max_depth=-1, relaxed_guard=None):
        """b200 build only: ptb::RenderControl::max_depth / relaxed_guard (None / negative = unchanged)."""
        self.lib.pth_set_render_control(int(max_depth), -1 if relaxed_guard is None else int(bool(relaxed_guard)))

    def set_sharding(self, shard_index, shard_count, fixed_seed=0):
        """b200 build only: the share of processJob's tile grid this process renders, and the job seed (0 = random)."""
        self.lib.pth_set_sharding(int(shard_index), int(shard_count), int(fixed_seed))

    # ---- camera
    def camera(self, origin, look_at, up, focal_length, height, aspect_ratio, aperture_width=0.0, aperture_height=0.0, sampler=0, hex_ratio=0.0,
               focal_plane_dist=0.0):
        o, l, u = _f32(origin), _f32(look_at), _f32(up)
        h = self.lib.pth_camera_new(_ptr(o), _ptr(l), _ptr(u), focal_length, height, aspect_ratio, aperture_width, aperture_height, sampler, hex_ratio,
                                    focal_plane_dist)
        return PthCamera(self, h)

    def aabb_intersect(self, low, high, rays):
        rays = _f32(rays, (-1, 6))
        out = np.empty(len(rays), np.float32)
        lo, hi = _f32(low), _f32(high)
        self.lib.pth_aabb_intersect(_ptr(lo), _ptr(hi), len(rays), _ptr(rays), _ptr(out))
        return out

    def post_process(self, mode, image, gamma=1.8):
        img = np.ascontiguousarray(image, dtype=np.float32).copy()
        h, w = img.shape[:2]
        self.lib.pth_post_process(mode, w, h, gamma, _ptr(img))
        return img

    def png_roundtrip(self, image):
        """b200 build only: encode with io::writeRGBImage, decode with io::readRGBImage; returns (decoded, n_bytes)."""
        img = np.ascontiguousarray(image, dtype=np.float32)
        h, w = img.shape[:2]
        out = np.zeros_like(img)
        n = self.lib.pth_png_roundtrip(w, h, _ptr(img), _ptr(out))
        return out, n

    def png_decode(self, data, max_pixels=1 << 22):
        """b200 build only: io::readRGBImage on raw bytes -> float image [h, w, 4] (raises ValueError on a decode failure)."""
        import ctypes as C

        out = np.zeros(max_pixels * 4, np.float32)
        w, h = C.c_int(0), C.c_int(0)
        self.lib.pth_png_decode.restype = C.c_int
        status = self.lib.pth_png_decode(bytes(data), C.c_long(len(data)), C.byref(w), C.byref(h), _ptr(out), C.c_long(max_pixels))
        if status != 0:
            raise ValueError(f"png decode failed with status {status}")
        return out[:w.value * h.value * 4].reshape(h.value, w.value, 4)

    def png_decode_status(self, data):
        """b200 build only: io::readRGBImage on raw bytes; 0 = decoded, 1 = std::logic_error, 2 = any other exception."""
        return self.lib.pth_png_decode_status(bytes(data), len(data))

    def builder(self):
        return PthBuilder(self)


class PthCamera:
    def __init__(self, pth, handle):
        self.pth, self.h = pth, handle

    def shoot(self, xy, pixel_width, pixel_height, seeds):
        xy = _f32(xy, (-1, 2))
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        out = np.empty((len(xy), 6), np.float32)
        self.pth.lib.pth_camera_shoot(self.h, len(xy), _ptr(xy), pixel_width, pixel_height, _ptr(seeds), _ptr(out))
        return out

    def close(self):
        if self.h:
            self.pth.lib.pth_camera_free(self.h)
            self.h = None


class PthBuilder:
    def __init__(self, pth):
        self.pth = pth
        self.h = pth.lib.pth_builder_new()

    def material(self, diffuse=(1, 1, 1, 1), ior=1.0, emission=(0, 0, 0, 0), bsdf=0, one_way=False):
        d, e = _f32(diffuse), _f32(emission)
        return self.pth.lib.pth_add_material(self.h, _ptr(d), ior, _ptr(e), bsdf, 1 if one_way else 0)

    def triangles(self, verts, normals=None, cull=False, material=-1):
        v = _f32(verts, (-1, 9))
        n = _f32(normals, (-1, 9)) if normals is not None else None
        return self.pth.lib.pth_add_triangles(self.h, len(v), _ptr(v), _ptr(n), 1 if cull else 0, material)

    def spheres(self, spheres, material=-1):
        s = _f32(spheres, (-1, 4))
        return self.pth.lib.pth_add_spheres(self.h, len(s), _ptr(s), material)

    def plane(self, a, b, cull=False, material=-1):
        a, b = _f32(a), _f32(b)
        return self.pth.lib.pth_add_plane(self.h, _ptr(a), _ptr(b), 1 if cull else 0, material)

    def box(self, a, b, cull=False, transform=None, material=-1):
        a, b = _f32(a), _f32(b)
        t = _f32(transform, (16,)) if transform is not None else None
        return self.pth.lib.pth_add_box(self.h, _ptr(a), _ptr(b), 1 if cull else 0, _ptr(t), material)

    def mesh_obj(self, text, transform=None, cull=True, smooth=True, material=-1):
        data = text.encode() if isinstance(text, str) else bytes(text)
        t = _f32(transform, (16,)) if transform is not None else None
        return self.pth.lib.pth_add_mesh_obj(self.h, data, len(data), _ptr(t), 1 if cull else 0, 1 if smooth else 0, material)

    def point_light(self, pos, rgba):
        p, c = _f32(pos), _f32(rgba)
        self.pth.lib.pth_add_point_light(self.h, _ptr(p), _ptr(c))

    def object_count(self):
        return self.pth.lib.pth_builder_object_count(self.h)

    def get_triangles(self, first=0, count=None):
        if count is None:
            count = self.object_count() - first
        out = np.empty((count, 18), np.float32)
        self.pth.lib.pth_builder_get_triangles(self.h, first, count, _ptr(out))
        return out

    def get_object_info(self, first=0, count=None):
        if count is None:
            count = self.object_count() - first
        out = np.empty((count, 7), np.float32)
        self.pth.lib.pth_builder_get_object_info(self.h, first, count, _ptr(out))
        return out

    # ---- one-element virtual methods of an object / a material of this builder (engines are RandomEngine(seed);
    # `next_draw` is the engine's next output after the call)
    def object_normal(self, index, positions):
        positions = _f32(positions, (-1, 3))
        out = np.empty_like(positions)
        self.pth.lib.pth_object_normal(self.h, index, len(positions), _ptr(positions), _ptr(out))
        return out

    def object_sample(self, index, seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        out = np.empty((len(seeds), 5), np.float32)
        next_draw = np.empty(len(seeds), np.uint32)
        self.pth.lib.pth_object_sample(self.h, index, len(seeds), _ptr(seeds), _ptr(out), _ptr(next_draw))
        return out, next_draw

    def bsdf_propagate(self, material, epsilon, inputs, seeds):
        inputs = _f32(inputs, (-1, 9))
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        out = np.empty((len(inputs), 8), np.float32)
        next_draw = np.empty(len(inputs), np.uint32)
        self.pth.lib.pth_bsdf_propagate(self.h, material, epsilon, len(inputs), _ptr(inputs), _ptr(seeds), _ptr(out), _ptr(next_draw))
        return out, next_draw

    def bsdf_spectrum(self, material, synthetic, inputs):
        inputs = _f32(inputs, (-1, 13))
        out = np.empty((len(inputs), 6), np.float32)
        self.pth.lib.pth_bsdf_spectrum(self.h, material, 1 if synthetic else 0, len(inputs), _ptr(inputs), _ptr(out))
        return out

    def scene(self):
        return PthScene(self.pth, self.pth.lib.pth_scene_new(self.h))

    def close(self):
        if self.h:
            self.pth.lib.pth_builder_free(self.h)
            self.h = None


class PthScene:
    def __init__(self, pth, handle):
        self.pth, self.h = pth, handle

    def intersect(self, rays):
        rays = _f32(rays, (-1, 6))
        t = np.empty(len(rays), np.float32)
        ids = np.empty(len(rays), np.int32)
        self.pth.lib.pth_scene_intersect(self.h, len(rays), _ptr(rays), _ptr(t), _ptr(ids))
        return t, ids

    def intersect_one(self, ray):
        ray = _f32(ray, (6,))
        t, i = C.c_float(), C.c_int()
        self.pth.lib.pth_scene_intersect_one(self.h, _ptr(ray), C.byref(t), C.byref(i))
        return t.value, i.value

    def sample_lights(self, pos, normal, seed, max_out=64):
        p, n = _f32(pos), _f32(normal)
        out = np.zeros((max_out, 8), np.float32)
        count = self.pth.lib.pth_scene_sample_lights(self.h, _ptr(p), _ptr(n), int(seed), max_out, _ptr(out))
        return out[: min(count, max_out)], count

    def render_samples(self, camera, width, height, epsilon, pixels, seeds):
        pixels = np.ascontiguousarray(pixels, dtype=np.int32).reshape(-1, 2)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        out = np.zeros((len(pixels), 4), np.float32)
        self.pth.lib.pth_render_samples(self.h, camera.h, width, height, epsilon, len(pixels), _ptr(pixels), _ptr(seeds), _ptr(out))
        return out

    def process_item(self, camera, width, height, min_spp, max_spp, epsilon, rect, seed):
        x0, y0, w, h = rect
        out = np.zeros((h, w, 4), np.float32)
        self.pth.lib.pth_process_item(self.h, camera.h, width, height, min_spp, max_spp, epsilon, x0, y0, w, h, int(seed), _ptr(out))
        return out

    def process_job(self, camera, width, height, min_spp, max_spp, epsilon, workers=0):
        out = np.empty((max(height, 0), max(width, 0), 4), np.float32)
        total, mono = C.c_int(), C.c_int()
        calls = self.pth.lib.pth_process_job(self.h, camera.h, width, height, min_spp, max_spp, epsilon, workers, _ptr(out), C.byref(total), C.byref(mono))
        timeline = np.zeros(3, np.float64)
        self.pth.lib.pth_last_job_timeline(_ptr(timeline))
        return out, {"callbacks": calls, "total_tiles": total.value, "monotonic": bool(mono.value), "seconds": float(timeline[0]),
                     "first_callback_s": float(timeline[1]), "half_callbacks_s": float(timeline[2])}

    def device_handle(self):
        """b200 build only: the ptb_scene* behind the C++ Scene, for direct C-ABI calls on the same scene."""
        return self.pth.lib.pth_scene_device_handle(self.h)

    def close(self):
        if self.h:
            self.pth.lib.pth_scene_free(self.h)
            self.h = None


def load_b200():
    """The harness over THIS repository's host library (requires the built extension; no fallback)."""
    return Pth(lib_path("libpth_b200.so"))


def load_reference(fast=False):
    """The harness over the unmodified reference — the oracle.  For tests and CPU baselines only."""
    path = REF_FAST if fast else REF_PARITY
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is missing: build it with `make -C oracle ref` where /root/reference is mounted")
    return Pth(path)
