"""ctypes mirror of include/ptb.h — the C-ABI of the render core.

Every structure and prototype below restates one declaration of ``include/ptb.h``; tests/test_abi.py checks that the
library exports each symbol and that structure sizes agree with the C compiler's.
"""
import ctypes as C

import numpy as np

from . import lib_path

PTB_ABI_VERSION = 2  # include/ptb.h: the struct layouts mirrored below belong to this version
PTB_OK = 0
PTB_ERR_INVALID_ARGUMENT = 1
PTB_ERR_NO_DEVICE = 2
PTB_ERR_CUDA = 3
PTB_ERR_OUT_OF_MEMORY = 4
PTB_ERR_UNSUPPORTED = 5

PTB_PRIM_TRIANGLE, PTB_PRIM_SPHERE, PTB_PRIM_NULL = 0, 1, 2
PTB_BSDF_LAMBERT, PTB_BSDF_GLASS, PTB_BSDF_MIRROR = 0, 1, 2
PTB_APERTURE_NONE, PTB_APERTURE_CIRCULAR, PTB_APERTURE_HEXAGONAL = 0, 1, 2
PTB_RNG_COUNTER, PTB_RNG_REFERENCE_XORSHIFT = 0, 1
PTB_BVH_REFERENCE = 0
PTB_BVH_REFERENCE_GPU_QUERY_TREE = 1

PTB_FLAG_DEVICE_IO = 0x1
PTB_FLAG_ANY_HIT_SHADOWS = 0x2
PTB_FLAG_SKIP_NULL_SHADOWS = 0x4
PTB_FLAG_COUNT_VISITS = 0x8
PTB_FLAG_CERTIFIED_CLOSEST = 0x10
PTB_FLAG_CERTIFIED_RELAXED = 0x20
PTB_FLAG_PROFILE_ALL = 0x40
PTB_FLAG_SINGLE_STREAM = 0x80

# numpy dtypes of the POD records (layout-identical to the C structs)
PRIM_DTYPE = np.dtype([("kind", "<u4"), ("material", "<u4"), ("cull_backface", "<u4"), ("reserved", "<u4"), ("p", "<f4", (18,))])
MATERIAL_DTYPE = np.dtype(
    [("diffuse", "<f4", (4,)), ("emission", "<f4", (4,)), ("refractive_index", "<f4"), ("bsdf", "<u4"), ("one_way", "<u4"), ("reserved", "<u4")]
)
LIGHT_DTYPE = np.dtype([("pos", "<f4", (3,)), ("rgba", "<f4", (4,))])


class SceneDesc(C.Structure):
    _fields_ = [
        ("prims", C.c_void_p),
        ("n_prims", C.c_uint64),
        ("materials", C.c_void_p),
        ("n_materials", C.c_uint32),
        ("lights", C.c_void_p),
        ("n_lights", C.c_uint32),
        ("bvh_mode", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class SceneInfo(C.Structure):
    _fields_ = [
        ("n_prims", C.c_uint64),
        ("n_inner_nodes", C.c_uint64),
        ("bvh_depth", C.c_uint32),
        ("n_emissive", C.c_uint32),
        ("object_sample_count", C.c_uint32),
        ("n_lights", C.c_uint32),
        ("device_bytes", C.c_uint64),
        ("build_seconds", C.c_double),
        ("upload_seconds", C.c_double),
        ("root_low", C.c_float * 3),
        ("root_high", C.c_float * 3),
        ("query_tree_on_device", C.c_uint32),
        ("certifiable", C.c_uint32),
        ("query_tree_device_ms", C.c_double),
        ("built_on_device", C.c_uint32),
        ("query_tree_kind", C.c_uint32),
        ("reference_tree_device_ms", C.c_double),
    ]


class Camera(C.Structure):
    _fields_ = [
        ("origin", C.c_float * 3),
        ("forward", C.c_float * 3),
        ("up", C.c_float * 3),
        ("right", C.c_float * 3),
        ("aperture_width_half", C.c_float),
        ("aperture_height_half", C.c_float),
        ("aperture_kind", C.c_uint32),
        ("hexagon_horizontal_ratio", C.c_float),
        ("focal_plane_dist", C.c_float),
    ]


class RenderOpts(C.Structure):
    _fields_ = [
        ("image_width", C.c_int32),
        ("image_height", C.c_int32),
        ("min_sample_count", C.c_int32),
        ("max_sample_count", C.c_int32),
        ("epsilon", C.c_float),
        ("max_depth", C.c_int32),
        ("rng_mode", C.c_uint32),
        ("flags", C.c_uint32),
        ("seed", C.c_uint64),
        ("tile_size", C.c_int32),
        ("shard_index", C.c_int32),
        ("shard_count", C.c_int32),
        ("reserved", C.c_uint32),
    ]


class RenderStats(C.Structure):
    _fields_ = [
        ("samples", C.c_uint64),
        ("closest_rays", C.c_uint64),
        ("shadow_rays", C.c_uint64),
        ("shadow_rays_skipped", C.c_uint64),
        ("path_vertices", C.c_uint64),
        ("inner_visits", C.c_uint64),
        ("leaf_visits", C.c_uint64),
        ("bounce_iterations", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("device_ms_total", C.c_double),
        ("device_ms_trace", C.c_double),
        ("device_ms_shade", C.c_double),
        ("device_ms_trace_shadow", C.c_double),
        ("shadow_inner_visits", C.c_uint64),
        ("shadow_leaf_visits", C.c_uint64),
        ("closest_rays_retraced", C.c_uint64),
        ("certified_suspect_hits", C.c_uint64),
        ("samples_used", C.c_uint64),
        ("adaptive_rounds", C.c_uint64),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


# every entry point of include/ptb.h: name -> (restype, argtypes)
_P = C.c_void_p
PROTOTYPES = {
    "ptb_abi_version": (C.c_int, []),
    "ptb_last_error": (C.c_char_p, []),
    "ptb_camera_init": (C.c_int, [C.POINTER(Camera), _P, _P, _P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_float, C.c_float]),
    "ptb_context_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "ptb_context_destroy": (C.c_int, [_P]),
    "ptb_context_device": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "ptb_context_synchronize": (C.c_int, [_P]),
    "ptb_scene_create": (C.c_int, [_P, C.POINTER(SceneDesc), C.POINTER(_P)]),
    "ptb_scene_destroy": (C.c_int, [_P]),
    "ptb_scene_get_info": (C.c_int, [_P, C.POINTER(SceneInfo)]),
    "ptb_scene_read": (C.c_int, [_P, C.c_uint32, _P, C.c_uint64]),
    "ptb_intersect": (C.c_int, [_P, _P, C.c_uint64, _P, _P, C.c_uint32, C.POINTER(RenderStats)]),
    "ptb_occluded": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_uint32, C.POINTER(RenderStats)]),
    "ptb_render": (C.c_int, [_P, C.POINTER(Camera), C.POINTER(RenderOpts), C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, C.POINTER(RenderStats)]),
    "ptb_render_with_progress": (C.c_int, [_P, C.POINTER(Camera), C.POINTER(RenderOpts), C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, C.POINTER(RenderStats), _P, _P]),
    "ptb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "ptb_scene_clone": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "ptb_render_multi": (C.c_int, [C.POINTER(_P), C.c_int32, C.POINTER(Camera), C.POINTER(RenderOpts), C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P,
                                   C.POINTER(RenderStats), _P, _P]),
    "ptb_render_samples": (C.c_int, [_P, C.POINTER(Camera), C.POINTER(RenderOpts), C.c_uint64, _P, _P, _P, C.POINTER(RenderStats)]),
    "ptb_camera_shoot": (C.c_int, [_P, C.POINTER(Camera), C.c_uint64, _P, C.c_float, C.c_float, _P, _P]),
    "ptb_aperture_sample": (C.c_int, [_P, C.c_uint32, C.c_float, C.c_uint64, _P, _P]),
    "ptb_sample_lights": (C.c_int, [_P, _P, C.POINTER(C.c_uint64), C.c_uint32, _P, C.POINTER(C.c_uint32)]),
    "ptb_aabb_intersect": (C.c_int, [_P, _P, _P, C.c_uint64, _P, _P]),
    "ptb_prim_intersect": (C.c_int, [_P, _P, C.c_uint64, _P, _P]),
    "ptb_prim_normal": (C.c_int, [_P, _P, C.c_uint64, _P, _P]),
    "ptb_prim_sample": (C.c_int, [_P, _P, C.c_uint64, _P, _P]),
    "ptb_bsdf_propagate": (C.c_int, [_P, _P, C.c_float, C.c_uint64, _P, _P, _P]),
    "ptb_bsdf_spectrum": (C.c_int, [_P, _P, C.c_uint32, C.c_uint64, _P, _P]),
    "ptb_post_process": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_uint32, C.c_float, C.c_uint32]),
}

_lib = None


def load():
    """Loads lib/libptb.so (raises FileNotFoundError if the extension is not built) and binds all prototypes."""
    global _lib
    if _lib is None:
        lib = C.CDLL(lib_path("libptb.so"), mode=C.RTLD_GLOBAL)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


class PtbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"ptb status {status}: {message}")
        self.status = status


def check(status):
    if status != PTB_OK:
        raise PtbError(status, load().ptb_last_error().decode("utf-8", "replace"))


def _ptr(array):
    return array.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def xorshift_state(seeds):
    """Raw engine state of a fresh RandomEngine(seed): seed ^ (~seed << 32) (reference base.h:26)."""
    s = np.asarray(seeds, dtype=np.uint64)
    return s ^ ((~s) << np.uint64(32))


def camera_init(origin, look_at, up, focal_length, height, aspect_ratio, aperture_width=0.0, aperture_height=0.0, aperture_kind=PTB_APERTURE_NONE,
                hex_ratio=0.0, focal_plane_dist=0.0):
    cam = Camera()
    o, l, u = _f32(origin), _f32(look_at), _f32(up)
    check(load().ptb_camera_init(C.byref(cam), _ptr(o), _ptr(l), _ptr(u), focal_length, height, aspect_ratio, aperture_width, aperture_height,
                                 aperture_kind, hex_ratio, focal_plane_dist))
    return cam


class Context:
    """One device context (ptb_context)."""

    def __init__(self, device=-1):
        self._h = C.c_void_p()
        check(load().ptb_context_create(device, C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def device(self):
        d = C.c_int()
        check(load().ptb_context_device(self._h, C.byref(d)))
        return d.value

    def close(self):
        if self._h:
            load().ptb_context_destroy(self._h)
            self._h = C.c_void_p()

    # ---- unit entries
    def aabb_intersect(self, low, high, rays):
        rays = _f32(rays, (-1, 6))
        out = np.empty(len(rays), np.float32)
        lo, hi = _f32(low), _f32(high)
        check(load().ptb_aabb_intersect(self._h, _ptr(lo), _ptr(hi), len(rays), _ptr(rays), _ptr(out)))
        return out

    def prim_intersect(self, prim, rays):
        rays = _f32(rays, (-1, 6))
        out = np.empty(len(rays), np.float32)
        prim = np.ascontiguousarray(prim, dtype=PRIM_DTYPE).reshape(1)
        check(load().ptb_prim_intersect(self._h, _ptr(prim), len(rays), _ptr(rays), _ptr(out)))
        return out

    def prim_normal(self, prim, positions):
        positions = _f32(positions, (-1, 3))
        out = np.empty_like(positions)
        prim = np.ascontiguousarray(prim, dtype=PRIM_DTYPE).reshape(1)
        check(load().ptb_prim_normal(self._h, _ptr(prim), len(positions), _ptr(positions), _ptr(out)))
        return out

    def prim_sample(self, prim, states):
        states = np.ascontiguousarray(states, dtype=np.uint64).copy()
        out = np.empty((len(states), 5), np.float32)
        prim = np.ascontiguousarray(prim, dtype=PRIM_DTYPE).reshape(1)
        check(load().ptb_prim_sample(self._h, _ptr(prim), len(states), _ptr(states), _ptr(out)))
        return out, states

    def camera_shoot(self, camera, xy, pixel_width, pixel_height, states):
        xy = _f32(xy, (-1, 2))
        states = np.ascontiguousarray(states, dtype=np.uint64).copy()
        out = np.empty((len(xy), 6), np.float32)
        check(load().ptb_camera_shoot(self._h, C.byref(camera), len(xy), _ptr(xy), pixel_width, pixel_height, _ptr(states), _ptr(out)))
        return out, states

    def aperture_sample(self, kind, ratio, states):
        states = np.ascontiguousarray(states, dtype=np.uint64).copy()
        out = np.empty((len(states), 2), np.float32)
        check(load().ptb_aperture_sample(self._h, kind, ratio, len(states), _ptr(states), _ptr(out)))
        return out, states

    def bsdf_propagate(self, material, epsilon, inputs, states):
        inputs = _f32(inputs, (-1, 9))
        states = np.ascontiguousarray(states, dtype=np.uint64).copy()
        out = np.empty((len(inputs), 8), np.float32)
        material = np.ascontiguousarray(material, dtype=MATERIAL_DTYPE).reshape(1)
        check(load().ptb_bsdf_propagate(self._h, _ptr(material), epsilon, len(inputs), _ptr(inputs), _ptr(states), _ptr(out)))
        return out, states

    def post_process(self, image, mode=2, gamma=1.8):
        """toneMap (0) / gammaCorrect (1) / postProcess (2) on a [h, w, 4] float image; returns the processed copy."""
        img = np.ascontiguousarray(image, dtype=np.float32).copy()
        h, w = img.shape[:2]
        check(load().ptb_post_process(self._h, _ptr(img), w, h, mode, gamma, 0))
        return img

    def post_process_device(self, ptr, width, height, mode=2, gamma=1.8):
        check(load().ptb_post_process(self._h, C.c_void_p(ptr), width, height, mode, gamma, PTB_FLAG_DEVICE_IO))

    def bsdf_spectrum(self, material, synthetic, inputs):
        inputs = _f32(inputs, (-1, 13))
        out = np.empty((len(inputs), 6), np.float32)
        material = np.ascontiguousarray(material, dtype=MATERIAL_DTYPE).reshape(1)
        check(load().ptb_bsdf_spectrum(self._h, _ptr(material), 1 if synthetic else 0, len(inputs), _ptr(inputs), _ptr(out)))
        return out


class Scene:
    """Device-resident scene (ptb_scene) built from POD arrays."""

    def __init__(self, ctx, prims, materials, lights=None, bvh_mode=0):
        self.ctx = ctx
        self.prims = np.ascontiguousarray(prims, dtype=PRIM_DTYPE)
        self.materials = np.ascontiguousarray(materials, dtype=MATERIAL_DTYPE)
        self.lights = np.ascontiguousarray(lights if lights is not None else np.zeros(0, LIGHT_DTYPE), dtype=LIGHT_DTYPE)
        desc = SceneDesc()
        desc.prims = self.prims.ctypes.data if len(self.prims) else None
        desc.n_prims = len(self.prims)
        desc.materials = self.materials.ctypes.data if len(self.materials) else None
        desc.n_materials = len(self.materials)
        desc.lights = self.lights.ctypes.data if len(self.lights) else None
        desc.n_lights = len(self.lights)
        desc.bvh_mode = bvh_mode
        self._h = C.c_void_p()
        check(load().ptb_scene_create(ctx.handle, C.byref(desc), C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            load().ptb_scene_destroy(self._h)
            self._h = C.c_void_p()

    def info(self):
        info = SceneInfo()
        check(load().ptb_scene_get_info(self._h, C.byref(info)))
        return info

    NODE_DTYPE = np.dtype([("left_lo", np.float32, 3), ("left_hi", np.float32, 3), ("right_lo", np.float32, 3), ("right_hi", np.float32, 3),
                           ("left", np.int32), ("right", np.int32), ("leaf_count", np.int32), ("parent", np.int32)])

    def read_nodes(self, query_tree=False):
        """The 64-byte inner records of the reference-topology tree (or the query tree) as a structured array."""
        n = max(len(self.prims) - 1, 0)
        if query_tree and self.info().query_tree_kind == 0:
            n = 0
        out = np.zeros(n, self.NODE_DTYPE)
        check(load().ptb_scene_read(self._h, 1 if query_tree else 0, _ptr(out) if n else None, out.nbytes))
        return out

    def read_slot_to_prim(self):
        out = np.zeros(len(self.prims), np.uint32)
        check(load().ptb_scene_read(self._h, 4, _ptr(out) if len(out) else None, out.nbytes))
        return out

    def read_geom(self):
        out = np.zeros((len(self.prims), 16), np.float32)
        check(load().ptb_scene_read(self._h, 2, _ptr(out) if len(out) else None, out.nbytes))
        return out

    def read_shade(self):
        out = np.zeros((len(self.prims), 12), np.float32)
        check(load().ptb_scene_read(self._h, 3, _ptr(out) if len(out) else None, out.nbytes))
        return out

    def intersect(self, rays, flags=0):
        rays = _f32(rays, (-1, 6))
        t = np.empty(len(rays), np.float32)
        prim = np.empty(len(rays), np.int32)
        stats = RenderStats()
        check(load().ptb_intersect(self._h, _ptr(rays), len(rays), _ptr(t), _ptr(prim), flags, C.byref(stats)))
        return t, prim, stats

    def intersect_device(self, rays_ptr, n, t_ptr, prim_ptr, flags=0):
        stats = RenderStats()
        check(load().ptb_intersect(self._h, C.c_void_p(rays_ptr), n, C.c_void_p(t_ptr), C.c_void_p(prim_ptr), flags | PTB_FLAG_DEVICE_IO, C.byref(stats)))
        return stats

    def occluded(self, rays, flags=0):
        rays = _f32(rays, (-1, 7))
        out = np.empty(len(rays), np.uint8)
        stats = RenderStats()
        check(load().ptb_occluded(self._h, _ptr(rays), len(rays), _ptr(out), flags, C.byref(stats)))
        return out, stats

    def occluded_device(self, rays_ptr, n, out_ptr, flags=0):
        stats = RenderStats()
        check(load().ptb_occluded(self._h, C.c_void_p(rays_ptr), n, C.c_void_p(out_ptr), flags | PTB_FLAG_DEVICE_IO, C.byref(stats)))
        return stats

    def sample_lights(self, pos, state, max_out=64):
        pos = _f32(pos)
        st = C.c_uint64(int(state))
        out = np.zeros((max_out, 8), np.float32)
        n = C.c_uint32()
        check(load().ptb_sample_lights(self._h, _ptr(pos), C.byref(st), max_out, _ptr(out), C.byref(n)))
        return out[: min(n.value, max_out)], n.value, st.value

    def render(self, camera, opts, rect=None, out_ptr=None):
        """Renders rect = (x0, y0, w, h) (default: the whole image). Returns (image[h, w, 4] or None, stats)."""
        if rect is None:
            rect = (0, 0, opts.image_width, opts.image_height)
        x0, y0, w, h = rect
        stats = RenderStats()
        if out_ptr is not None:
            opts.flags |= PTB_FLAG_DEVICE_IO
            check(load().ptb_render(self._h, C.byref(camera), C.byref(opts), x0, y0, w, h, C.c_void_p(out_ptr), C.byref(stats)))
            return None, stats
        out = np.zeros((h, w, 4), np.float32)
        check(load().ptb_render(self._h, C.byref(camera), C.byref(opts), x0, y0, w, h, _ptr(out), C.byref(stats)))
        return out, stats

    def render_samples(self, camera, opts, pixels, seeds):
        pixels = np.ascontiguousarray(pixels, dtype=np.int32).reshape(-1, 2)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        out = np.zeros((len(pixels), 4), np.float32)
        stats = RenderStats()
        check(load().ptb_render_samples(self._h, C.byref(camera), C.byref(opts), len(pixels), _ptr(pixels), _ptr(seeds), _ptr(out), C.byref(stats)))
        return out, stats


def render_opts(width, height, min_spp, max_spp, epsilon=1e-3, max_depth=0, rng_mode=PTB_RNG_COUNTER, flags=0, seed=1, tile_size=0, shard_index=0,
                shard_count=1):
    o = RenderOpts()
    o.image_width, o.image_height = width, height
    o.min_sample_count, o.max_sample_count = min_spp, max_spp
    o.epsilon = epsilon
    o.max_depth = max_depth
    o.rng_mode = rng_mode
    o.flags = flags
    o.seed = seed
    o.tile_size = tile_size
    o.shard_index = shard_index
    o.shard_count = shard_count
    return o
