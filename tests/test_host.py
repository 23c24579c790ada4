"""CPU tests (-m "not gpu") of the C++ host layer (include/PathTrace/** + cpupathtrace_b200/host/**) through the
API-only harness: everything that is host-side set-up in the reference must give the reference's values."""
import numpy as np
import pytest

from cpupathtrace_b200 import scenes, sharding


def _both(ref, b200, spec):
    a, b = spec.replay(ref), spec.replay(b200)
    try:
        return a.get_triangles(), b.get_triangles(), a.get_object_info(), b.get_object_info()
    finally:
        a.close()
        b.close()


def test_reference_mesh_kats(b200):
    """reference test/scene/mesh_test.cpp:12-31"""
    for text in ("", "  \t  \n\t  \r\n \r", "\n# f 0 1 2\n"):
        s = scenes.SceneSpec()
        s.mesh_obj(text, None, True, True, -1)
        b = s.replay(b200)
        assert b.object_count() == 0
        b.close()
    s = scenes.SceneSpec()
    s.mesh_obj("v 0 0 0\nv 1 0 0\nv 1 0 1\nv 0 0 1\nf 1 2 3\nf 3 4 1", None, True, True, -1)
    b = s.replay(b200)
    assert b.object_count() == 2
    b.close()


def test_planes_boxes_and_obj_meshes_match_reference(ref, b200):
    spec = scenes.cornell_demo(("obj", scenes.standin_obj(48, 36)))
    ta, tb, ia, ib = _both(ref, b200, spec)
    assert ta.shape == tb.shape and np.array_equal(ta, tb, equal_nan=True)  # vertices and smooth normals, bit for bit
    assert np.array_equal(ia, ib)  # getSurfaceArea and getBoundingVolume of every object

    spec = scenes.SceneSpec()
    spec.plane((0, 0, 0), (1, 1, 0))          # valid: flat in z
    spec.plane((0, 0, 0), (1, 0, 0))          # invalid: flat in two axes
    spec.plane((0, 0, 0), (1, 1, 1))          # invalid: not flat
    spec.plane((2, 1, -3), (2, -1, 4), True)  # flat in x, culled
    spec.box((0, 0, 0), (1, 2, 3), True)
    spec.box((0, 0, 0), (1, 0, 3))            # invalid: flat
    ta, tb, ia, ib = _both(ref, b200, spec)
    assert len(ta) == 2 + 2 + 12 and np.array_equal(ta, tb) and np.array_equal(ia, ib)


def test_obj_parser_edge_cases_match_reference(ref, b200):
    """Slash forms, trailing tokens, bad numbers, out-of-range / degenerate faces, CRLF, missing final newline
    (SURVEY.md Appendix D)."""
    texts = [
        "v 0 0 0\nv 1 0 0\n  v 1 0 1 extra\nv 0 0 1\n# c\nvn 0 1 0\nf 1 2 3\nf 3//1 4//2 1//3\nf 1/1 2/2 3/3\nf 1 2 2\nf 5 1 2\nf -1 2 3\n"
        "v 1e40 0 0\nv 2 . 3\nf 1 2 6\n\r\n f 4 3 2 1\nf 1 2",
        "v 0 0 0\r\nv 1 0 0\r\nv 0 1 0\r\nf 1 2 3\r\n",
        "v 1.5e-1 +2 -3\nv 4 5 6\nv 7 8 10\nvt 0 0\nf 1/1/1 2/2/2 3/3/3\nf 1 2 3\ng group\nusemtl x\n",
        "v 0 0 0\nv 1 0 0\nv 2 0 0\nf 1 2 3\nv 0 1 0\nf 1 2 4\nf 4 2 1\n",
        "vv 1 2 3\nv1 2 3\nf\nv\nv 1\nv 1 2\nf 1 1 1\n",
    ]
    transform = (2.0, 0.0, 0.0, 0.5, 0.0, 1.5, 0.0, -1.0, 0.0, 0.0, 0.5, 2.0, 0.0, 0.0, 0.0, 1.0)
    for text in texts:
        for smooth in (True, False):
            for tr in (None, transform):
                spec = scenes.SceneSpec()
                spec.mesh_obj(text, tr, False, smooth, -1)
                ta, tb, _, _ = _both(ref, b200, spec)
                assert ta.shape == tb.shape and np.array_equal(ta, tb, equal_nan=True), (text, smooth, tr)


def test_png_round_trip(b200):
    """reference test/image/image_io_test.cpp:12-40: encode/decode within 0.004 per channel (seeded random 256x128)."""
    rng = np.random.Generator(np.random.PCG64(1234))
    image = rng.uniform(0, 1, size=(128, 256, 4)).astype(np.float32)
    image[0, 0] = (1.5, -0.2, 0.5, 1.0)  # clamped by the writer
    decoded, n_bytes = b200.png_roundtrip(image)
    assert n_bytes > 64
    clipped = np.clip(image, 0, 1)
    assert np.abs(decoded - clipped).max() <= 0.004
    levels = decoded * np.float32(255.0)
    assert np.abs(levels - np.round(levels)).max() < 1e-3  # 8-bit levels
    assert np.abs(decoded - clipped).max() <= 0.5 / 255 + 1e-6  # round(255 v), not truncation


def test_png_quantisation_matches_the_reference_formula(b200):
    """reference src/image/image_io.cpp:139-142 quantises round(255.0 * v) in double; values whose float product would
    round across a .5 boundary (v = 0x1.383838p-1 -> 155, not 156) must come out as the double formula says."""
    levels = np.arange(0, 255, dtype=np.float64) + 0.5
    centre = (levels / 255.0).astype(np.float32)
    near = np.concatenate([np.nextafter(centre, np.float32(0)), centre, np.nextafter(centre, np.float32(1)), np.float32([float.fromhex("0x1.383838p-1")])])
    rng = np.random.Generator(np.random.PCG64(5))
    values = np.concatenate([near, rng.uniform(0, 1, 4 * 64 * 64 - len(near)).astype(np.float32)])
    image = values.reshape(64, 64, 4)
    decoded, n_bytes = b200.png_roundtrip(image)
    assert n_bytes > 0
    want = np.clip(np.round(255.0 * image.astype(np.float64)), 0, 255)
    assert np.array_equal(np.round(decoded.astype(np.float64) * 255.0), want)


def test_png_reader_rejects_hostile_headers_with_logic_error(b200):
    """A tiny file whose IHDR promises 65535 x 65535 pixels must fail with the documented std::logic_error, not allocate
    gigabytes or throw std::bad_alloc; so must truncated streams and 16-bit / interlaced files (documented as unsupported)."""
    import struct
    import zlib

    def chunk(kind, payload):
        return struct.pack(">I", len(payload)) + kind + payload + struct.pack(">I", zlib.crc32(kind + payload) & 0xFFFFFFFF)

    def png(width, height, depth=8, colour=6, interlace=0, data=b"\x00" * 16):
        ihdr = struct.pack(">IIBBBBB", width, height, depth, colour, 0, 0, interlace)
        return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(data)) + chunk(b"IEND", b"")

    assert b200.png_decode_status(png(65535, 65535)) == 1
    assert b200.png_decode_status(png(2, 2, depth=16, data=b"\x00" * 34)) == 1
    assert b200.png_decode_status(png(2, 2, interlace=1, data=b"\x00" * 18)) == 1
    assert b200.png_decode_status(png(2, 2, data=b"\x00" * 18)[:-20]) == 1
    assert b200.png_decode_status(png(2, 2, data=b"\x00" * 18)) == 0
    assert b200.png_decode_status(b"not a png") == 1


def test_tile_sharding_partitions_the_frame():
    for (w, h, world) in [(1920, 1080, 8), (132, 68, 3), (16, 16, 2), (1, 1, 4), (100, 37, 5)]:
        owners = sharding.owner_map(w, h, world)
        assert owners.shape == (h, w) and owners.min() >= 0 and owners.max() < world
        counts = np.bincount(owners.ravel(), minlength=world)
        tiles_x, tiles_y, tile = sharding.tile_grid(w, h)
        if tiles_x * tiles_y >= world:
            assert counts.min() > 0
        total = sum(sharding.owned_pixels(w, h, r, world).sum() for r in range(world))
        assert total == w * h
    assert sharding.reference_tile_size(1920, 1080) == 32 and sharding.reference_tile_size(132, 68) == 17 and sharding.reference_tile_size(3, 9) == 1
