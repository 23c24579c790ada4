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


def test_chunked_obj_parse_equals_the_sequential_parse(ref, b200, monkeypatch):
    """Files above 1 MB are parsed in chunks on several threads (host/mesh.cpp).  The reference's grammar is sequential: a
    short vertex record eats the first token of the next line, faces are kept only if their indices are below the number
    of vertices read so far.  A file full of such cases -- also right at the chunk cuts, which move with the thread
    count -- must give the triangles of the sequential parse and of the reference's own parser."""
    rng = np.random.Generator(np.random.PCG64(2024))
    lines = []
    n_vertices = 0
    while len(lines) < 70000:
        kind = rng.integers(0, 100)
        if kind < 45:
            lines.append("v %.5f %.5f %.5f" % tuple(rng.uniform(-3, 3, 3)))
            n_vertices += 1
        elif kind < 85:
            hi = max(n_vertices + 40, 3)  # some indices point at vertices that come later (or never): dropped
            a, b, c = (int(v) for v in rng.integers(1, hi, 3))
            lines.append(rng.choice(["f %d %d %d", "f %d//7 %d//8 %d//9", "f   %d %d %d 12"]) % (a, b, c))
        elif kind < 88:
            lines.append("v %.4f %.4f" % tuple(rng.uniform(-3, 3, 2)))  # short record: spills into the next line
            n_vertices += 1
        elif kind < 91:
            lines.append("f %d %d" % tuple(int(v) for v in rng.integers(1, max(n_vertices, 2), 2)))
        elif kind < 94:
            lines.append("# comment v 1 2 3")
        elif kind < 96:
            lines.append("vn 0 0 1")
        elif kind < 98:
            lines.append("")
        else:
            lines.append("  v 1e-2 -2E+0 +3.5 trailing words")
            n_vertices += 1
    text = "\n".join(lines) + "\n"
    crlf = text.replace("\n", "\r\n")
    for data in (text, crlf):
        assert len(data) > (1 << 20)
        spec = scenes.SceneSpec()
        spec.mesh_obj(data, None, False, True, -1)
        results = []
        for threads in ("1", "3", "7", "16"):
            monkeypatch.setenv("PTB_MESH_THREADS", threads)
            b = spec.replay(b200)
            results.append(b.get_triangles())
            b.close()
        b = spec.replay(ref)
        want = b.get_triangles()
        b.close()
        assert len(want) > 1000
        for got in results:
            assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))


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
    gigabytes or throw std::bad_alloc; so must truncated streams and headers with invalid depths or interlace methods."""
    import struct
    import zlib

    def chunk(kind, payload):
        return struct.pack(">I", len(payload)) + kind + payload + struct.pack(">I", zlib.crc32(kind + payload) & 0xFFFFFFFF)

    def png(width, height, depth=8, colour=6, interlace=0, data=b"\x00" * 16):
        ihdr = struct.pack(">IIBBBBB", width, height, depth, colour, 0, 0, interlace)
        return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(data)) + chunk(b"IEND", b"")

    assert b200.png_decode_status(png(65535, 65535)) == 1
    assert b200.png_decode_status(png(2, 2, depth=16, data=b"\x00" * 34)) == 0  # 16-bit: read since round 2
    assert b200.png_decode_status(png(2, 2, depth=3, data=b"\x00" * 18)) == 1  # no such bit depth
    assert b200.png_decode_status(png(2, 2, interlace=2, data=b"\x00" * 18)) == 1  # no such interlace method
    assert b200.png_decode_status(png(2, 2, interlace=1, data=b"\x00" * 18)) == 1  # Adam7 needs 4 pass rows here, not 2
    assert b200.png_decode_status(png(2, 2, data=b"\x00" * 18)[:-20]) == 1
    assert b200.png_decode_status(png(2, 2, data=b"\x00" * 18)) == 0
    assert b200.png_decode_status(b"not a png") == 1


def test_png_reader_reads_what_libpng_reads(b200):
    """The reference reads PNGs through libpng with EXPAND | PACKING | STRIP_16 and copies 3- and 4-channel rows
    (src/image/image_io.cpp:50-79).  io::readRGBImage here: every colour type and bit depth, Adam7, palette with tRNS, colour
    keys, 16-bit samples (high byte); grey and grey + alpha files yield zero pixels, as in the reference.  Expected values
    come from PIL (an independent decoder) for the files PIL can write, and from the source arrays for hand-assembled
    16-bit and interlaced files."""
    import io
    import struct
    import zlib

    from PIL import Image

    rng = np.random.Generator(np.random.PCG64(77))
    w, h = 37, 23  # odd sizes: partial bytes at the end of low-bit rows, ragged Adam7 passes

    def pil_bytes(img, **kw):
        buf = io.BytesIO()
        img.save(buf, format="PNG", **kw)
        return buf.getvalue()

    def expect_rgba(img):
        return np.asarray(img.convert("RGBA"), np.float32) / np.float32(255.0)

    rgb = Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), "RGB")
    rgba = Image.fromarray(rng.integers(0, 256, (h, w, 4), dtype=np.uint8), "RGBA")
    assert np.array_equal(b200.png_decode(pil_bytes(rgb)), expect_rgba(rgb))
    assert np.array_equal(b200.png_decode(pil_bytes(rgba)), expect_rgba(rgba))
    for colours in (2, 4, 16, 200):  # palette images of 1, 2, 4 and 8 bits per index
        pal = Image.fromarray(rng.integers(0, colours, (h, w), dtype=np.uint8), "P")
        pal.putpalette(rng.integers(0, 256, 3 * colours, dtype=np.uint8).tobytes())
        data = pil_bytes(pal, bits={2: 1, 4: 2, 16: 4, 200: 8}[colours])
        assert np.array_equal(b200.png_decode(data), expect_rgba(pal)), colours
        alpha = bytes(rng.integers(0, 256, colours // 2 + 1, dtype=np.uint8))  # tRNS shorter than the palette
        data = pil_bytes(pal, bits={2: 1, 4: 2, 16: 4, 200: 8}[colours], transparency=alpha)
        assert np.array_equal(b200.png_decode(data), expect_rgba(Image.open(io.BytesIO(data)))), colours
    for mode, array in (("L", rng.integers(0, 256, (h, w), dtype=np.uint8)), ("1", rng.integers(0, 2, (h, w), dtype=np.uint8) * 255),
                        ("LA", rng.integers(0, 256, (h, w, 2), dtype=np.uint8))):
        grey = Image.fromarray(array.astype(np.uint8), "L" if mode == "1" else mode).convert(mode)
        got = b200.png_decode(pil_bytes(grey))
        assert got.shape == (h, w, 4) and not got.any(), mode  # the reference copies no rows of 1- and 2-channel images

    def chunk(kind, payload):
        return struct.pack(">I", len(payload)) + kind + payload + struct.pack(">I", zlib.crc32(kind + payload) & 0xFFFFFFFF)

    def assemble(width, height, depth, colour, interlace, raw, extra=b""):
        ihdr = struct.pack(">IIBBBBB", width, height, depth, colour, 0, 0, interlace)
        return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + extra + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")

    def filtered_rows(rows, bpp):
        """Scanlines with PNG filters 0..4 in turn (rows: [n, bytes] uint8)."""
        out = bytearray()
        prev = np.zeros(rows.shape[1], np.int32)
        for y, row in enumerate(rows.astype(np.int32)):
            f = y % 5
            left = np.concatenate([np.zeros(bpp, np.int32), row[:-bpp]]) if rows.shape[1] > bpp else np.zeros_like(row)
            upleft = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]]) if rows.shape[1] > bpp else np.zeros_like(row)
            if f == 0:
                pred = np.zeros_like(row)
            elif f == 1:
                pred = left
            elif f == 2:
                pred = prev
            elif f == 3:
                pred = (left + prev) // 2
            else:
                p = left + prev - upleft
                pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - upleft)
                pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, upleft))
            out.append(f)
            out += bytes(((row - pred) & 0xFF).astype(np.uint8))
            prev = row
        return bytes(out)

    # 16-bit RGB and RGBA: STRIP_16 keeps the high byte; an RGB colour key is compared at 16 bits
    for channels, colour in ((3, 2), (4, 6)):
        src = rng.integers(0, 65536, (h, w, channels), dtype=np.uint16)
        src[3, 5] = src[4, 6] = (0x1234, 0x5678, 0x9ABC, 0xFFFF)[:channels]
        src[5, 7] = (0x1234, 0x5678, 0x9ABD, 0xFFFF)[:channels]  # differs from the key in a LOW byte only
        rows = src.astype(">u2").tobytes()
        rows = np.frombuffer(rows, np.uint8).reshape(h, w * channels * 2)
        want = np.ones((h, w, 4), np.float32)
        want[..., :channels] = (src >> 8).astype(np.float32) / np.float32(255.0)
        assert np.array_equal(b200.png_decode(assemble(w, h, 16, colour, 0, filtered_rows(rows, 2 * channels))), want), colour
        if channels == 3:
            key = chunk(b"tRNS", struct.pack(">HHH", 0x1234, 0x5678, 0x9ABC))
            want[3, 5, 3] = want[4, 6, 3] = 0.0
            assert np.array_equal(b200.png_decode(assemble(w, h, 16, colour, 0, filtered_rows(rows, 2 * channels), extra=key)), want)

    # Adam7: 8-bit RGBA and 4-bit palette, every pass filtered on its own
    passes = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]
    src = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    raw = b""
    for x0, y0, dx, dy in passes:
        sub = src[y0::dy, x0::dx]
        if sub.size:
            raw += filtered_rows(sub.reshape(sub.shape[0], -1), 4)
    assert np.array_equal(b200.png_decode(assemble(w, h, 8, 6, 1, raw)), src.astype(np.float32) / np.float32(255.0))
    index = rng.integers(0, 16, (h, w), dtype=np.uint8)
    palette = rng.integers(0, 256, (16, 3), dtype=np.uint8)
    raw = b""
    for x0, y0, dx, dy in passes:
        sub = index[y0::dy, x0::dx]
        if sub.size:
            padded = np.zeros((sub.shape[0], (sub.shape[1] + 1) // 2 * 2), np.uint8)
            padded[:, :sub.shape[1]] = sub
            raw += filtered_rows((padded[:, 0::2] << 4) | padded[:, 1::2], 1)
    want = np.ones((h, w, 4), np.float32)
    want[..., :3] = palette[index].astype(np.float32) / np.float32(255.0)
    assert np.array_equal(b200.png_decode(assemble(w, h, 4, 3, 1, raw, extra=chunk(b"PLTE", palette.tobytes()))), want)


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
