"""CPU tests (-m "not gpu") of the drop-in boundary: the C-ABI library loads, exports every symbol include/ptb.h
declares, its POD layouts agree with the ctypes mirror, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from cpupathtrace_b200 import REPO_ROOT, capi, lib_path

HEADER = os.path.join(REPO_ROOT, "include", "ptb.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ptb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(lib_path("libptb.so"))
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/ptb.h but not exported by libptb.so"
    # and the ctypes mirror binds exactly the declared set
    assert sorted(capi.PROTOTYPES) == names


def test_pod_layouts_match_the_c_compiler(tmp_path):
    source = tmp_path / "sizes.c"
    source.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "ptb.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(ptb_prim), sizeof(ptb_material), sizeof(ptb_point_light),"
        " sizeof(ptb_scene_desc), sizeof(ptb_scene_info), sizeof(ptb_camera), sizeof(ptb_render_opts), sizeof(ptb_render_stats),"
        " offsetof(ptb_render_opts, seed), offsetof(ptb_prim, p));return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c11", f"-I{os.path.join(REPO_ROOT, 'include')}", str(source), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], check=True, stdout=subprocess.PIPE, text=True).stdout.split()]
    assert sizes == [capi.PRIM_DTYPE.itemsize, capi.MATERIAL_DTYPE.itemsize, capi.LIGHT_DTYPE.itemsize, C.sizeof(capi.SceneDesc), C.sizeof(capi.SceneInfo),
                     C.sizeof(capi.Camera), C.sizeof(capi.RenderOpts), C.sizeof(capi.RenderStats), capi.RenderOpts.seed.offset,
                     capi.PRIM_DTYPE.fields["p"][1]]


def test_header_is_plain_c():
    """include/ptb.h must compile as C (no C++-isms): it is what cgo / JNI / ctypes bindings consume."""
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", HEADER], check=True)


def test_host_libraries_load_and_identify():
    from cpupathtrace_b200 import pth

    b200 = pth.load_b200()
    assert b200.name == "b200"
    assert capi.load().ptb_abi_version() == capi.PTB_ABI_VERSION == 2


def test_camera_init_is_host_side_and_matches_reference(ref):
    """ptb_camera_init restates Camera::Camera; the derived frame must give the reference's rays (checked on the
    CPU through the oracle restatement, which consumes the POD camera)."""
    from oracle import pto

    kw = dict(origin=(0.3, -0.2, -2.5), look_at=(0.1, 0.0, 0.2), up=(0.1, 1.0, 0.0), focal_length=1.3, height=0.8, aspect_ratio=-1.7,
              aperture_width=0.07, aperture_height=0.04, sampler=2, hex_ratio=0.35, focal_plane_dist=3.1)
    cam = capi.camera_init(kw["origin"], kw["look_at"], kw["up"], kw["focal_length"], kw["height"], kw["aspect_ratio"], kw["aperture_width"],
                           kw["aperture_height"], 2, kw["hex_ratio"], kw["focal_plane_dist"])
    rng = np.random.Generator(np.random.PCG64(4))
    xy = rng.uniform(-1, 1, size=(2000, 2)).astype(np.float32)
    seeds = rng.integers(1, 2**62, 2000, dtype=np.int64).astype(np.uint64)
    assert np.array_equal(ref.camera(**kw).shoot(xy, 1 / 640, 1 / 360, seeds), pto.camera_shoot(cam, xy, 1 / 640, 1 / 360, seeds))


def test_no_cpu_fallback_without_a_device():
    """Where no CUDA device exists (this container), context creation must fail with PTB_ERR_NO_DEVICE and say so;
    where one exists the test checks that a context can be made."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from cpupathtrace_b200 import capi\n"
            "import ctypes as C\n"
            "lib = capi.load(); h = C.c_void_p()\n"
            "st = lib.ptb_context_create(-1, C.byref(h))\n"
            "print(st, lib.ptb_last_error().decode())\n" % REPO_ROOT)
    out = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, text=True, check=True).stdout.strip()
    status = int(out.split()[0])
    if status != capi.PTB_OK:
        assert status == capi.PTB_ERR_NO_DEVICE
        assert "no CPU fallback" in out


def test_warp_collectives_never_assume_a_full_warp():
    """The wavefront kernels must not hard-code the full member mask in a warp collective: ptxas (12.9, sm_100a) elides
    the barrier in front of such a vote wherever it believes the lanes reconverged, and on B200 they sometimes have
    not -- which once made a trace kernel vote with half a warp and never terminate.  Every
    collective runs over the lanes that are really executing together (__activemask()) or over a mask derived from
    such a vote, and the loops are written so that any group of lanes makes progress on its own."""
    csrc = os.path.join(REPO_ROOT, "cpupathtrace_b200", "csrc")
    calls = 0
    for name in sorted(os.listdir(csrc)):
        if not name.endswith((".cuh", ".cu")):
            continue
        text = open(os.path.join(csrc, name)).read()
        text = re.sub(r"//.*", "", text)
        for m in re.finditer(r"__(ballot|any|all|shfl|shfl_up|shfl_down|shfl_xor|reduce_add|reduce_or|match_any)_sync\s*\(\s*([^,]+),", text):
            calls += 1
            mask = m.group(2).strip()
            assert not re.fullmatch(r"0[xX][fF]{8}[uU]?|~0[uU]?|-1", mask), f"{name}: {m.group(0)} hard-codes the full warp"
    assert calls >= 12
