"""The C ABI library: loads, exports every symbol include/pb2.h declares, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pb2.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb2_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pb):
    L = pb.lib()
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(L, s), "libpb2.so does not export %s declared in include/pb2.h" % s
    declared = int(re.search(r"#define\s+PB2_ABI_VERSION\s+(\d+)", open(os.path.join(ROOT, "include", "pb2.h")).read()).group(1))
    assert L.pb2_abi_version() == declared == pb.PB2_ABI_VERSION


def test_struct_layouts_match_header(pb):
    """ctypes mirrors of the ABI structs have the sizes the C compiler gives them (checked through the host helpers)."""
    assert C.sizeof(pb.BvhNode) == 32 and C.sizeof(pb.Ray) == 28 and C.sizeof(pb.Hit) == 88
    assert C.sizeof(pb.Material) == 192 and C.sizeof(pb.Instance) == 144 and C.sizeof(pb.Bvh) == 32 and C.sizeof(pb.Light) == 32 and C.sizeof(pb.Mesh) == 48 and C.sizeof(pb.Texture) == 80
    assert C.sizeof(pb.PathParams) == 48 and C.sizeof(pb.FilmDesc) == 56
    hs = pb.HostScene.soup(10, xres=16, yres=16, spp=1)
    d = hs.desc.contents
    assert d.n_prims == 14 and d.n_tris == 14 and d.n_lights == 2 and d.n_nodes >= 7
    assert hs.params.contents.samples_per_pixel == 1 and hs.params.contents.max_depth == 8
    assert tuple(hs.film.contents.cropped_pixel_bounds) == (0, 0, 16, 16)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_no_device_fails_loudly(pb):
    """No CPU fallback: without a CUDA device every compute entry point reports PB2_ERR_NO_DEVICE."""
    L = pb.lib()
    assert L.pb2_init(0) == pb.PB2_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.pb2_last_error()
    hs = pb.HostScene.soup(10, xres=16, yres=16, spp=1)
    handle = C.c_void_p()
    assert L.pb2_scene_create(hs.desc, C.byref(handle)) == pb.PB2_ERR_NO_DEVICE
    rays = np.zeros(1, pb.RAY_DTYPE)
    hits = np.zeros(1, pb.HIT_DTYPE)
    assert L.pb2_intersect(None, pb.ptr(rays), 1, pb.ptr(hits)) == pb.PB2_ERR_NO_DEVICE
    film = np.zeros((16, 16, 4), np.float32)
    assert L.pb2_render_path(None, hs.camera, hs.film, hs.params, pb.ptr(film), None) == pb.PB2_ERR_NO_DEVICE
    with pytest.raises(pb.Pb2Error):
        hs.render()


def test_product_does_not_import_the_oracle():
    """The checker must never sit on the product path."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pbrt_v3_b200")):
        if os.sep + "lib" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in text and "libpb2_oracle" not in text and "libpbrt_ref" not in text and "oracle/" not in text.replace("# oracle/", ""), \
                    "%s references the oracle" % os.path.join(dirpath, f)
