"""Developer script: small renders of every scene class, meant to run under compute-sanitizer (memcheck / initcheck)."""
import os
import sys

sys.path.insert(0, ".")
import numpy as np

import pbrt_v3_b200 as pb

os.environ.setdefault("PB2_POOL", "65536")
os.environ.setdefault("PB2_LIGHTDIST_LAZY", "1")   # the on-demand light-distribution path for every scene with several lights
makers = [("soup", lambda: pb.HostScene.soup(3000, xres=48, yres=27, spp=2))]
for _n in ("killeroo_like", "materials", "instances", "specular", "substrate", "metal", "uber", "roughglass", "lights"):
    makers.append((_n, lambda _n=_n: pb.HostScene.from_file(os.path.join("tests", "scenes", _n + ".pbrt"))))
makers.append(("instanced_soup", lambda: pb.HostScene.instanced_soup(500, grid=3, xres=48, yres=27, spp=2)))
sys.path.insert(0, "tests")
import golden_cases as gc  # noqa: E402

for _c in ("gaussian", "sinc", "gaussian_aniso_crop"):
    makers.append(("filter_" + _c, lambda _c=_c: pb.HostScene.from_string(gc.filter_scene_text(os.path.join("tests", "scenes"), _c))))
makers.append(("hlbvh", lambda: pb.HostScene.from_string(gc.with_accelerator(open(os.path.join("tests", "scenes", "killeroo_like.pbrt")).read(), "hlbvh", 4))))
sys.path.insert(0, "tests")
import test_gpu_parity as tgp  # noqa: E402

makers.append(("emissive_mesh_lazy_lightdist", lambda: pb.HostScene.from_string(tgp.emissive_mesh_scene(12, res=(32, 20), spp=2))))
for _n in ("textured", "textured_lens", "sobol", "envlight", "envmap", "bumpmap", "texcombine", "checker"):   # image textures + alpha masks, the SobolSampler, the infinite light
    makers.append((_n, lambda _n=_n: pb.HostScene.from_file(os.path.join("tests", "scenes", _n + ".pbrt"))))
for name, make in makers:
    hs = make()   # the host front end keeps ONE parsed scene: build, use, then build the next
    img, st = hs.render()
    film, st1 = hs.render_rgbw(hs.params_copy(flags=1))
    film2, st2 = hs.render_rgbw(hs.params_copy(flags=2))
    # every trace-kernel selection of round 2 (four-child records, 16-byte loads, small stacks, TMA-staged leaves, ray pool,
    # one thread per ray) through a render and through pb2_trace_wavefront
    for flags in (pb.PB2_FLAG_WIDE4, pb.PB2_FLAG_LD128, pb.PB2_FLAG_SMALL_STACK, pb.PB2_FLAG_SMALL_STACK | pb.PB2_FLAG_WIDE4,
                  pb.PB2_FLAG_LEAF_TMA, pb.PB2_FLAG_POOL, pb.PB2_FLAG_PLAIN_TRACE, pb.PB2_FLAG_CHAIN):
        hs.render_rgbw(hs.params_copy(flags=flags))
        wr = np.zeros(300, pb.RAY_DTYPE)
        wr["o"] = (0, -3, 1)
        wr["d"] = np.random.RandomState(2).normal(size=(300, 3)).astype(np.float32)
        wr["t_max"] = np.inf
        hs.trace_wavefront(wr, any_hit=(np.arange(300) % 3 == 0).astype(np.uint8), flags=flags)
    rays = np.zeros(256, pb.RAY_DTYPE)
    rays["o"] = (0, -3, 1)
    rays["d"] = np.random.RandomState(1).normal(size=(256, 3)).astype(np.float32)
    rays["t_max"] = np.inf
    hs.intersect(rays)
    hs.intersect_p(rays)
    print(name, "ok", float(img.mean()), st.regular_rays, st1.node_visits)
