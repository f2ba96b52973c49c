"""Developer script: small renders of every scene class, meant to run under compute-sanitizer (memcheck / initcheck)."""
import os
import sys

sys.path.insert(0, ".")
import numpy as np

import pbrt_v3_b200 as pb

os.environ.setdefault("PB2_POOL", "65536")
makers = [("soup", lambda: pb.HostScene.soup(3000, xres=48, yres=27, spp=2))]
for _n in ("killeroo_like", "materials", "instances", "specular", "substrate", "metal", "uber", "roughglass", "lights"):
    makers.append((_n, lambda _n=_n: pb.HostScene.from_file(os.path.join("tests", "scenes", _n + ".pbrt"))))
makers.append(("instanced_soup", lambda: pb.HostScene.instanced_soup(500, grid=3, xres=48, yres=27, spp=2)))
sys.path.insert(0, "tests")
import golden_cases as gc  # noqa: E402

for _c in ("gaussian", "sinc", "gaussian_aniso_crop"):
    makers.append(("filter_" + _c, lambda _c=_c: pb.HostScene.from_string(gc.filter_scene_text(os.path.join("tests", "scenes"), _c))))
makers.append(("hlbvh", lambda: pb.HostScene.from_string(gc.with_accelerator(open(os.path.join("tests", "scenes", "killeroo_like.pbrt")).read(), "hlbvh", 4))))
for name, make in makers:
    hs = make()   # the host front end keeps ONE parsed scene: build, use, then build the next
    img, st = hs.render()
    film, st1 = hs.render_rgbw(hs.params_copy(flags=1))
    film2, st2 = hs.render_rgbw(hs.params_copy(flags=2))
    rays = np.zeros(256, pb.RAY_DTYPE)
    rays["o"] = (0, -3, 1)
    rays["d"] = np.random.RandomState(1).normal(size=(256, 3)).astype(np.float32)
    rays["t_max"] = np.inf
    hs.intersect(rays)
    hs.intersect_p(rays)
    print(name, "ok", float(img.mean()), st.regular_rays, st1.node_visits)
