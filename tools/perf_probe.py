"""Developer script: time pb2_render_path_device-style renders of the config-2 scene (1 M triangle soup)."""
import os
import sys
import time

sys.path.insert(0, ".")
import pbrt_v3_b200 as pb  # noqa: E402

nt = int(os.environ.get("PROBE_TRIS", "1000000"))
spp = int(os.environ.get("PROBE_SPP", "4"))
res = (int(os.environ.get("PROBE_XRES", "1920")), int(os.environ.get("PROBE_YRES", "1080")))
big = pb.HostScene.soup(nt, xres=res[0], yres=res[1], spp=spp)
big.device_scene()
best = None
for it in range(int(os.environ.get("PROBE_ITERS", "3"))):
    rgbw, st = big.render_rgbw()
    best = st.render_ms if best is None else min(best, st.render_ms)
import hashlib  # noqa: E402
import numpy as np  # noqa: E402

digest = hashlib.sha1(np.ascontiguousarray(rgbw).tobytes()).hexdigest()[:12]
ns = res[0] * res[1] * spp
print("probe %s [film %s]: %d tris %dx%dx%d: %.1f ms -> %.1f Msamples/s, %.1f Mrays/s; nodes %d prims %d"
      % (os.environ.get("PROBE_TAG", ""), digest, nt, res[0], res[1], spp, best, ns / best / 1e3,
         (st.regular_rays + st.shadow_rays) / best / 1e3, st.node_visits, st.prim_tests))
