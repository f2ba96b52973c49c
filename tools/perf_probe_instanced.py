"""Developer probe: the synthetic instanced workload at a few spp."""
import os
import sys

sys.path.insert(0, ".")
import pbrt_v3_b200 as pb

spp = int(os.environ.get("PROBE_SPP", "8"))
hs = pb.HostScene.instanced_soup(100000, grid=10, xres=1920, yres=1080, spp=spp, maxdepth=5)
for i in range(3):
    film, st = hs.render_rgbw()
n = 1920 * 1080 * spp
print("instanced probe %s: %.1f ms -> %.1f Msamples/s, %.1f Mrays/s (trace %.1f ms)" % (
    os.environ.get("PROBE_TAG", ""), st.render_ms, n / st.render_ms / 1e3, (st.regular_rays + st.shadow_rays) / st.render_ms / 1e3, st.trace_ms))
