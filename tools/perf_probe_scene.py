"""Developer probe: a .pbrt test scene at 1920x1080 (resolution / spp / subdivision levels overridden)."""
import os
import re
import sys

sys.path.insert(0, ".")
import pbrt_v3_b200 as pb

path = sys.argv[1]
spp = int(os.environ.get("PROBE_SPP", "16"))
text = open(path).read()
text = re.sub(r'"integer xresolution" \[\d+\]', '"integer xresolution" [1920]', text)
text = re.sub(r'"integer yresolution" \[\d+\]', '"integer yresolution" [1080]', text)
text = re.sub(r'"integer pixelsamples" \[\d+\]', '"integer pixelsamples" [%d]' % spp, text)
if "PROBE_LEVELS" in os.environ:
    text = re.sub(r'"integer (n?levels)" \[\d+\]', r'"integer \1" [%s]' % os.environ["PROBE_LEVELS"], text)
hs = pb.HostScene.from_string(text)
d = hs.desc.contents
for i in range(3):
    film, st = hs.render_rgbw()
n = 1920 * 1080 * spp
print("scene probe %s: %d prims, %d spheres: %.1f ms -> %.1f Msamples/s, %.1f Mrays/s (trace %.1f ms)" % (
    os.environ.get("PROBE_TAG", ""), d.n_prims, d.n_spheres, st.render_ms, n / st.render_ms / 1e3,
    (st.regular_rays + st.shadow_rays) / st.render_ms / 1e3, st.trace_ms))
