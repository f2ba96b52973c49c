"""Developer probe: one scene, several trace-kernel selections (pb2_path_params.flags), device-timed renders.

    python tools/probe.py soup <tris> <spp> "<flags...>" [iters]
    python tools/probe.py instanced <object tris> <spp> "<flags...>"
    python tools/probe.py file <scene.pbrt> <spp> "<flags...>"      (1920x1080 forced)
"""
import re
import sys
import time

sys.path.insert(0, ".")
import pbrt_v3_b200 as pb  # noqa: E402

kind, what, spp = sys.argv[1], sys.argv[2], int(sys.argv[3])
flag_list = [int(f) for f in sys.argv[4].split()]
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
t0 = time.time()
if kind == "soup":
    hs = pb.HostScene.soup(int(what), xres=1920, yres=1080, spp=spp)
elif kind == "instanced":
    hs = pb.HostScene.instanced_soup(int(what), grid=10, xres=1920, yres=1080, spp=spp, maxdepth=5)
else:
    text = open(what).read()
    text = re.sub(r'"integer xresolution" \[\d+\]', '"integer xresolution" [1920]', text)
    text = re.sub(r'"integer yresolution" \[\d+\]', '"integer yresolution" [1080]', text)
    text = re.sub(r'"integer pixelsamples" \[\d+\]', '"integer pixelsamples" [%d]' % spp, text)
    hs = pb.HostScene.from_string(text)
hs.device_scene()
print("probe %s %s: scene ready in %.1f s" % (kind, what, time.time() - t0), flush=True)
n = 1920 * 1080 * spp
names = {0: "wide2 ld256", 32: "wide2 ld128", 4: "wide4 ld256", 36: "wide4 ld128", 64: "wide2 ld256 + TMA-staged leaves", 128: "ray pool", 256: "light step chained into the trace kernel", 2: "linear", 8: "plain"}
for flags in flag_list:
    best = None
    for _ in range(iters):
        film, st = hs.render_rgbw(hs.params_copy(flags=flags))
        if best is None or st.render_ms < best.render_ms:
            best = st
    print("probe %s %s %dspp flags=%d (%s): %.1f ms -> %.1f Msamples/s, %.1f Mrays/s, trace %.1f ms (%.0f %%), %d launches"
          % (kind, what, spp, flags, names.get(flags, "?"), best.render_ms, n / best.render_ms / 1e3,
             (best.regular_rays + best.shadow_rays) / best.render_ms / 1e3, best.trace_ms, 100 * best.trace_ms / best.render_ms,
             best.kernel_launches), flush=True)
