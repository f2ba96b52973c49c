"""Developer probe: how close a 1/N share of the frame (pb2_path_params.tile_count = N, this GPU renders tile_rank 0) comes to
1/N of the full frame's time - the device-side part of the multi-GPU efficiency, measured on one GPU.
    [PB2_POOL=...] python tools/probe_partition.py <tris> <spp> "<tile counts>" [iters]"""
import sys

sys.path.insert(0, ".")
import pbrt_v3_b200 as pb  # noqa: E402

tris, spp = int(sys.argv[1]), int(sys.argv[2])
counts = [int(c) for c in sys.argv[3].split()]
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 4
hs = pb.HostScene.soup(tris, xres=1920, yres=1080, spp=spp)
hs.device_scene()
full = None
for tc in counts:
    best = None
    for _ in range(iters):
        film, st = hs.render_rgbw(hs.params_copy(tile_count=tc, tile_rank=0))
        if best is None or st.render_ms < best.render_ms:
            best = st
    if full is None:
        full = best.render_ms * tc
    print("partition 1/%d: %.2f ms, %d launches; efficiency against the first line %.3f"
          % (tc, best.render_ms, best.kernel_launches, full / (best.render_ms * tc)), flush=True)
