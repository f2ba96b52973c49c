"""Developer script (runs on the GPU box through gpurun): CUDA path vs the compiled reference.

Prints one line per check; tests/ holds the asserted versions of the same comparisons.
usage: python tools/gpu_check.py [n_tris] [--perf]
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import pbrt_v3_b200 as pb  # noqa: E402
from oracle import pyoracle  # noqa: E402


def random_rays(n, seed=1):
    rng = np.random.RandomState(seed)
    rays = np.zeros(n, pb.RAY_DTYPE)
    o = rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["o"] = o
    rays["d"] = d.astype(np.float32)
    rays["t_max"] = np.inf
    return rays


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def main():
    n_tris = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20000
    ref = pyoracle.best()
    print("oracle kind:", ref.kind)
    hs = pb.HostScene.soup(n_tris, xres=96, yres=54, spp=8)
    rs = ref.scene(hs)

    rays = random_rays(20000)
    t0 = time.time()
    g = hs.intersect(rays)
    print("gpu intersect %.2fs" % (time.time() - t0))
    r = rs.intersect(rays)
    same_prim = g["prim"] == r["prim"]
    hit = r["prim"] >= 0
    print("intersect: prim equal %d/%d, hits %d" % (same_prim.sum(), len(rays), hit.sum()))
    both = same_prim & hit
    for f in ("t", "p", "p_error", "n", "ns", "dpdu", "uv"):
        eq = bits(g[f][both]) == bits(r[f][both])
        print("  %-8s bit-equal %d/%d" % (f, eq.all(axis=-1).sum() if eq.ndim > 1 else eq.sum(), both.sum()))
    # shadow-style rays: finite tMax, unnormalised d
    rays2 = random_rays(20000, 2)
    rays2["d"] *= 0.7
    rays2["t_max"] = 1 - 1e-4
    go, ro = hs.intersect_p(rays2), rs.intersect_p(rays2)
    print("intersect_p: equal %d/%d, occluded %d" % ((go == ro).sum(), len(go), ro.sum()))

    # halton
    rng = np.random.RandomState(3)
    n = 20000
    pix = np.stack([rng.randint(0, 96, n), rng.randint(0, 54, n)], 1).astype(np.int32)
    sn = rng.randint(0, 8, n).astype(np.int64)
    dim = rng.randint(0, 150, n).astype(np.int32)
    gh, rh = hs.halton(pix, sn, dim), ref.halton(hs.film, hs.params, pix, sn, dim)
    print("halton: bit-equal %d/%d" % ((bits(gh) == bits(rh)).sum(), n))

    # light distribution
    pts = rng.uniform(-4, 4, (2000, 3)).astype(np.float32)
    pts[:, 2] = rng.uniform(-1.05, 2.5, 2000)
    gl, rl = hs.light_distribution(pts), rs.light_distribution(pts)
    print("light distribution: bit-equal rows %d/%d, max abs diff %.3g" % ((bits(gl) == bits(rl)).all(axis=1).sum(), len(pts), np.abs(gl - rl).max()))

    # per-sample Li
    n = 20000
    pix = np.stack([rng.randint(0, 96, n), rng.randint(0, 54, n)], 1).astype(np.int32)
    sn = rng.randint(0, 8, n).astype(np.int64)
    gL, gp = hs.li_samples(pix, sn)
    rL, rp = rs.li_samples(pix, sn)
    err = np.abs(gL - rL).max(axis=1) / np.maximum(1, np.abs(rL).max(axis=1))
    print("li: pfilm bit-equal %d/%d; |dL| <= 1e-4 rel: %d/%d; bit-equal %d; max rel err %.3g; mean ref L %.4f mean gpu L %.4f"
          % ((bits(gp) == bits(rp)).all(axis=1).sum(), n, (err <= 1e-4).sum(), n, (bits(gL) == bits(rL)).all(axis=1).sum(), err.max(), rL.mean(), gL.mean()))
    bad = np.where(err > 1e-4)[0][:5]
    for i in bad:
        print("   sample", pix[i], sn[i], "gpu", gL[i], "ref", rL[i])

    # whole image
    img_g, st = hs.render()
    img_r, secs, st_r = rs.render(n_threads=0)
    rel = np.abs(img_g - img_r) / np.maximum(np.abs(img_r), 1e-3)
    print("image: gpu %.1f ms, ref %.2f s; rays gpu %d/%d/%d ref %d/%d/%d" % (st.render_ms, secs, st.camera_rays, st.regular_rays, st.shadow_rays, st_r.camera_rays, st_r.regular_rays, st_r.shadow_rays))
    print("image: mean gpu %.5f ref %.5f; pixels within 1%%: %.4f; mean rel err %.3g; max abs diff %.3g"
          % (img_g.mean(), img_r.mean(), (rel.max(axis=2) <= 0.01).mean(), rel.mean(), np.abs(img_g - img_r).max()))

    if "--perf" in sys.argv:
        for nt, res, spp in ((1000000, (1920, 1080), 4), (1000000, (1920, 1080), 16)):
            t0 = time.time()
            big = pb.HostScene.soup(nt, xres=res[0], yres=res[1], spp=spp)
            t1 = time.time()
            big.device_scene()
            t2 = time.time()
            for it in range(2):
                rgbw, st = big.render_rgbw()
                ns = res[0] * res[1] * spp
                print("perf: %d tris %dx%dx%d: host build %.1fs upload %.1fs render %.1f ms -> %.1f Msamples/s, %.1f Mrays/s (d2h %.1f ms)"
                      % (nt, res[0], res[1], spp, t1 - t0, t2 - t1, st.render_ms, ns / st.render_ms / 1e3, (st.regular_rays + st.shadow_rays) / st.render_ms / 1e3, st.d2h_ms))
            print("   film mean", rgbw[..., :3].mean(), "weight mean", rgbw[..., 3].mean())


if __name__ == "__main__":
    main()
