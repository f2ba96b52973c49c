"""BASELINE.json configs[1] at full size (1 M random triangles, 1920x1080) on the GPU: properties that do not need a
full CPU render — conservation of sample weight, additivity over the multi-GPU tile partition, determinism of the
ray counters, Monte-Carlo consistency between sample counts — plus per-sample parity with the CPU checker on a few
thousand samples of the full-size scene.  Sample counts are kept low (2-4 spp) so the file runs in about a minute."""
import numpy as np
import pytest

import golden_cases as gc

pytestmark = pytest.mark.gpu

XRES, YRES, TRIS = 1920, 1080, 1000000


@pytest.fixture(scope="module")
def big(pb):
    return pb.HostScene.soup(TRIS, xres=XRES, yres=YRES, spp=2, maxdepth=8)


def test_every_sample_is_deposited_once(pb, big):
    film, st = big.render_rgbw()
    n = XRES * YRES * 2
    assert st.camera_rays == n
    w = film[..., 3]
    # box filter, radius 0.5: a sample weighs 1 in its own pixel, and additionally in a neighbour only when it falls
    # exactly on a pixel boundary (film.h:128-131; the first Halton samples of a pixel have u = 0 exactly, so this
    # is common at low sample numbers) -> total weight = n + (boundary samples), every pixel holds >= spp
    assert n <= w.sum() <= n * 1.02
    assert (w >= 2).all() and w.max() <= 2 + 8
    assert np.isfinite(film).all() and (film[..., :3] >= 0).all()
    # ~4.0 Scene::Intersect + IntersectP calls per camera sample on this scene
    assert 3.5 < (st.regular_rays + st.shadow_rays) / n < 4.5 and st.shadow_rays < st.regular_rays


def test_ray_counters_are_deterministic_and_films_agree(pb, big):
    a, sa = big.render_rgbw()
    b, sb = big.render_rgbw()
    assert (sa.camera_rays, sa.regular_rays, sa.shadow_rays) == (sb.camera_rays, sb.regular_rays, sb.shadow_rays)
    assert np.array_equal(a[..., 3], b[..., 3])
    # float atomics reorder the per-pixel sums: equal up to rounding
    assert np.allclose(a, b, rtol=1e-5, atol=1e-5)


def test_tile_partition_is_additive_at_full_size(pb, big):
    full, st = big.render_rgbw()
    parts = [big.render_rgbw(big.params_copy(tile_rank=r, tile_count=2)) for r in range(2)]
    assert sum(int(s.camera_rays) for _, s in parts) == st.camera_rays
    assert np.array_equal(parts[0][0][..., 3] + parts[1][0][..., 3], full[..., 3])
    assert np.allclose(parts[0][0] + parts[1][0], full, rtol=1e-4, atol=1e-4)


def test_more_samples_converge_to_the_same_image(pb, big):
    lo, _ = big.render_rgbw()
    hi, _ = big.render_rgbw(big.params_copy(samples_per_pixel=8))
    mean_lo = (lo[..., :3].sum(axis=(0, 1)) / lo[..., 3].sum())
    mean_hi = (hi[..., :3].sum(axis=(0, 1)) / hi[..., 3].sum())
    assert np.allclose(mean_lo, mean_hi, rtol=0.02)
    # 64x64 block means of the two estimates agree within Monte-Carlo noise
    def blocks(f):
        rgb = f[: YRES // 64 * 64, : XRES // 64 * 64, :3] / f[: YRES // 64 * 64, : XRES // 64 * 64, 3:4]
        return rgb.reshape(YRES // 64, 64, XRES // 64, 64, 3).mean(axis=(1, 3))
    bl, bh = blocks(lo), blocks(hi)
    rel = np.abs(bl - bh) / np.maximum(bh, 0.05)
    assert np.median(rel) < 0.05


def test_per_sample_parity_on_the_full_size_scene(pb, big, port):
    sc = port.scene(big)            # builds the 1 M-triangle BVH on the CPU (a few seconds)
    pix, sn = gc.sample_ids(XRES, YRES, 2, 4000, 31)
    li, pfilm = big.li_samples(pix, sn)
    ref_li, ref_pfilm = sc.li_samples(pix, sn)
    assert np.array_equal(gc.bits(pfilm), gc.bits(ref_pfilm))
    err = np.abs(li - ref_li).max(axis=1) / np.maximum(1, np.abs(ref_li).max(axis=1))
    assert (err <= 1e-4).mean() >= 0.999
    nodes = big.nodes()
    rays = gc.rays_for(pb, nodes, 20000, 32)
    g, r = big.intersect(rays), sc.intersect(rays)
    assert np.array_equal(g["prim"], r["prim"]) and np.array_equal(gc.bits(g["t"]), gc.bits(r["t"]))
    assert np.array_equal(gc.bits(g["p"]), gc.bits(r["p"])) and np.array_equal(gc.bits(g["n"]), gc.bits(r["n"]))


def test_wavefront_trace_records_at_full_size(pb, big, port):
    """The benchmarked kernel on the benchmarked scene: k_wf_trace_w<2> (and the other variants) over 200 000 path rays
    and 200 000 shadow rays of the 1 M-triangle soup leave (found, primitive, t, b0, b1, b2) bit-identical to
    pb2_intersect / pb2_intersect_p, which the test above pins to the CPU checker on this very scene."""
    from test_gpu_parity import check_wavefront_records
    nodes = big.nodes()
    rays, srays = gc.rays_for(pb, nodes, 200000, 33), gc.rays_for(pb, nodes, 200000, 34, shadow=True)
    hits, occ = big.intersect(rays), big.intersect_p(srays)
    assert 0.2 < (hits["prim"] >= 0).mean() and 0.05 < occ.mean() < 0.95
    sc = port.scene(big)
    sub = slice(0, 20000)
    ref = sc.intersect(rays[sub])
    assert np.array_equal(hits["prim"][sub], ref["prim"]) and np.array_equal(gc.bits(hits["t"][sub]), gc.bits(ref["t"]))
    assert np.array_equal(occ[sub], sc.intersect_p(srays[sub]))
    check_wavefront_records(pb, big, rays, srays, hits, occ)


CROPS = [(256, 128), (1408, 208), (832, 496), (320, 880), (1600, 864)]   # x0, y0 of 128 x 72 windows spread over the frame


def test_wavefront_image_parity_on_crops_of_the_full_size_frame(pb, checker):
    """BASELINE.json configs[1] exactly (1 M triangles, 1920x1080, 64 spp, maxdepth 8) through the wavefront renderer,
    compared pixel by pixel with the CPU checker (the compiled reference where present) on five 128 x 72 crop windows
    (PathIntegrator "pixelbounds") spread over the frame; image tolerances of DESIGN.md section 4."""
    hs = pb.HostScene.soup(TRIS, xres=XRES, yres=YRES, spp=64, maxdepth=8)
    sc = checker.scene(hs)
    for x0, y0 in CROPS:
        p = hs.params_copy()
        p.pixel_bounds[0], p.pixel_bounds[1], p.pixel_bounds[2], p.pixel_bounds[3] = x0, y0, x0 + 128, y0 + 72
        rgbw, st = hs.render_rgbw(p)
        assert st.camera_rays == 128 * 72 * 64
        got = hs.resolve(rgbw)[y0:y0 + 72, x0:x0 + 128]
        ref_img, _, ref_st = sc.render(n_threads=0, params=p)
        want = ref_img[y0:y0 + 72, x0:x0 + 128]
        # a sample exactly on a pixel boundary also weighs in the neighbouring pixel (film.h:128-131), window border included
        outside = rgbw[..., 3].sum() - rgbw[y0:y0 + 72, x0:x0 + 128, 3].sum()
        assert 0 <= outside <= 64 and (rgbw[y0:y0 + 72, x0:x0 + 128, 3] >= 64).all()
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-3)
        assert (rel.max(axis=2) <= 0.01).mean() >= 0.999 and rel.mean() <= 1e-4, ((x0, y0), float(rel.mean()))
        assert abs(float(got.mean()) - float(want.mean())) <= 1e-4 * float(want.mean())
        assert abs(int(st.regular_rays) - int(ref_st.regular_rays)) <= ref_st.regular_rays // 1000 + 2
        assert abs(int(st.shadow_rays) - int(ref_st.shadow_rays)) <= ref_st.shadow_rays // 1000 + 2


# ---------------------------------------------------------------- BASELINE.json configs[3]: 10 M instanced triangles
@pytest.fixture(scope="module")
def instanced(pb):
    """One 100 000-triangle object instanced 10 x 10 times, 1920x1080, 1 spp for the tests."""
    return pb.HostScene.instanced_soup(100000, grid=10, xres=XRES, yres=YRES, spp=1, maxdepth=5)


def test_instanced_scene_properties_at_full_size(pb, instanced):
    d = instanced.desc.contents
    assert d.n_instances == 100 and d.n_bvhs == 2 and instanced.bvh_range(1)[3] == 100000
    film, st = instanced.render_rgbw()
    n = XRES * YRES
    assert st.camera_rays == n
    w = film[..., 3]
    assert n <= w.sum() <= n * 1.02 and (w >= 1).all()
    assert np.isfinite(film).all() and (film[..., :3] >= 0).all() and film[..., :3].sum() > 0
    # the two-GPU tile partition adds up to the same film
    parts = [instanced.render_rgbw(instanced.params_copy(tile_rank=r, tile_count=2)) for r in range(2)]
    assert sum(int(s.camera_rays) for _, s in parts) == st.camera_rays
    assert np.array_equal(parts[0][0][..., 3] + parts[1][0][..., 3], w)
    assert np.allclose(parts[0][0] + parts[1][0], film, rtol=1e-4, atol=1e-4)
    assert sum(int(s.regular_rays) for _, s in parts) == st.regular_rays and sum(int(s.shadow_rays) for _, s in parts) == st.shadow_rays


def test_instanced_scene_per_sample_parity_at_full_size(pb, instanced, port):
    sc = port.scene(instanced)
    pix, sn = gc.sample_ids(XRES, YRES, 1, 3000, 41)
    li, pfilm = instanced.li_samples(pix, sn)
    ref_li, ref_pfilm = sc.li_samples(pix, sn)
    assert np.array_equal(gc.bits(pfilm), gc.bits(ref_pfilm))
    err = np.abs(li - ref_li).max(axis=1) / np.maximum(1, np.abs(ref_li).max(axis=1))
    assert (err <= 1e-4).mean() >= 0.999
    rays = gc.rays_for(pb, instanced.nodes(), 20000, 42)
    g, r = instanced.intersect(rays), sc.intersect(rays)
    assert (g["prim"] >= 0).mean() > 0.05
    assert np.array_equal(g["prim"], r["prim"]) and np.array_equal(gc.bits(g["t"]), gc.bits(r["t"]))
    assert np.array_equal(gc.bits(g["p"]), gc.bits(r["p"])) and np.array_equal(gc.bits(g["n"]), gc.bits(r["n"]))
    srays = gc.rays_for(pb, instanced.nodes(), 20000, 43, shadow=True)
    assert np.array_equal(instanced.intersect_p(srays), sc.intersect_p(srays))
