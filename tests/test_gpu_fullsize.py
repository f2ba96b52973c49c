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
