"""Prototype for the next step of the trace kernel: collapse the reference's binary BVH into records of up to
four children and traverse it so that the leaves are visited in exactly the order - and with exactly the verdicts - of
BVHAccel::Intersect (bvh.cpp:662-700).

Rule: a record holds the grandchildren [LL, LR, RL, RR] of a node (a child that is a leaf stays as it is), with the split
axes of the node and of its two children.  A visit tests all child boxes with the current tMax, orders the survivors
near-first level by level from dirIsNeg (the reference's order), continues with the first and pushes the others with their
entry distance; a popped entry is re-checked against the tMax of that moment.  Because a child's slab interval lies inside
its parent's and tMax only shrinks, a leaf is reached iff its own box passes at the moment the binary traversal would test
it, so both traversals test the same leaves in the same order.

This script checks that claim on the host-built tree of a random mesh for random rays.  A "primitive test" is stood in for by
the primitive's own bounding box (hit distance = entry distance), which shrinks tMax like a real hit would; the argument
does not depend on what shrinks it.    python tools/wide4_prototype.py [n_tris] [n_rays]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import golden_cases as gc  # noqa: E402
import pbrt_v3_b200 as pb  # noqa: E402

f32 = np.float32
GAMMA3 = f32(3 * 2.0 ** -24 / (1 - 3 * 2.0 ** -24))


def slab(bmin, bmax, o, inv, neg, tmax):
    """Bounds3::IntersectP(ray, invDir, dirIsNeg) (geometry.h:1412-1438): returns (hit, entry distance)."""
    with np.errstate(all="ignore"):
        lo = [bmax[k] if neg[k] else bmin[k] for k in range(3)]
        hi = [bmin[k] if neg[k] else bmax[k] for k in range(3)]
        t0 = f32((lo[0] - o[0]) * inv[0])
        t1 = f32(f32((hi[0] - o[0]) * inv[0]) * (1 + 2 * GAMMA3))
        ty0 = f32((lo[1] - o[1]) * inv[1])
        ty1 = f32(f32((hi[1] - o[1]) * inv[1]) * (1 + 2 * GAMMA3))
        if t0 > ty1 or ty0 > t1:
            return False, t0
        if ty0 > t0:
            t0 = ty0
        if ty1 < t1:
            t1 = ty1
        tz0 = f32((lo[2] - o[2]) * inv[2])
        tz1 = f32(f32((hi[2] - o[2]) * inv[2]) * (1 + 2 * GAMMA3))
        if t0 > tz1 or tz0 > t1:
            return False, t0
        if tz0 > t0:
            t0 = tz0
        if tz1 < t1:
            t1 = tz1
        return bool(t0 < tmax and t1 > 0), t0


def leaf_tests(nodes, prims, pbounds, i, o, inv, neg, tmax, log):
    """stand-in for the primitive loop of a leaf: every primitive is 'tested', a box hit shrinks tMax"""
    for j in range(nodes["offset"][i], nodes["offset"][i] + nodes["n_prims"][i]):
        p = prims[j]
        log.append(int(p))
        hit, t = slab(pbounds[p, 0], pbounds[p, 1], o, inv, neg, tmax)
        if hit and t > 0:
            tmax = t
    return tmax


def binary(nodes, prims, pbounds, o, d):
    with np.errstate(divide="ignore"):
        inv = f32(1) / d
    neg = inv < 0
    tmax, log, stack, cur = f32(np.inf), [], [], 0
    while True:
        hit, _ = slab(nodes["bmin"][cur], nodes["bmax"][cur], o, inv, neg, tmax)
        if hit:
            if nodes["n_prims"][cur] > 0:
                tmax = leaf_tests(nodes, prims, pbounds, cur, o, inv, neg, tmax, log)
                if not stack:
                    break
                cur = stack.pop()
            elif neg[nodes["axis"][cur]]:
                stack.append(cur + 1)
                cur = int(nodes["offset"][cur])
            else:
                stack.append(int(nodes["offset"][cur]))
                cur = cur + 1
        else:
            if not stack:
                break
            cur = stack.pop()
    return log, tmax


def children4(nodes, i):
    """[(node, level-1 side, level-2 side or None)] of interior node i, in the canonical [LL, LR, RL, RR] order"""
    out = []
    for side, c in enumerate((i + 1, int(nodes["offset"][i]))):
        if nodes["n_prims"][c] > 0:
            out.append((c, side, None))
        else:
            out.append((c + 1, side, 0))
            out.append((int(nodes["offset"][c]), side, 1))
    return out


def wide4(nodes, prims, pbounds, o, d):
    with np.errstate(divide="ignore"):
        inv = f32(1) / d
    neg = inv < 0
    tmax, log, stack = f32(np.inf), [], []
    hit, _ = slab(nodes["bmin"][0], nodes["bmax"][0], o, inv, neg, tmax)
    if not hit:
        return log, tmax
    cur = 0
    while True:
        if nodes["n_prims"][cur] > 0:
            tmax = leaf_tests(nodes, prims, pbounds, cur, o, inv, neg, tmax, log)
            cur = None
        else:
            kids = []
            for c, s1, s2 in children4(nodes, cur):
                h, t = slab(nodes["bmin"][c], nodes["bmax"][c], o, inv, neg, tmax)
                if h:
                    parent = cur + 1 if s1 == 0 else int(nodes["offset"][cur])
                    k1 = s1 if not neg[nodes["axis"][cur]] else 1 - s1                       # 0 = visited first at level 1
                    k2 = 0 if s2 is None else (s2 if not neg[nodes["axis"][parent]] else 1 - s2)
                    kids.append((k1, k2, c, t))
            kids.sort()
            cur = kids[0][2] if kids else None
            for k1, k2, c, t in reversed(kids[1:]):
                stack.append((c, t))
        while cur is None:
            if not stack:
                return log, tmax
            c, t = stack.pop()
            if t < tmax:          # the re-check with the tMax of this moment
                cur = c
    return log, tmax


def main():
    n_tris = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    hs = pb.HostScene.soup(n_tris, seed=11, jitter=0.25, xres=16, yres=16, spp=1)    # fat triangles: a ray crosses many leaves
    nodes, prims = hs.nodes(), hs.bvh_prims(0)
    d = hs.desc.contents
    P = np.ctypeslib.as_array(d.P, shape=(d.n_vertices, 3))
    idx = np.ctypeslib.as_array(d.tri_index, shape=(d.n_tris, 3))
    assert d.n_prims == d.n_tris
    tri = P[idx][np.ctypeslib.as_array(d.prim_index, shape=(d.n_prims,))]       # per primitive number
    pbounds = np.stack([tri.min(axis=1), tri.max(axis=1)], 1).astype(f32)
    rays = gc.rays_for(pb, nodes, n_rays, 2)
    # aim most rays at triangles, so that leaves are reached and tMax shrinks along the way; keep the axis-parallel ones
    rng = np.random.RandomState(3)
    target = tri[rng.randint(0, len(tri), n_rays)].mean(axis=1)
    aimed = (target - rays["o"]).astype(f32)
    keep = np.arange(n_rays) < n_rays // 20
    rays["d"] = np.where(keep[:, None], rays["d"], aimed)
    same, visits_b, visits_w = 0, 0, 0
    for r in rays:
        o, dd = r["o"].astype(f32), r["d"].astype(f32)
        lb, tb = binary(nodes, prims, pbounds, o, dd)
        lw, tw = wide4(nodes, prims, pbounds, o, dd)
        assert lb == lw and (tb == tw or (np.isinf(tb) and np.isinf(tw))), "the two traversals must test the same primitives in the same order"
        same += 1
        visits_b += len(lb)
    print("wide4 prototype: %d rays over %d nodes: identical primitive-test sequences (%d tests in total)" % (same, len(nodes), visits_b))


if __name__ == "__main__":
    main()
