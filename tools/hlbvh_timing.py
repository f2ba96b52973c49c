"""Developer script: wall time of the three HLBVH builds (host threads; device treelets + host upper tree and flatten; everything
on the device) on the synthetic soup, and a check that the three give the same nodes.  usage: hlbvh_timing.py <n_tris>"""
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np

import pbrt_v3_b200 as pb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
os.environ["PB2_SOUP_SPLIT"] = "hlbvh"
os.environ["PB2_VERBOSE"] = "1"
pb.init(0)
ref = None
for label, dev, upper in (("host threads", "0", "host"), ("device treelets, host upper tree + flatten", "1", "host"),
                          ("all on the device", "1", "device"), ("all on the device (again)", "1", "device")):
    os.environ["PB2_DEVICE_BVH"] = dev
    os.environ["PB2_DEVICE_BVH_UPPER"] = upper
    t0 = time.time()
    hs = pb.HostScene.soup(n, xres=64, yres=36, spp=1)
    nodes = hs.nodes()   # forces the build
    dt = time.time() - t0
    h = hash(nodes.tobytes())
    if ref is None:
        ref = nodes.copy()
    same = len(nodes) == len(ref) and all(np.array_equal(nodes[f], ref[f]) for f in ("bmin", "bmax", "offset", "n_prims"))
    print("hlbvh %d tris, %s: scene ready in %.2f s, %d nodes, equal to the host build: %s" % (n, label, dt, len(nodes), same), flush=True)
