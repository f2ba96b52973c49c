"""Developer script: device time of pb2_hlbvh_treelets (Morton codes + sort + treelets) for n random primitive bounds."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import pbrt_v3_b200 as pb  # noqa: E402

L = pb.lib()
pb.init()
L.pb2_hlbvh_treelets.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
for n in [int(a) for a in sys.argv[1:]] or [1000000]:
    rng = np.random.RandomState(1)
    c = rng.rand(n, 3).astype(np.float32)
    r = (0.002 * rng.rand(n, 3)).astype(np.float32)
    bounds = np.concatenate([c - r, c + r], 1).astype(np.float32)
    pool = np.zeros((2 * n, 11), np.int32)
    ordered = np.zeros(n, np.int32)
    roots = np.zeros(4096, np.int32)
    nt, ms = C.c_int32(), C.c_double()
    for it in range(3):
        t = time.time()
        rc = L.pb2_hlbvh_treelets(pb.ptr(bounds), n, 4, pb.ptr(pool), pb.ptr(ordered), pb.ptr(roots), C.byref(nt), C.byref(ms))
        wall = time.time() - t
    leaves = int((pool[:, 10] > 0).sum())
    print("hlbvh probe: n %d rc %d: %d treelets, %d leaves; kernels %.2f ms (%.0f Mprims/s), call incl. copies %.1f ms"
          % (n, rc, nt.value, leaves, ms.value, n / ms.value / 1e3 if ms.value else 0, wall * 1e3))
