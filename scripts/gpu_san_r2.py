"""Small run through every kernel family added in round 2, for compute-sanitizer (memcheck / racecheck):
slot-polygon clip kernel (Lloyd + Newton), simple and multinerve RDT, initial sampling, volumetric cell-first path + tet path."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphitethree_b200 import capi, shapes
V, F = shapes.icosphere(8)
X = shapes.sample_surface(V, F, 400, 1)
h = capi.Handle(3)
h.set_mesh(V, F)
x0, ok = h.initial_sampling(300)
x = h.lloyd(X, 3)
x, info = h.newton(x, 3, 5)
tri = h.rdt()
mn = h.rdt_multinerve(True, True)
h.close()
Vt, T = shapes.kuhn_cube(6)
Xv = 0.02 + 0.96 * np.random.default_rng(3).random((T.shape[0] // 10, 3))
for vcell in ("1", "0"):
    os.environ["B200CVT_VCELL"] = vcell
    hv = capi.Handle(3, volumetric=True)
    hv.set_mesh(Vt, T)
    xv = hv.lloyd(Xv, 2)
    xv, info = hv.newton(xv, 2, 5)
    hv.set_seeds(xv); mg, m = hv.centroids(True)
    assert abs(m.sum() - 1.0) < 1e-12
    hv.close()
print("sanitizer run ok: surface %d seeds, %d RDT triangles, %d multinerve vertices, volumetric %d seeds" % (x.shape[0], tri.shape[0], mn[1].shape[0], xv.shape[0]))
