"""Walk / redo statistics of the 6D anisotropic case (C4) in Lloyd (check_SR = false) and Newton (check_SR = true) mode."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphitethree_b200 import capi, shapes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 268
S = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
V, F = shapes.cad_like(n)
V6 = shapes.lift_anisotropic(V, F, 0.04)
X = shapes.sample_surface(V6, F, S, 1)
h = capi.Handle(6)
h.set_mesh(V6, F)
x = h.lloyd(X, 4)
for mode, sr in (("lloyd", False), ("newton", True)):
    h.stats()
    h.set_seeds(x)
    t0 = time.time()
    if sr: h.funcgrad(True)
    else: h.centroids(False)
    t = time.time() - t0
    st = h.stats()
    fl = h.flags()
    print(mode, "wall_ms=%.2f" % (t * 1e3), {k: st[k] for k in ("redo_seeds", "candidate_pairs", "nonempty_pairs", "facets_uncertified", "facets_subdivided")},
          "flag_exhausted=%d kmax=%d" % ((fl & 1).sum(), (fl & 8).sum()), h.timings())
