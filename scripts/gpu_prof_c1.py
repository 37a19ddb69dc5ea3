"""C1 size (Graphite's default sphere: 20 480 triangles, 5 000 seeds): where a launch-bound Lloyd iteration goes.
usage: [ncu --metrics gpu__time_duration.sum ...] python scripts/gpu_prof_c1.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphitethree_b200 import capi, shapes
V, F = shapes.icosphere_split(5)
X = shapes.sample_surface(V, F, 5000, 1)
h = capi.Handle(3)
h.set_mesh(V, F)
x = h.lloyd(X, 5)
h.cumulative(reset=True)
t0 = time.time(); x = h.lloyd(x, 30); t = time.time() - t0
c = h.cumulative(reset=True)
print("30 Lloyd iterations: %.2f ms wall = %.3f ms per iteration; device phases per iteration:" % (t * 1e3, t * 1e3 / 30),
      {k: round(c[k] / c["evals"], 4) for k in ("sort", "knn", "pairs", "clip", "clip_kernel")}, "launches", h.launch_count())
t0 = time.time(); xn, info = h.newton(x, 30, 7); t = time.time() - t0
print("Newton 30: %.2f ms wall, %d evaluations = %.3f ms per evaluation" % (t * 1e3, info["nfev"], t * 1e3 / info["nfev"]))
