"""Volumetric evaluation for ncu (one Lloyd-mode evaluation on relaxed seeds). usage: gpu_prof_volume.py [n]"""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from graphitethree_b200 import capi, shapes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 47
V, T = shapes.kuhn_cube(n); S = T.shape[0] // 10
X = 0.01 + 0.98 * np.random.default_rng(5).random((S, 3))
h = capi.Handle(3, volumetric=True); h.set_mesh(V, T)
x = h.lloyd(X, 4)
h.stats(); h.set_seeds(x); h.centroids(False); print(h.stats()); print(h.cumulative())
