"""C4 (6D anisotropic) evaluations for ncu. usage: gpu_prof_c4.py [n] [S]"""
import sys, os
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
h.cumulative(reset=True)
h.stats(); h.set_seeds(x); h.centroids(False); print(h.stats()); print(h.cumulative())
