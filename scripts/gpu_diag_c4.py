import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from graphitethree_b200 import capi, shapes
n=int(sys.argv[1]); S=int(sys.argv[2])
V,F=shapes.cad_like(n); V6=shapes.lift_anisotropic(V,F,0.04)
X=shapes.sample_surface(V6,F,S,1)
h=capi.Handle(6); h.set_mesh(V6,F)
x=h.lloyd(X,3)
h.stats()
h.set_seeds(x); h.centroids(False)
print(h.stats()); print(h.timings())
