"""Profiling target: a few steady-state Lloyd evaluations and one func/grad evaluation at C2 size (no oracle)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphitethree_b200 import shapes, capi

def run(freq, S, noise=True, iters=4):
    V, F = (shapes.noise_sphere(freq) if noise else shapes.icosphere(freq))
    X = shapes.sample_surface(V, F, S, 1)
    h = capi.Handle(3)
    h.set_mesh(V, F)
    x = h.lloyd(X, iters)
    print("lloyd", h.timings())
    h.set_seeds(x)
    f, g = h.funcgrad(True)
    print("funcgrad f=%.6g" % f, h.timings())
    h.stats()
    h.set_seeds(x)
    h.centroids(False)
    print("stats", h.stats())
    h.close()

if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), iters=int(sys.argv[3]) if len(sys.argv) > 3 else 4)
