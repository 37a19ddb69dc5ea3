"""Per-phase device timings of Lloyd iterations at bench sizes (no oracle)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphitethree_b200 import shapes, capi

def run(freq, S, iters=6, noise=False):
    t = time.time()
    V, F = (shapes.noise_sphere(freq) if noise else shapes.icosphere(freq))
    X = shapes.sample_surface(V, F, S, 1)
    print("mesh %d tri, %d seeds, gen %.1fs" % (F.shape[0], S, time.time() - t), flush=True)
    h = capi.Handle(3)
    t = time.time(); h.set_mesh(V, F); print("set_mesh %.2fs" % (time.time() - t))
    x = X
    for it in range(iters):
        t = time.time(); x = h.lloyd(x, 1); dt = time.time() - t
        tm = h.timings()
        print("iter %d wall %.2f ms | dev %s | flagged %d" % (it, dt * 1e3, " ".join("%s=%.3f" % kv for kv in tm.items()),
              int((h.flags() & 1).astype(bool).sum()) if False else -1), flush=True)
    t = time.time(); x = h.lloyd(x, 10); dt = time.time() - t
    print("10 iters (incl H2D/D2H) %.2f ms -> %.3e seed-iter/s" % (dt * 1e3, S * 10 / dt))
    print(h.stats())
    f, g = h.funcgrad(True); print("funcgrad f=%.6g" % f, h.timings(), h.stats())
    h.close()

if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), noise=len(sys.argv) > 3)
