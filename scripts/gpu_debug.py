"""First-contact GPU check: CUDA path vs the port oracle on small seeded inputs, with timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphitethree_b200 import shapes, capi
from oracle import port

def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)

def run(freq, S, iters=3, seed=1):
    V, F = shapes.icosphere(freq)
    X = shapes.sample_surface(V, F, S, seed)
    h = capi.Handle(3)
    h.set_mesh(V, F)
    t = time.time(); h.set_seeds(X); print("set_seeds %.3fs" % (time.time() - t))
    idx, cnt, sqd, fl = h.knn(20)
    pidx, pcnt, psqd, ptie = port.knn(X, 20)
    print("knn: idx equal", np.array_equal(idx, pidx), "cnt", np.array_equal(cnt, pcnt), "sqd bit-equal", np.array_equal(sqd, psqd),
          "mismatch rows", int((idx != pidx).any(1).sum()), "ties gpu/port", int((fl & 2).astype(bool).sum()), int(ptie.sum()))
    h.set_seeds(X)
    t = time.time(); mg, m = h.centroids(False); print("centroids %.3fs" % (time.time() - t), h.timings())
    e = port.surface_eval(V, F, X, 0, False)
    ok = (e.flags & 1) == 0
    print("m rel all %.3e  unflagged %.3e | mg rel unflagged %.3e | flagged oracle %d gpu %d" % (
        rel(m, e.m), rel(m[ok], e.m[ok]), rel(mg[ok], e.mg[ok]), int((~ok).sum()), int((h.flags() & 1).astype(bool).sum())))
    bad = np.nonzero(np.abs(m - e.m) > 1e-9 * e.m.max())[0]
    print("seeds with |dm|>1e-9:", len(bad), "of which flagged:", int((~ok[bad]).sum()))
    print("stats", h.stats(), "oracle", e.counters)
    f, g = h.funcgrad(True)
    e2 = port.surface_eval(V, F, X, 1, True)
    print("f gpu %.17g oracle %.17g rel %.3e | g rel %.3e | f_seed rel %.3e | flags %s" % (
        f, e2.f, abs(f - e2.f) / e2.f, rel(g, e2.g), rel(h.seed_energy(), e2.f_seed), np.unique(h.flags())))
    print("stats", h.stats())
    t = time.time(); xl = h.lloyd(X, iters); dt = time.time() - t
    xo, _ = port.lloyd(V, F, X, iters)
    print("lloyd %d iters %.3fs (%.0f seed-iter/s) max|dx| vs oracle %.3e" % (iters, dt, S * iters / dt, np.abs(xl - xo).max()), h.timings())
    t = time.time(); xn, info = h.newton(xl, 3, 7); dt = time.time() - t
    xno, oinfo = port.newton(V, F, xo, 3, 7)
    print("newton", info, "%.3fs" % dt, "oracle iters/nfev", oinfo["iters"], oinfo["nfev"], "max|dx| %.3e" % np.abs(xn - xno).max())
    h.close()

if __name__ == "__main__":
    freq = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    run(freq, S)
