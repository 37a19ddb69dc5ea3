"""BASELINE.json configs C3 (trefoil tube 20 M triangles / 5 M seeds, Lloyd, here on ONE B200) and C4 (6D anisotropic
CAD-like surface 2 M triangles / 500 k seeds): seed-iterations/s, phase times and size-independent properties.
usage: gpu_configs.py [c3] [c4] [--small] [--ref]"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphitethree_b200 import capi, shapes


def area(V, F):
    P = V[F.astype(np.int64)]
    e1, e2 = P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]
    return 0.5 * np.sqrt(np.maximum((e1 * e1).sum(1) * (e2 * e2).sum(1) - ((e1 * e2).sum(1)) ** 2, 0.0)).sum()


def run(name, V, F, S, newton=0):
    X = shapes.sample_surface(V, F, S, 1)
    h = capi.Handle(V.shape[1])
    t0 = time.time(); h.set_mesh(V, F); t_mesh = time.time() - t0
    x = h.lloyd(X, 4)
    h.cumulative(reset=True)
    t0 = time.time(); x = h.lloyd(x, 5); t_l = (time.time() - t0) / 5
    c = h.cumulative(reset=True)
    out = {"config": name, "dim": int(V.shape[1]), "triangles": int(F.shape[0]), "seeds": int(S), "set_mesh_s": round(t_mesh, 2),
           "lloyd_ms_per_iter_e2e": round(t_l * 1e3, 3), "seed_iterations_per_s_lloyd": S / t_l,
           "phase_ms": {k: round(c[k] / c["evals"], 3) for k in ("sort", "knn", "pairs", "clip", "clip_kernel")}}
    if newton:
        t0 = time.time(); x, info = h.newton(x, newton, 7); t_n = time.time() - t0
        out["newton"] = {"iters": info["iters"], "nfev": info["nfev"], "ms_per_evaluation_e2e": round(t_n * 1e3 / info["nfev"], 3),
                         "seed_iterations_per_s": S * info["nfev"] / t_n}
    h.set_seeds(x)
    mg, m = h.centroids(True)
    fl = h.flags()
    A = area(V, F)
    out["sum_m_rel_err"] = float(abs(m.sum() - A) / A)                 # the cells tile the surface
    h.set_seeds(x)
    f, g = h.funcgrad(True)
    out["g_identity_rel"] = float(np.abs(g - 2.0 * (m[:, None] * x - mg)).max() / np.abs(g).max())
    out["flags_overflow_or_kmax"] = int((fl & (capi.FLAG_POLY_OVERFLOW | capi.FLAG_KMAX)).sum())
    h.close()
    if "--ref" in sys.argv:
        from oracle import ref
        if ref.available():
            r = ref.RefCVT(V, F, multithread=True)
            r.set_points(x)
            r.lloyd(1)                         # thread partition of the mesh
            t = r.lloyd(2)
            out["reference_seed_iterations_per_s_lloyd"] = S * 2 / t
            out["reference_threads"] = ref.RefCVT.nb_threads()
            r.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    small = "--small" in sys.argv
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c3", "c4"]
    if "c4" in which:
        V, F = shapes.cad_like(60 if small else 268)                   # 28 n^2 -> 2.0 M triangles
        run("C4", shapes.lift_anisotropic(V, F, 0.04), F, 20000 if small else 500000, newton=5)
    if "c3" in which:
        V, F = shapes.trefoil_tube(1000 if small else 10000, 100 if small else 1000)
        run("C3 on one GPU", V, F, 50000 if small else 5000000)
