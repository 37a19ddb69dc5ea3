"""C5 (volumetric RVD): Lloyd + func/grad on the unit cube split into n^3 x 6 Kuhn tets, S = T/10 seeds; seed-iterations/s
per size, size-independent properties, and the reference (oracle/_ref, all host threads) at the smallest size.
usage: gpu_volume.py [n ...]   (n = 119 -> 10.1 M tets, 1 M seeds)"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphitethree_b200 import capi, shapes

def run(n, ref_too=False):
    V, T = shapes.kuhn_cube(n)
    S = T.shape[0] // 10
    X = 0.01 + 0.98 * np.random.default_rng(5).random((S, 3))
    h = capi.Handle(3, volumetric=True)
    t0 = time.time(); h.set_mesh(V, T); t_mesh = time.time() - t0
    x = h.lloyd(X, 3)                      # relaxation + warm-up
    h.cumulative(reset=True)
    t0 = time.time(); x = h.lloyd(x, 3); t_l = (time.time() - t0) / 3
    c = h.cumulative(reset=True)
    h.set_seeds(x)
    t0 = time.time(); f, g = h.funcgrad(True); t_f = time.time() - t0
    h.set_seeds(x)
    mg, m = h.centroids(True)
    out = {"n": n, "tets": int(T.shape[0]), "seeds": int(S), "set_mesh_s": round(t_mesh, 2),
           "lloyd_ms_per_iter_e2e": round(t_l * 1e3, 2), "seed_iterations_per_s_lloyd": S / t_l,
           "funcgrad_ms_e2e": round(t_f * 1e3, 2),
           "phase_ms": {k: round(c[k] / c["evals"], 3) for k in ("sort", "knn", "pairs", "clip", "clip_kernel", "cells")},
           "sum_m_minus_1": float(m.sum() - 1.0), "g_identity": float(np.abs(g - 2.0 * (m[:, None] * x - mg)).max()),
           "flags": int((h.flags() & (capi.FLAG_POLY_OVERFLOW | capi.FLAG_KMAX)).sum())}
    h.close()
    if ref_too:
        from oracle import ref
        if ref.available():
            r = ref.RefCVT(V, T, volumetric=True, multithread=True)
            r.set_points(x)
            r.lloyd(1)
            t = r.lloyd(2)
            out["reference_seed_iterations_per_s"] = S * 2 / t
            out["reference_threads"] = ref.RefCVT.nb_threads()
            r.close()
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    ns = [int(a) for a in sys.argv[1:]] or [47, 75, 119]
    for i, n in enumerate(ns):
        run(n, ref_too=(i == 0))
