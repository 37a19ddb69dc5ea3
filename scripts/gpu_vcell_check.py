"""Volumetric cell-first path (vcell.cuh) against the (tet, seed) path of the same library (B200CVT_VCELL=0): per-seed mass,
first moment, energy and gradient on a Kuhn cube, raw and relaxed seeds, Lloyd mode and exact cells; counts of the two paths.
usage: gpu_vcell_check.py [n ...]"""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphitethree_b200 import capi, shapes

def handle(vcell, V, T):
    os.environ["B200CVT_VCELL"] = "1" if vcell else "0"
    os.environ["B200CVT_VCELL_TET"] = os.environ.get("VCELL_TET", "1")
    h = capi.Handle(3, volumetric=True)
    h.set_mesh(V, T)
    return h

def run(n):
    V, T = shapes.kuhn_cube(n)
    S = T.shape[0] // 10
    X = 0.01 + 0.98 * np.random.default_rng(5).random((S, 3))
    ha, hb = handle(False, V, T), handle(True, V, T)
    out = {"n": n, "tets": int(T.shape[0]), "seeds": int(S)}
    for tag, x in (("raw", X), ("relaxed", ha.lloyd(X, 4))):
        for sr in (False, True):
            res = {}
            for name, h in (("tet", ha), ("cell", hb)):
                h.stats()
                h.set_seeds(x); mg, m = h.centroids(sr); fl0 = h.flags().copy()
                st = h.stats()
                h.set_seeds(x); f, g = h.funcgrad(sr); fl1 = h.flags().copy()
                res[name] = (m, mg, f, g, fl0, fl1, st)
            a, b = res["tet"], res["cell"]
            ok = ((a[4] | b[4] | a[5] | b[5]) & (capi.FLAG_POLY_OVERFLOW | capi.FLAG_KMAX)) == 0
            key = "%s_sr%d" % (tag, int(sr))
            out[key] = {"dm": float(np.abs(a[0] - b[0])[ok].max() / a[0].max()), "dmg": float(np.abs(a[1] - b[1])[ok].max() / np.abs(a[1]).max()),
                        "df": float(abs(a[2] - b[2]) / abs(a[2])), "dg": float(np.abs(a[3] - b[3])[ok].max() / np.abs(a[3]).max()),
                        "exh_tet": int((a[4] & 1).sum()), "exh_cell": int((b[4] & 1).sum()),
                        "sum_m": float(b[0].sum()), "cells_direct": b[6]["volumetric_cells"]["direct"], "cells_tet_path": b[6]["volumetric_cells"]["tet_path"],
                        "redo": b[6]["redo_seeds"]}
    if S <= 5000:
        # Lloyd mode (20 neighbours, no radius check) against the oracle on EVERY seed, truncated cells included
        from oracle import port
        x = ha.lloyd(X, 4)
        e = port.surface_eval(V, T, x, 0, False)
        for name, h in (("tet", ha), ("cell", hb)):
            h.set_seeds(x); mg, m = h.centroids(False)
            bad = np.abs(m - e.m) > 1e-9 * e.m.max()
            out["lloyd_mode_vs_oracle_" + name] = {"seeds_off": int(bad.sum()), "of_which_flagged_by_oracle": int((bad & ((e.flags & 1) != 0)).sum()),
                                                   "oracle_flagged": int((e.flags & 1).sum()), "max_rel": float(np.abs(m - e.m).max() / e.m.max())}
    # trajectories
    x0 = ha.lloyd(X, 2)
    t0 = time.time(); xa = ha.lloyd(x0, 5); ta = time.time() - t0
    t0 = time.time(); xb = hb.lloyd(x0, 5); tb = time.time() - t0
    out["lloyd5_dx"] = float(np.abs(xa - xb).max()); out["lloyd_ms_tet"] = ta * 200; out["lloyd_ms_cell"] = tb * 200
    xa, ia = ha.newton(xa, 10, 7); xb2, ib = hb.newton(xb, 10, 7)
    out["newton_dx"] = float(np.abs(xa - xb2).max()); out["newton_info"] = [str(ia), str(ib)]
    ha.close(); hb.close()
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    for n in [int(a) for a in sys.argv[1:]] or [12, 30]:
        run(n)
