"""Multinerve RDT on the device against the oracle restatement (itself bit-equal to the live reference), order-independent form."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphitethree_b200 import capi, shapes
from oracle import port

def unoriented(t):
    return np.unique(np.sort(t, axis=1), axis=0)

ok_all = True
cases = [("ico", shapes.icosphere(12), 300, 3), ("trefoil", shapes.trefoil_tube(300, 24), 1500, 3), ("box", shapes.box_surface(10), 400, 3),
         ("thinbox", shapes.box_surface(8, (1.0, 1.0, 0.02)), 150, 3), ("thinbox_raw", shapes.box_surface(8, (1.0, 1.0, 0.02)), 150, 0),
         ("noise", shapes.noise_sphere(60), 8000, 2)]
for name, (V, F), S, nl in cases:
    X = shapes.sample_surface(V, F, S, 3)
    x = port.lloyd(V, F, X, nl)[0] if nl else X
    for uc, ps in ((0, 0), (1, 1), (1, 0)):
        to, vo, so = port.rdt_multinerve(V, F, x, uc, ps)
        co = port.canonical_multinerve(to, vo, so)
        h = capi.Handle(3)
        h.set_mesh(V, F)
        h.set_seeds(x)
        t0 = time.time()
        tg, vg, sg = h.rdt_multinerve(bool(uc), bool(ps))
        dt = time.time() - t0
        fl = h.flags()
        h.close()
        cg = port.canonical_multinerve(tg, vg, sg)
        same_v = cg[1].shape == co[1].shape and np.array_equal(cg[2], co[2]) and np.abs(cg[1] - co[1]).max() <= 1e-12
        same_t = np.array_equal(unoriented(cg[0]), unoriented(co[0]))
        ok = same_v and same_t
        ok_all &= ok
        print("%-12s centroids=%d prefer=%d: components %d/%d (seeds with several: %d)  triangles %d/%d  vertices %s  triangle set %s  flags %s  %.3f s" % (
            name, uc, ps, len(vg), len(vo), int(np.sum(np.bincount(so, minlength=S) > 1)), len(unoriented(cg[0])), len(unoriented(co[0])),
            "ok" if same_v else "DIFF", "ok" if same_t else "DIFF", np.unique(fl).tolist(), dt), flush=True)
        if not ok and cg[1].shape == co[1].shape:
            d = np.abs(cg[1] - co[1]).max(1)
            print("   max vertex diff", d.max(), "at", np.argmax(d), "seed", co[2][np.argmax(d)])
            a, b = set(map(tuple, unoriented(cg[0]))), set(map(tuple, unoriented(co[0])))
            print("   only gpu", list(a - b)[:5], "only oracle", list(b - a)[:5])
print("ALL OK" if ok_all else "FAILURES")
