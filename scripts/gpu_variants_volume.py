"""A/B timing of library variants (gpurun_in/lib_<name>.so) on the volumetric path. usage: gpu_variants_volume.py n name..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r)
from graphitethree_b200 import capi, shapes
capi.LIB_PATH = %(lib)r
V, T = shapes.kuhn_cube(%(n)d); S = T.shape[0] // 10
X = 0.01 + 0.98 * np.random.default_rng(5).random((S, 3))
h = capi.Handle(3, volumetric=True); h.set_mesh(V, T)
x = h.lloyd(X, 3); h.cumulative(reset=True); x = h.lloyd(x, 3); c = h.cumulative(reset=True)
import time
h.set_seeds(x); h.funcgrad(True); h.cumulative(reset=True); h.stats(); h.set_seeds(x); t0 = time.time(); h.funcgrad(True); tf = time.time() - t0
cf = h.cumulative(reset=True); st = h.stats()
print("  funcgrad phases", " ".join("%%s=%%.3f" %% (k, cf[k]) for k in ("sort", "knn", "pairs", "clip", "clip_kernel", "cells")), "redo_total", st["redo_seeds"], st["volumetric_cells"])
print("%(name)s", " ".join("%%s=%%.3f" %% (k, c[k] / c["evals"]) for k in ("knn", "pairs", "clip", "clip_kernel", "cells")), "funcgrad_ms=%%.2f" %% (tf * 1e3))
'''
n = int(sys.argv[1])
for name in sys.argv[2:]:
    lib = os.path.join(ROOT, "gpurun_in", "lib_%s.so" % name)
    r = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT, lib=lib, name=name, n=n)], capture_output=True, text=True)
    sys.stdout.write(r.stdout + (r.stderr[-1500:] if r.returncode else "")); sys.stdout.flush()
