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
print("%(name)s", " ".join("%%s=%%.3f" %% (k, c[k] / c["evals"]) for k in ("knn", "pairs", "clip")))
'''
n = int(sys.argv[1])
for name in sys.argv[2:]:
    lib = os.path.join(ROOT, "gpurun_in", "lib_%s.so" % name)
    r = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT, lib=lib, name=name, n=n)], capture_output=True, text=True)
    sys.stdout.write(r.stdout + (r.stderr[-1500:] if r.returncode else "")); sys.stdout.flush()
