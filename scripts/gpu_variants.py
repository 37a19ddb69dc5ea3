"""A/B timing of library variants (built with -D knobs into gpurun_in/lib_<name>.so): phase times of steady-state Lloyd
evaluations and of a short Newton run at C2 size. usage: gpu_variants.py <name> [<name> ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, numpy as np
sys.path.insert(0, %(root)r)
from graphitethree_b200 import capi, shapes
capi.LIB_PATH = %(lib)r
d = np.load("/tmp/c2_input.npz")
V, F, X = d["V"], d["F"], d["X"]
h = capi.Handle(3)
h.set_mesh(V, F)
x = h.lloyd(X, 6)
h.cumulative(reset=True)
x = h.lloyd(x, 5)
c = h.cumulative(reset=True)
n = c["evals"]
print("%(name)s lloyd  " + " ".join("%%s=%%.3f" %% (k, c[k] / n) for k in ("sort", "knn", "pairs", "clip", "clip_kernel")), "sum=%%.3f" %% ((c["sort"] + c["knn"] + c["pairs"] + c["clip"]) / n))
import time
h.newton(x, 10, 7)                       # warm-up: allocations, cooperative-launch set-up
t0 = time.time(); xn, info = h.newton(x, 20, 7); t1 = time.time()
c = h.cumulative(reset=True); n = c["evals"]
print("%(name)s newton " + " ".join("%%s=%%.3f" %% (k, c[k] / n) for k in ("sort", "knn", "pairs", "clip", "clip_kernel")), "evals=%%d wall_ms_per_eval=%%.3f lbfgs_dir_ms_per_iter=%%.4f" %% (n, (t1 - t0) * 1e3 / n, c["cells"] / max(info["iters"], 1)))
h.close()
'''
if not os.path.exists("/tmp/c2_input.npz"):
    sys.path.insert(0, ROOT)
    import numpy as np
    from graphitethree_b200 import shapes
    V, F = shapes.noise_sphere(316)
    X = shapes.sample_surface(V, F, 200000, 1)
    np.savez("/tmp/c2_input.npz", V=V, F=F, X=X)
for name in sys.argv[1:]:
    lib = os.path.join(ROOT, "gpurun_in", "lib_%s.so" % name)
    r = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT, lib=lib, name=name)], capture_output=True, text=True)
    sys.stdout.write(r.stdout)
    if r.returncode != 0:
        sys.stdout.write(r.stderr[-2000:])
    sys.stdout.flush()
