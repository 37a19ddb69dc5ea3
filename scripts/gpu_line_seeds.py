"""Seeds on a line next to two facets much longer than their spacing (the neighbour-cap provocation of tests/test_gpu_parity.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphitethree_b200 import capi
V = np.array([[-1.0, -5.0, 0.0], [2.0, -5.0, 0.0], [2.0, 5.0, 0.0], [-1.0, 5.0, 0.0]])
F = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 200
x = np.column_stack([(np.arange(S) + 0.5) / S, 1e-3 * np.sin(np.arange(S)), np.zeros(S)])
h = capi.Handle(3)
h.set_mesh(V, F)
h.set_seeds(x)
mg, m = h.centroids(False)
print("lloyd-mode ok, sum m", m.sum())
h.set_seeds(x)
f, g = h.funcgrad(True)
print("f", f, "flags", np.unique(h.flags(), return_counts=True))
