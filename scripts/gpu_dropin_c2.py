"""The drop-in checker (integration/_build/dropin_check: stock CentroidalVoronoiTesselation vs the B200 adapter classes through
the reference's own C++ API) at C2 size: 2 M triangles, 200 k seeds, 10 Lloyd + 30 Newton, from the sampling after two stock Lloyd iterations (no truncated cell is left there: the oracle flags 194 seeds after one iteration, none after two). Prints the checker's JSON line."""
import os, sys, subprocess, tempfile, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
sys.path.insert(0, os.path.join(ROOT, "tests"))
V, F, X = bench.workload(1, "--small" in sys.argv)
d = tempfile.mkdtemp()
mp, sp = os.path.join(d, "mesh.bin"), os.path.join(d, "seeds.bin")
with open(mp, "wb") as f:
    f.write(np.array([V.shape[0], F.shape[0], V.shape[1]], dtype=np.uint32).tobytes())
    f.write(np.ascontiguousarray(V, dtype=np.float64).tobytes()); f.write(np.ascontiguousarray(F, dtype=np.uint32).tobytes())
with open(sp, "wb") as f:
    f.write(np.array([X.shape[0], X.shape[1]], dtype=np.uint32).tobytes()); f.write(np.ascontiguousarray(X, dtype=np.float64).tobytes())
exe = os.path.join(ROOT, "integration", "_build", "dropin_check")
out = subprocess.run([exe, mp, sp, "10", "30", "7", "2", "0"], capture_output=True, text=True, timeout=1500)
print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-2000:])
