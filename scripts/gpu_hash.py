"""Bit-level regression fingerprints of the C-ABI results on fixed procedural inputs (run under gpurun).

Every kernel rewrite of round 2 must leave these hashes unchanged (the clipped polygons, the integration order and the
reduction trees are results, not implementation details): `python scripts/gpu_hash.py > gpurun_out/hash_<tag>.json`,
then compare with profiles/r2_hashes_baseline.json (the round-1 kernels)."""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphitethree_b200 import shapes, capi


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def surface_case(name, V, F, S, out, weights=None, dim=3, lloyd=3, newton=4):
    if dim == 6:
        V = shapes.lift_anisotropic(V, F, 0.04)
    X = shapes.sample_surface(V, F, S, 1)
    h = capi.Handle(dim)
    h.set_mesh(V, F, weights=weights)
    t0 = time.time()
    h.set_seeds(X)
    mg, m = h.centroids(False)
    fl = h.flags()
    out[name + ".lloyd_eval"] = sha(m, mg, fl)
    h.set_seeds(X)
    f, g = h.funcgrad(True)
    out[name + ".funcgrad"] = sha(np.array([f]), g, h.seed_energy(), h.flags())
    x = h.lloyd(X, lloyd)
    out[name + ".lloyd%d" % lloyd] = sha(x)
    x2, info = h.newton(x, newton, 7)
    out[name + ".newton%d" % newton] = sha(x2) + "/%d/%d" % (info["iters"], info["nfev"])
    h.set_seeds(x2)
    out[name + ".rdt"] = sha(h.rdt())
    out[name + ".seconds"] = round(time.time() - t0, 2)
    h.close()


def main():
    out = {}
    V, F = shapes.noise_sphere(120)
    surface_case("noise120_s30k", V, F, 30000, out)
    V, F = shapes.noise_sphere(316)
    surface_case("c2_noise316_s200k", V, F, 200000, out, lloyd=2, newton=2)
    V, F = shapes.icosphere(30)
    surface_case("ico30_s9k_ts2", V, F, 9000, out)
    V, F = shapes.trefoil_tube(400, 40)
    surface_case("trefoil_s4k", V, F, 4000, out)
    V, F = shapes.box_surface(20)
    w = 1.0 + 4.0 * (V[:, 0] - V[:, 0].min())
    surface_case("box_weighted_s1k", V, F, 1000, out, weights=w)
    V, F = shapes.cad_like(40)
    surface_case("cad6d_s3k", V, F, 3000, out, dim=6, lloyd=2, newton=2)
    # regular lattice seeds on a flat box face: exact ties and on-plane vertices
    V, F = shapes.box_surface(12)
    g = (np.arange(10) + 0.5) / 10.0
    P = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    lo, hi = V.min(0), V.max(0)
    X = np.concatenate([np.column_stack([lo[0] + P[:, 0] * (hi[0] - lo[0]), lo[1] + P[:, 1] * (hi[1] - lo[1]), np.full(len(P), z)]) for z in (lo[2], hi[2])])
    h = capi.Handle(3)
    h.set_mesh(V, F)
    h.set_seeds(X)
    mg, m = h.centroids(False)
    out["box_lattice.lloyd_eval"] = sha(m, mg)
    h.set_seeds(X)
    f, gg = h.funcgrad(True)
    out["box_lattice.funcgrad"] = sha(np.array([f]), gg)
    h.close()
    # volumetric
    V, T = shapes.kuhn_cube(12)
    X = shapes.sample_volume(V, T, 1500, 1)
    h = capi.Handle(3, volumetric=True)
    h.set_mesh(V, T)
    h.set_seeds(X)
    mg, m = h.centroids(False)
    out["kuhn12_s1500.lloyd_eval"] = sha(m, mg)
    h.set_seeds(X)
    f, gg = h.funcgrad(True)
    out["kuhn12_s1500.funcgrad"] = sha(np.array([f]), gg)
    h.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
