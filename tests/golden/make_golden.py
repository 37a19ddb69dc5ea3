"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile.ref). Run in the build container only:

    python tests/golden/make_golden.py

Every array is an output of the reference's own classes (GEO::Delaunay "NN",
GEO::RestrictedVoronoiDiagram, GEO::CentroidalVoronoiTesselation) driven single-threaded
through oracle/ref_driver.cpp on procedural inputs from graphitethree_b200/shapes.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from graphitethree_b200 import shapes  # noqa: E402
from oracle.ref import RefCVT  # noqa: E402


def case(V, F, X, weights=None, lloyd_iters=5, newton_iters=4):
    out = dict(V=V, F=F, X=X)
    if weights is not None:
        out["weights"] = weights

    # only one CentroidalVoronoiTesselation may exist at a time (instance_, CVT.cpp:53-70), and
    # neighbour-list sizes are sticky (delaunay.cpp:260-268): one fresh object per quantity
    def fresh(x):
        r = RefCVT(V, F, weights=weights, multithread=False)
        r.set_points(x)
        return r
    r = fresh(X); r.update_delaunay()
    out["knn_idx"], out["knn_cnt"] = r.neighbors(20)
    out["mg"], out["m"], _ = r.centroids(False)
    r.close()
    r = fresh(X); r.update_delaunay()
    out["f"], out["g"], _ = r.funcgrad(True)
    out["f"] = np.float64(out["f"])
    r.close()
    r = fresh(X); r.update_delaunay()
    out["f_seed"], _, cnt, pairs = r.polygons(True, want_pairs=True, pairs_cap=64 * X.shape[0])
    out["pairs_exact"] = pairs
    r.close()
    r = fresh(X); r.lloyd(lloyd_iters)
    out["x_lloyd"] = r.points()
    out["lloyd_iters"] = np.int32(lloyd_iters)
    r.close()
    # Newton from the Lloyd result (the order remesh_smooth uses, mesh_remesh.cpp:96-106)
    r = fresh(out["x_lloyd"]); r.newton(newton_iters, 7)
    out["x_newton"] = r.points()
    out["newton_iters"] = np.int32(newton_iters)
    r.close()
    r = fresh(out["x_newton"]); r.update_delaunay()
    out["f_after_newton"] = np.float64(r.funcgrad(True)[0])
    r.close()
    r = fresh(out["x_lloyd"]); r.update_delaunay()
    tri, vtx = r.rdt(0)
    out["rdt_tri"] = tri
    r.close()
    # the mode compute_surface uses by default: RDT_MULTINERVE | RDT_RVC_CENTROIDS | RDT_PREFER_SEEDS (CVT.cpp:180-197)
    r = fresh(out["x_lloyd"]); r.update_delaunay()
    out["rdt_mn_tri"], out["rdt_mn_vert"] = r.rdt(7)
    r.close()
    return out


def multinerve_case(V, F, X, lloyd_iters=2):
    """A surface whose cells have several connected components (two sheets closer than the seed spacing)."""
    out = dict(V=V, F=F, X=X)
    r = RefCVT(V, F, multithread=False); r.set_points(X); r.lloyd(lloyd_iters)
    out["x_lloyd"] = r.points()
    r.close()
    for mode in (1, 3, 7):       # multinerve with seeds / with centroids / with centroids and seed preference
        r = RefCVT(V, F, multithread=False); r.set_points(out["x_lloyd"]); r.update_delaunay()
        out["rdt_mn%d_tri" % mode], out["rdt_mn%d_vert" % mode] = r.rdt(mode)
        r.close()
    return out


def sampling_cases():
    """compute_initial_sampling of the reference (single thread: no Hilbert reordering of the mesh) on small meshes."""
    out = {}
    for name, V, E, S, kw in [("noise3d", *shapes.noise_sphere(10), 300, {}),
                              ("sphere6d", shapes.lift_anisotropic(*shapes.icosphere(6), 0.04), shapes.icosphere(6)[1], 200, {}),
                              ("boxw", *shapes.box_surface(6), 250, dict(weights=1.0 + shapes.box_surface(6)[0][:, 0] + 0.5 * shapes.box_surface(6)[0][:, 1])),
                              ("kuhn", *shapes.kuhn_cube(5), 200, dict(volumetric=True))]:
        r = RefCVT(V, E, multithread=False, **kw)
        r.initial_sampling(S)
        out[name + "_V"], out[name + "_E"], out[name + "_x"] = V, E, r.points()
        if "weights" in kw:
            out[name + "_w"] = kw["weights"]
        r.close()
    return out


def volume_case(V, T, X, lloyd_iters=4, newton_iters=3):
    """Volumetric mode (RestrictedVoronoiDiagram::set_volumetric(true)): T holds tetrahedra."""
    out = dict(V=V, F=T, X=X, volumetric=np.int32(1))

    def fresh(x):
        r = RefCVT(V, T, volumetric=True, multithread=False)
        r.set_points(x)
        return r
    r = fresh(X); r.update_delaunay()
    out["knn_idx"], out["knn_cnt"] = r.neighbors(20)
    out["mg"], out["m"], _ = r.centroids(False)
    r.close()
    r = fresh(X); r.update_delaunay()
    out["mg_exact"], out["m_exact"], _ = r.centroids(True)
    r.close()
    r = fresh(X); r.update_delaunay()
    out["f"], out["g"], _ = r.funcgrad(True)
    out["f"] = np.float64(out["f"])
    r.close()
    r = fresh(X); r.lloyd(lloyd_iters)
    out["x_lloyd"] = r.points()
    out["lloyd_iters"] = np.int32(lloyd_iters)
    r.close()
    r = fresh(out["x_lloyd"]); r.newton(newton_iters, 7)
    out["x_newton"] = r.points()
    out["newton_iters"] = np.int32(newton_iters)
    r.close()
    r = fresh(out["x_newton"]); r.update_delaunay()
    out["f_after_newton"] = np.float64(r.funcgrad(True)[0])
    r.close()
    return out


def rdt_volume_case(V, T, X, lloyd_iters=3):
    """compute_RDT in volumetric mode (CentroidalVoronoiTesselation::compute_volume): tets of seed indices, on the raw
    sampling and after Lloyd iterations."""
    out = dict(V=V, F=T, X=X, volumetric=np.int32(1))
    r = RefCVT(V, T, volumetric=True, multithread=False)
    r.set_points(X); r.update_delaunay()
    out["rdt_tets_raw"], _ = r.rdt(0)
    r.lloyd(lloyd_iters)
    out["x_lloyd"] = r.points()
    r.update_delaunay()
    out["rdt_tets_lloyd"], emb = r.rdt(0)
    assert np.array_equal(emb, out["x_lloyd"])
    r.close()
    return out


def main():
    V, F = shapes.icosphere(6)
    np.savez_compressed(os.path.join(HERE, "sphere_s150.npz"), **case(V, F, shapes.sample_surface(V, F, 150, 3)))
    V, F = shapes.box_surface(6)
    w = 1.0 + V[:, 0] + 0.5 * V[:, 1]
    np.savez_compressed(os.path.join(HERE, "box_weighted_s100.npz"), **case(V, F, shapes.sample_surface(V, F, 100, 5), weights=w))
    V, F = shapes.icosphere(6)
    V6 = shapes.lift_anisotropic(V, F, 0.04)
    np.savez_compressed(os.path.join(HERE, "sphere6d_s120.npz"), **case(V6, F, shapes.sample_surface(V6, F, 120, 7)))
    V, F = shapes.trefoil_tube(48, 10)
    np.savez_compressed(os.path.join(HERE, "trefoil_s200.npz"), **case(V, F, shapes.sample_surface(V, F, 200, 9)))
    V, F = shapes.box_surface(8, (1.0, 1.0, 0.02))
    np.savez_compressed(os.path.join(HERE, "thinbox_multinerve_s150.npz"), **multinerve_case(V, F, shapes.sample_surface(V, F, 150, 3)))
    np.savez_compressed(os.path.join(HERE, "sampling.npz"), **sampling_cases())
    V, T = shapes.kuhn_cube(6)
    X = np.random.default_rng(13).random((130, 3))
    np.savez_compressed(os.path.join(HERE, "volume_cube_s130.npz"), **volume_case(V, T, X))
    V, T = shapes.kuhn_cube(6)
    np.savez_compressed(os.path.join(HERE, "rdtvol_cube_s160.npz"), **rdt_volume_case(V, T, 0.02 + 0.96 * np.random.default_rng(17).random((160, 3))))
    for n in sorted(os.listdir(HERE)):
        if n.endswith(".npz"):
            print(n, os.path.getsize(os.path.join(HERE, n)))


if __name__ == "__main__":
    main()
