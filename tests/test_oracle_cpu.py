"""CPU suite: the plain-C oracle against (i) golden vectors generated from the unmodified
reference (tests/golden/make_golden.py), (ii) the reference itself when oracle/_ref is built,
(iii) analytic known answers. No GPU needed."""
import glob
import os

import numpy as np
import pytest

from graphitethree_b200 import shapes
from oracle import port, ref

ALL_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
GOLDEN = [p for p in ALL_GOLDEN if not os.path.basename(p).startswith(("volume_", "thinbox_multinerve", "sampling", "rdtvol"))]
GOLDEN_VOLUME = [p for p in ALL_GOLDEN if os.path.basename(p).startswith("volume_")]


def load(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    G = load(path)
    V, F, X = G["V"], G["F"], G["X"]
    w = G.get("weights")
    idx, cnt, sqd, tie = port.knn(X, 20)
    assert tie.sum() == 0
    assert np.array_equal(idx, G["knn_idx"]) and np.array_equal(cnt, G["knn_cnt"])
    e = port.surface_eval(V, F, X, 0, False, weights=w)
    assert np.array_equal(e.m, G["m"]) and np.array_equal(e.mg, G["mg"])          # bit-exact (same traversal order)
    e = port.surface_eval(V, F, X, 1, True, weights=w, want_pairs=True)
    assert e.f == float(G["f"]) and np.array_equal(e.g, G["g"])
    if w is None:
        assert np.array_equal(e.f_seed, G["f_seed"])
    assert np.array_equal(e.pairs, G["pairs_exact"])
    x, _ = port.lloyd(V, F, X, int(G["lloyd_iters"]), weights=w)
    assert np.array_equal(x, G["x_lloyd"])
    xn, info = port.newton(V, F, G["x_lloyd"], int(G["newton_iters"]), 7, weights=w)
    assert info["iters"] == int(G["newton_iters"]) + 1          # HLBFGS runs max_iter+1 iterations (HLBFGS.cpp:580-584)
    assert np.abs(xn - G["x_newton"]).max() <= 1e-12


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_live_reference_c1_like():
    V, F = shapes.icosphere(20)
    X = shapes.sample_surface(V, F, 2000, 11)
    r = ref.RefCVT(V, F, multithread=False)
    try:
        r.set_points(X)
        r.update_delaunay()
        idx, cnt = r.neighbors(20)
        pidx, pcnt, _, tie = port.knn(X, 20)
        assert np.array_equal(idx, pidx) and np.array_equal(cnt, pcnt) and tie.sum() == 0
        mg, m, _ = r.centroids(False)
        e = port.surface_eval(V, F, X, 0, False)
        assert np.array_equal(m, e.m) and np.array_equal(mg, e.mg)
        q = V[::7]
        assert np.array_equal(port.nearest(X, q), np.array([r.nearest_vertex(p) for p in q], dtype=np.uint32))
    finally:
        r.close()


def test_duplicate_seed_rule():
    # delaunay_nn.cpp:123-134: the later duplicate gets no neighbour, the earlier one skips it
    V, F = shapes.icosphere(4)
    X = shapes.sample_surface(V, F, 60, 2)
    X[40] = X[10]
    idx, cnt, sqd, tie = port.knn(X, 20)
    assert cnt[40] == 0 and cnt[10] == 19 and 40 not in idx[10, :19]
    assert tie[10] or tie[40]


def test_small_seed_count():
    X = np.random.default_rng(0).random((7, 3))
    idx, cnt, _, _ = port.knn(X, 20)
    assert (cnt == 6).all() and (idx[:, 6:] == 0xffffffff).all()


def test_hlbfgs_rosenbrock_known_answer():
    # geogram/src/tests/test_HLBFGS/main.cpp: minimum f=0 at x=1
    x, f, it, nfev = port.hlbfgs_rosenbrock(1000, 5, 1000)
    assert f < 1e-20 and np.abs(x - 1.0).max() < 1e-9


def test_analytic_identities_exact_cells():
    V, F = shapes.icosphere(12)
    X = shapes.sample_surface(V, F, 400, 4)
    P = V[F.astype(np.int64)]
    area = 0.5 * np.linalg.norm(np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]), axis=1)
    ec = port.surface_eval(V, F, X, 0, True)
    assert abs(ec.m.sum() - area.sum()) <= 1e-12 * area.sum()                      # cells tile the surface
    cen = (area[:, None] * P.mean(1)).sum(0)
    assert np.abs(ec.mg.sum(0) - cen).max() <= 1e-12 * area.sum()
    eg = port.surface_eval(V, F, X, 1, True)
    assert np.abs(eg.g - 2.0 * (ec.m[:, None] * X - ec.mg)).max() <= 1e-13         # g = 2 m (x - c)
    # gradient against central finite differences of f
    rng = np.random.default_rng(1)
    d = rng.standard_normal(X.shape)
    h = 1e-6
    fp = port.surface_eval(V, F, X + h * d, 1, True).f
    fm = port.surface_eval(V, F, X - h * d, 1, True).f
    assert abs((fp - fm) / (2 * h) - (eg.g * d).sum()) <= 1e-6 * abs((eg.g * d).sum())


def test_facet_adjacency_closed_surface():
    V, F = shapes.icosphere(5)
    adj = port.facet_adjacency(F)
    assert (adj >= 0).all()
    f = np.arange(F.shape[0])
    for lv in range(3):
        back = adj[adj[:, lv]]
        assert ((back == f[:, None]).sum(1) == 1).all()


# ---------------------------------------------------------------------------------------
# volumetric mode (tetrahedra, GEOGen::ConvexCell restatement)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", GOLDEN_VOLUME, ids=[os.path.basename(p) for p in GOLDEN_VOLUME])
def test_volume_oracle_matches_reference_golden(path):
    G = load(path)
    V, T, X = G["V"], G["F"], G["X"]
    assert T.shape[1] == 4
    e = port.surface_eval(V, T, X, 0, False)
    assert np.array_equal(e.m, G["m"]) and np.array_equal(e.mg, G["mg"])          # bit-exact (same traversal order)
    e = port.surface_eval(V, T, X, 0, True, kcap=129)
    assert np.array_equal(e.m, G["m_exact"]) and np.array_equal(e.mg, G["mg_exact"])
    e = port.surface_eval(V, T, X, 1, True, kcap=129)
    assert e.f == float(G["f"]) and np.array_equal(e.g, G["g"])
    x, _ = port.lloyd(V, T, X, int(G["lloyd_iters"]))
    assert np.array_equal(x, G["x_lloyd"])
    xn, info = port.newton(V, T, G["x_lloyd"], int(G["newton_iters"]), 7, kcap=129)
    assert info["iters"] == int(G["newton_iters"]) + 1
    assert np.abs(xn - G["x_newton"]).max() <= 1e-12


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_volume_oracle_matches_live_reference():
    V, T = shapes.kuhn_cube(10)
    X = np.random.default_rng(3).random((600, 3))
    r = ref.RefCVT(V, T, volumetric=True, multithread=False)
    try:
        r.set_points(X)
        r.update_delaunay()
        mg, m, _ = r.centroids(False)
        e = port.surface_eval(V, T, X, 0, False)
        assert np.array_equal(m, e.m) and np.array_equal(mg, e.mg)
        f, g, _ = r.funcgrad(True)
        e = port.surface_eval(V, T, X, 1, True, kcap=256)
        assert f == e.f and np.array_equal(g, e.g)
    finally:
        r.close()


RDTVOL = os.path.join(os.path.dirname(ALL_GOLDEN[0]), "rdtvol_cube_s160.npz")


def exact_orient3d(p):
    """sign of det(p1 - p0, p2 - p0, p3 - p0) in rational arithmetic (the reference: PCK::orient_3d)"""
    from fractions import Fraction
    q = [[Fraction(float(c)) for c in row] for row in p]
    a, b, c = ([q[i][k] - q[0][k] for k in range(3)] for i in (1, 2, 3))
    d = a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0])
    return (d > 0) - (d < 0)


def test_volume_rdt_oracle_matches_reference_golden():
    """compute_RDT in volumetric mode (RVD.cpp:2308-2335): the oracle returns the reference's rows in the reference's order."""
    G = load(RDTVOL)
    V, T = G["V"], G["F"]
    for x, key in ((G["X"], "rdt_tets_raw"), (G["x_lloyd"], "rdt_tets_lloyd")):
        tet, unc = port.rdt_volume(V, T, x)
        assert unc == 0 and np.array_equal(tet, G[key])
        assert all(exact_orient3d(x[row]) > 0 for row in tet[::7])
        # independent check: the Delaunay tets whose circumcentre lies in the cube
        from scipy.spatial import Delaunay
        D = Delaunay(x)
        P = x[D.simplices]
        cc = np.linalg.solve(2 * (P[:, 1:] - P[:, :1]), ((P[:, 1:] ** 2).sum(2) - (P[:, :1] ** 2).sum(2))[..., None])[..., 0]
        inside = ((cc > 0) & (cc < 1)).all(1)
        lex = lambda a: a[np.lexsort(a.T[::-1])]
        assert np.array_equal(lex(np.sort(D.simplices[inside], axis=1).astype(np.uint32)), lex(np.sort(tet, axis=1)))


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_volume_rdt_oracle_matches_live_reference():
    V, T = shapes.kuhn_cube(8)
    X = 0.02 + 0.96 * np.random.default_rng(23).random((400, 3))
    x, _ = port.lloyd(V, T, X, 2)
    r = ref.RefCVT(V, T, volumetric=True, multithread=False)
    try:
        r.set_points(x)
        r.update_delaunay()
        rt, emb = r.rdt(0)
    finally:
        r.close()
    tet, unc = port.rdt_volume(V, T, x)
    assert np.array_equal(tet, rt) and np.array_equal(emb, x)


def test_volume_analytic_identities_exact_cells():
    V, T = shapes.kuhn_cube(7)
    X = 0.05 + 0.9 * np.random.default_rng(2).random((200, 3))
    ec = port.surface_eval(V, T, X, 0, True, kcap=199)
    assert abs(ec.m.sum() - 1.0) <= 1e-12                                          # cells tile the unit cube
    assert np.abs(ec.mg.sum(0) - 0.5).max() <= 1e-12
    eg = port.surface_eval(V, T, X, 1, True, kcap=199)
    assert np.abs(eg.g - 2.0 * (ec.m[:, None] * X - ec.mg)).max() <= 1e-13         # g = 2 m (x - c)
    assert abs(eg.f_seed.sum() - eg.f) <= 1e-15
    rng = np.random.default_rng(1)
    d = rng.standard_normal(X.shape)
    h = 1e-6
    fp = port.surface_eval(V, T, X + h * d, 1, True, kcap=199).f
    fm = port.surface_eval(V, T, X - h * d, 1, True, kcap=199).f
    assert abs((fp - fm) / (2 * h) - (eg.g * d).sum()) <= 1e-5 * abs((eg.g * d).sum())


def test_tet_adjacency_kuhn_cube():
    V, T = shapes.kuhn_cube(3)
    adj = port.tet_adjacency(T)
    # 6 n^3 tets, 4 faces each; the cube's boundary carries 2 * 6 n^2 triangles
    assert (adj < 0).sum() == 12 * 9
    t = np.arange(T.shape[0])
    for lf in range(4):
        ok = adj[:, lf] >= 0
        back = adj[adj[ok, lf]]
        assert ((back == t[ok, None]).sum(1) == 1).all()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_rdt_matches_reference_golden(path):
    # compute_RDT, simple mode, at the reference's own Lloyd result: same triangles in the same (traversal) order
    G = load(path)
    tri = port.rdt(G["V"], G["F"], G["x_lloyd"])
    assert np.array_equal(tri, G["rdt_tri"])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_rdt_matches_live_reference():
    # a raw random sampling (cells that need enlarged neighbourhoods) and a relaxed one, C1-like size
    V, F = shapes.icosphere(20)
    X = shapes.sample_surface(V, F, 2000, 11)
    xl, _ = port.lloyd(V, F, X, 3)
    for x in (X, xl):
        r = ref.RefCVT(V, F, multithread=False)
        try:
            r.set_points(x)
            r.update_delaunay()
            tri, _ = r.rdt(0)
        finally:
            r.close()
        assert np.array_equal(port.rdt(V, F, x), tri)
    # closed genus-0 surface after Lloyd: 2 S - 4 distinct triangles
    t = np.unique(np.sort(port.rdt(V, F, xl).astype(np.int64), axis=1), axis=0)
    assert t.shape[0] == 2 * 2000 - 4


# ---------------------------------------------------------------------------------------
# multinerve RDT (RVD.cpp:1901-2264): the restatement returns the reference's own rows and vertices
# ---------------------------------------------------------------------------------------
MN_GOLDEN = [p for p in GOLDEN if "rdt_mn_tri" in np.load(p).files]


@pytest.mark.parametrize("path", MN_GOLDEN, ids=[os.path.basename(p) for p in MN_GOLDEN])
def test_oracle_multinerve_matches_reference_golden(path):
    d = load(path)
    tri, vert, vseed = port.rdt_multinerve(d["V"], d["F"], d["x_lloyd"], True, True)
    assert np.array_equal(tri, d["rdt_mn_tri"])            # same rows in the reference's traversal order
    assert np.array_equal(vert, d["rdt_mn_vert"])          # same vertices bit for bit, in order of discovery


def test_oracle_multinerve_thin_plate_golden():
    # two sheets closer than the seed spacing: most cells have two connected components
    d = load(os.path.join(os.path.dirname(GOLDEN[0]), "thinbox_multinerve_s150.npz"))
    for mode, (uc, ps) in {1: (False, False), 3: (True, False), 7: (True, True)}.items():
        tri, vert, vseed = port.rdt_multinerve(d["V"], d["F"], d["x_lloyd"], uc, ps)
        assert np.array_equal(tri, d["rdt_mn%d_tri" % mode])
        assert np.array_equal(vert, d["rdt_mn%d_vert" % mode])
        assert (np.bincount(vseed, minlength=d["X"].shape[0]) > 1).sum() > 50


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_multinerve_matches_live_reference():
    V, F = shapes.trefoil_tube(120, 16)
    X = shapes.sample_surface(V, F, 700, 4)
    xl, _ = port.lloyd(V, F, X, 2)
    for x in (X, xl):
        for mode, (uc, ps) in {1: (False, False), 7: (True, True)}.items():
            r = ref.RefCVT(V, F, multithread=False)
            try:
                r.set_points(x)
                r.update_delaunay()
                tri, vert = r.rdt(mode)
            finally:
                r.close()
            to, vo, _ = port.rdt_multinerve(V, F, x, uc, ps)
            assert np.array_equal(to, tri) and np.array_equal(vo, vert)


# ---------------------------------------------------------------------------------------
# initial sampling (RVD.cpp:1658-1698, mesh_sampling.h): the restatement returns the reference's points bit for bit
# ---------------------------------------------------------------------------------------
def sampling_golden():
    d = np.load(os.path.join(os.path.dirname(ALL_GOLDEN[0]), "sampling.npz"))
    for name in ("noise3d", "sphere6d", "boxw", "kuhn"):
        yield name, d[name + "_V"], d[name + "_E"], d[name + "_x"], (d[name + "_w"] if name + "_w" in d.files else None)


def test_oracle_initial_sampling_matches_reference_golden():
    for name, V, E, x, w in sampling_golden():
        xo, elem, ok = port.initial_sampling(V, E, x.shape[0], weights=w)
        assert ok and np.array_equal(xo, x), name
        assert np.all(np.diff(elem.astype(np.int64)) >= 0)       # sorted uniforms walk the elements in order


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_initial_sampling_matches_live_reference():
    for V, E, S, kw in [(*shapes.noise_sphere(40), 20000, {}), (*shapes.trefoil_tube(300, 24), 5000, {}), (*shapes.kuhn_cube(6), 2000, dict(volumetric=True))]:
        r = ref.RefCVT(V, E, multithread=False, **kw)
        try:
            r.initial_sampling(S)
            xr = r.points()
        finally:
            r.close()
        assert np.array_equal(port.initial_sampling(V, E, S)[0], xr)
