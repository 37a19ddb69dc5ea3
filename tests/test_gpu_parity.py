"""GPU parity suite (pytest -m gpu): the CUDA path, called through the C-ABI, against the
oracle on the same seeded inputs, against the golden vectors generated from the reference,
and — at BASELINE.json's full sizes — through size-independent properties.

Tolerances (BASELINE.json north_star): kNN lists bit-exact (ties flagged); per-seed mass,
centroid, energy, gradient within 1e-9 relative.

Seeds whose cell is truncated to the 20 stored neighbours in Lloyd mode (flag EXHAUSTED) are
"flagged near-degenerate configurations": there the reference's result depends on its facet
flood-fill order (and on its thread count); parity is asserted on every seed that is neither
flagged nor a kNN neighbour of a flagged seed, and the flagged fraction is bounded.
"""
import glob
import os
import threading

import numpy as np
import pytest

from graphitethree_b200 import capi, shapes, sharding
from oracle import port

pytestmark = pytest.mark.gpu
RTOL = 1e-9
ALL_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
GOLDEN = [p for p in ALL_GOLDEN if not os.path.basename(p).startswith(("volume_", "thinbox_multinerve", "sampling", "rdtvol"))]
GOLDEN_VOLUME = [p for p in ALL_GOLDEN if os.path.basename(p).startswith("volume_")]


def load(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}


def handle_for(V, F, weights=None):
    h = capi.Handle(V.shape[1])
    h.set_mesh(V, F, weights=weights)
    return h


def tainted(V, F, x, weights=None, gpu_flags=None):
    """Seeds whose Lloyd-mode (k = 20 truncated) result is not defined by geometry alone: the flagged seeds
    (neighbour list exhausted before the radius test passed) and every seed whose true restricted cell meets
    a facet that a flagged seed's truncated cell also meets — there the reference's flood-fill may or may not
    reach the pair (generic_RVD.h:1385-1418), depending on its facet order and thread count."""
    eL = port.surface_eval(V, F, x, 0, False, weights=weights, want_pairs=True)
    exh = (eL.flags & port.FLAG_EXHAUSTED).astype(bool)
    if gpu_flags is not None:
        # the two flag sets differ only where the candidate pair sets differ (false candidates on the GPU,
        # flood-fill reach in the reference): both mark truncated cells
        gexh = (gpu_flags & capi.FLAG_EXHAUSTED).astype(bool)
        assert (gexh != exh).mean() <= 0.02, "flag sets differ on %.1f%% of the seeds" % (100 * (gexh != exh).mean())
        exh = exh | gexh
    t = exh.copy()
    if exh.any():
        facets = np.zeros(F.shape[0], dtype=bool)
        facets[eL.pairs[exh[eL.pairs[:, 0]], 1]] = True
        eX = port.surface_eval(V, F, x, 0, True, weights=weights, want_pairs=True)
        t[eX.pairs[facets[eX.pairs[:, 1]], 0]] = True
        t[eL.pairs[facets[eL.pairs[:, 1]], 0]] = True
    return t, eL


def assert_close(a, b, mask=None, what=""):
    if mask is not None:
        a, b = a[mask], b[mask]
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b).max() / scale
    assert err <= RTOL, "%s: relative error %.3e" % (what, err)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_against_reference_golden(built, path):
    G = load(path)
    V, F, X = G["V"], G["F"], G["X"]
    w = G.get("weights")
    h = handle_for(V, F, w)
    h.set_seeds(X)
    idx, cnt, sqd, fl = h.knn(20)
    assert np.array_equal(idx, G["knn_idx"]) and np.array_equal(cnt, G["knn_cnt"])
    # exact cells (check_SR): no exemption at all
    h.set_seeds(X)
    f, g = h.funcgrad(True)
    assert abs(f - float(G["f"])) <= RTOL * abs(float(G["f"]))
    assert_close(g, G["g"], what="gradient")
    assert (h.flags() & (capi.FLAG_EXHAUSTED | capi.FLAG_POLY_OVERFLOW | capi.FLAG_KMAX)).sum() == 0
    if w is None:
        assert_close(h.seed_energy(), G["f_seed"], what="per-seed energy")
    # Lloyd mode (k = 20 truncation): parity away from truncated cells
    h.set_seeds(X)
    mg, m = h.centroids(False)
    t, e = tainted(V, F, X, w, h.flags())
    ok = ~t
    assert ok.sum() > 0
    assert_close(m, G["m"], ok, "mass")
    assert_close(mg, G["mg"], ok, "mass*centroid")
    # random sampling of a coarse mesh: truncated cells are common, yet few seeds actually differ
    assert (np.abs(m - G["m"]) > RTOL * G["m"].max()).mean() <= 0.10
    # iterations from the reference's own Lloyd result (regular sampling)
    x0 = G["x_lloyd"]
    xg = h.lloyd(x0, 3)
    xo, fo = port.lloyd(V, F, x0, 3, weights=w)
    if (fo & port.FLAG_EXHAUSTED).sum() == 0:
        assert np.abs(xg - xo).max() <= 1e-10
    xn, info = h.newton(x0, int(G["newton_iters"]), 7)
    assert info["iters"] == int(G["newton_iters"]) + 1
    assert np.abs(xn - G["x_newton"]).max() <= 1e-8
    h.set_seeds(xn)
    f2, _ = h.funcgrad(True)
    assert abs(f2 - float(G["f_after_newton"])) <= 1e-8 * abs(float(G["f_after_newton"]))
    h.close()


def test_knn_bit_exact_and_edge_cases(built):
    rng = np.random.default_rng(5)
    V, F = shapes.noise_sphere(30)
    for S in (2, 7, 21, 22, 500, 20000):
        X = shapes.sample_surface(V, F, S, S)
        h = capi.Handle(3)
        h.set_seeds(X)                       # kNN needs no mesh
        for k in (20, 5, 33, 70):
            idx, cnt, sqd, fl = h.knn(k)
            pidx, pcnt, psqd, ptie = port.knn(X, k)
            assert np.array_equal(cnt, pcnt)
            clean = (ptie == 0)
            assert np.array_equal(idx[clean], pidx[clean])
            assert np.array_equal(sqd[clean], psqd[clean])           # bit-equal FP64 distances
            assert np.array_equal((fl & capi.FLAG_TIE).astype(bool), ptie.astype(bool))
        q = rng.standard_normal((50, 3))
        assert np.array_equal(h.nearest(q), port.nearest(X, q))
        h.close()
    # duplicated seeds (delaunay_nn.cpp:123-134) and exact ties on a lattice
    X = shapes.sample_surface(V, F, 300, 1)
    X[200] = X[17]
    X[250] = X[17]
    h = capi.Handle(3)
    h.set_seeds(X)
    idx, cnt, sqd, fl = h.knn(20)
    pidx, pcnt, psqd, ptie = port.knn(X, 20)
    assert np.array_equal(cnt, pcnt) and cnt[200] == 0 and cnt[250] == 0
    assert np.array_equal(sqd, psqd)
    g = np.stack(np.meshgrid(np.arange(6.), np.arange(6.), np.arange(6.), indexing="ij"), -1).reshape(-1, 3)
    h.set_seeds(g)
    idx, cnt, sqd, fl = h.knn(20)
    pidx, pcnt, psqd, ptie = port.knn(g, 20)
    assert np.array_equal(sqd, psqd) and (fl & capi.FLAG_TIE).all()   # distances equal, order among ties flagged
    h.close()


@pytest.mark.parametrize("shape,S", [("icosphere", 3000), ("trefoil", 4000), ("cad", 2500), ("noise", 5000)])
def test_per_seed_parity_random_and_regular_sampling(built, shape, S):
    if shape == "icosphere":
        V, F = shapes.icosphere(25)
    elif shape == "trefoil":
        V, F = shapes.trefoil_tube(300, 24)
    elif shape == "cad":
        V, F = shapes.cad_like(14)
    else:
        V, F = shapes.noise_sphere(40)
    X = shapes.sample_surface(V, F, S, 21)
    h = handle_for(V, F)
    for stage, x in (("random", X), ("regular", port.lloyd(V, F, X, 6)[0])):
        h.set_seeds(x)
        idx, cnt, _, _ = h.knn(20)
        h.set_seeds(x)
        mg, m = h.centroids(False)
        fl = h.flags()
        t, e = tainted(V, F, x, None, fl)
        exh = (fl & capi.FLAG_EXHAUSTED).astype(bool)
        ok = ~t
        differ = np.abs(m - e.m) > RTOL * e.m.max()
        if stage == "regular":
            assert exh.mean() < 0.02 and ok.mean() > 0.9 and differ.mean() < 0.005
        else:
            assert differ.mean() < 0.05          # truncation artefacts of the random initial sampling
        assert_close(m, e.m, ok, stage + " mass")
        assert_close(mg, e.mg, ok, stage + " mass*centroid")
        # exact cells: every seed, no exemption; accumulate-into semantics (CVT.cpp:149-150)
        h.set_seeds(x)
        g0 = np.ones((S, 3))
        f, g = h.funcgrad(True, g=g0.copy(), f0=2.0)
        e2 = port.surface_eval(V, F, x, 1, True)
        assert abs((f - 2.0) - e2.f) <= RTOL * e2.f
        assert_close(g - 1.0, e2.g, what=stage + " gradient")
        assert_close(h.seed_energy(), e2.f_seed, what=stage + " per-seed energy")
        assert (h.flags() & (capi.FLAG_EXHAUSTED | capi.FLAG_KMAX | capi.FLAG_POLY_OVERFLOW)).sum() == 0
        h.set_seeds(x)
        mgx, mx = h.centroids(True)
        ex = port.surface_eval(V, F, x, 0, True)
        assert_close(mx, ex.m, what=stage + " exact mass")
        assert_close(mgx, ex.mg, what=stage + " exact mass*centroid")
    h.close()


def test_c1_lloyd_trajectory_stepwise(built):
    """C1: 40 500-triangle icosphere, 10 000 seeds, 30 Lloyd iterations. Each GPU iterate is checked
    against one oracle step from the previous GPU iterate."""
    V, F = shapes.icosphere(45)
    X = shapes.sample_surface(V, F, 10000, 1)
    h = handle_for(V, F)
    traj = [X]
    h.lloyd(X, 30, callback=lambda u, it, f, g: traj.append(h.get_seeds()) or 0)
    assert len(traj) == 31
    for k in (0, 1, 4, 12, 29):
        xo, fo = port.lloyd(V, F, traj[k], 1)
        t, _ = tainted(V, F, traj[k])
        ok = ~t
        assert np.abs(traj[k + 1] - xo)[ok].max() <= 1e-9
        if k >= 4:
            assert ok.mean() > 0.97
    # determinism: same input, same bits
    x1 = h.lloyd(X, 5)
    x2 = h.lloyd(X, 5)
    assert np.array_equal(x1, x2)
    h.close()


def test_newton_trajectory_and_locked_points(built):
    V, F = shapes.icosphere(20)
    X0 = shapes.sample_surface(V, F, 2000, 3)
    X, _ = port.lloyd(V, F, X0, 5)
    h = handle_for(V, F)
    hist = []
    xn, info = h.newton(X, 6, 7, callback=lambda u, it, f, g: hist.append((f, g)) or 0)
    xo, oi = port.newton(V, F, X, 6, 7)
    assert info["iters"] == oi["iters"] == 7 and info["nfev"] == oi["nfev"]
    assert np.abs(np.array([a for a, _ in hist]) - oi["f"]).max() <= 1e-9 * oi["f"][0]
    assert np.abs(np.array([b for _, b in hist]) - oi["gnorm"]).max() <= 1e-7 * oi["gnorm"][0]
    assert np.abs(xn - xo).max() <= 1e-8
    locked = np.zeros(2000, dtype=np.uint8)
    locked[::3] = 1
    xl = h.lloyd(X0, 3, locked=locked)
    assert np.array_equal(xl[::3], X0[::3]) and not np.array_equal(xl[1::3], X0[1::3])
    xo, _ = port.lloyd(V, F, X0, 1, locked=locked)
    x1 = h.lloyd(X0, 1, locked=locked)
    assert np.array_equal(x1[::3], xo[::3])
    xn, info = h.newton(X, 3, 7, locked=locked)
    xo, oi = port.newton(V, F, X, 3, 7, locked=locked)
    assert np.array_equal(xn[::3], X[::3]) and np.abs(xn - xo).max() <= 1e-8
    # cancel from the progress callback (TaskCanceled)
    with pytest.raises(capi.B200CVTError) as ei:
        h.lloyd(X, 10, callback=lambda u, it, f, g: 1 if it == 2 else 0)
    assert ei.value.code == 5
    h.close()


def test_error_behaviour(built):
    with pytest.raises(capi.B200CVTError) as ei:
        capi.Handle(4)
    assert ei.value.code == 1
    h = capi.Handle(3)
    X = np.random.default_rng(0).random((100, 3))
    h.set_seeds(X)
    with pytest.raises(capi.B200CVTError) as ei:
        h.centroids(False)
    assert ei.value.code == 3                       # no mesh
    with pytest.raises(capi.B200CVTError) as ei:
        h.knn(500)
    assert ei.value.code == 1
    V, F = shapes.icosphere(4)
    bad = F.copy()
    bad[0, 0] = 10 ** 6
    with pytest.raises(capi.B200CVTError) as ei:
        h.set_mesh(V, bad)
    assert ei.value.code == 1
    # a failed set_mesh leaves the handle without a mesh; RDT has the same preconditions as an evaluation
    with pytest.raises(capi.B200CVTError) as ei:
        h.rdt()
    assert ei.value.code == 3
    h.set_mesh(V, F)
    tri = h.rdt()                                   # 100 random seeds on a coarse sphere: still a valid call
    assert tri.shape[1] == 3 and (tri[:, 0] < tri[:, 2]).all() and (tri[:, 2] < tri[:, 1]).all()
    h.close()
    hv = capi.Handle(3, volumetric=True)
    Vt, T = shapes.kuhn_cube(3)
    hv.set_mesh(Vt, T)
    hv.set_seeds(X)
    with pytest.raises(capi.B200CVTError) as ei:
        hv.rdt_multinerve()                         # the reference has no multinerve mode for volumes either (RVD.cpp:2309)
    assert ei.value.code == 1
    tets = hv.rdt()                                 # seeds on a sphere, most of them outside the unit cube: still a valid call
    assert tets.shape[1] == 4
    hv.close()


def test_full_size_properties_c2(built):
    """C2 size: 2 M-triangle noise sphere, 200 k seeds. The oracle is too slow here; check
    size-independent properties instead."""
    V, F = shapes.noise_sphere(316)
    S = 200000
    X = shapes.sample_surface(V, F, S, 1)
    P = V[F.astype(np.int64)]
    area = 0.5 * np.linalg.norm(np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]), axis=1)
    h = handle_for(V, F)
    x = h.lloyd(X, 3)
    h.set_seeds(x)
    mg, m = h.centroids(True)
    assert (h.flags() & (capi.FLAG_KMAX | capi.FLAG_POLY_OVERFLOW)).sum() == 0
    assert abs(m.sum() - area.sum()) <= 1e-10 * area.sum()                       # cells tile the surface
    assert np.abs(mg.sum(0) - (area[:, None] * P.mean(1)).sum(0)).max() <= 1e-10 * area.sum()
    h.set_seeds(x)
    f, g = h.funcgrad(True)
    assert np.abs(g - 2.0 * (m[:, None] * x - mg)).max() <= 1e-12 * np.abs(g).max() + 1e-15   # g = 2 m (x - c)
    assert abs(h.seed_energy().sum() - f) <= 1e-12 * f
    # Lloyd step == mg/m of the truncated cells; fixed seeds of a second call give the same bits
    h.set_seeds(x)
    mgl, ml = h.centroids(False)
    x1 = h.lloyd(x, 1)
    assert np.array_equal(x1, (1.0 / ml)[:, None] * mgl)
    # kNN: ascending exact distances, and brute force on a sample
    h.set_seeds(x)
    idx, cnt, sqd, fl = h.knn(20)
    assert (cnt == 20).all() and (np.diff(sqd, axis=1) >= 0).all()
    for i in np.random.default_rng(0).integers(0, S, 64):
        d = x - x[i]
        d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
        o = np.lexsort((np.arange(S), d2))[1:21]
        assert np.array_equal(np.sort(d2[o]), sqd[i])
        if not (fl[i] & capi.FLAG_TIE):
            assert np.array_equal(o.astype(np.uint32), idx[i])
    # energy decreases along Lloyd
    h.set_seeds(X)
    f0, _ = h.funcgrad(True)
    assert f < f0
    h.close()


def test_sharded_lloyd_two_partitions_one_gpu(built):
    """The partition + pack/exchange/unpack path of the C library, with two handles on one GPU
    standing in for two ranks (threads + a barrier play the all-gather)."""
    import torch
    V, F = shapes.icosphere(20)
    X = shapes.sample_surface(V, F, 3001, 5)
    S, dim, world = X.shape[0], 3, 2
    chunk = sharding.chunk_doubles(dim, S, world)
    shared = torch.zeros(chunk * world, dtype=torch.float64, device="cuda")
    barrier = threading.Barrier(world)
    results, errors = [None] * world, []

    def worker(rank):
        try:
            h = handle_for(V, F)
            h.set_partition(rank, world)
            sl = torch.zeros(chunk, dtype=torch.float64, device="cuda")
            al = torch.zeros(chunk * world, dtype=torch.float64, device="cuda")

            def exchange():
                shared[rank * chunk:(rank + 1) * chunk].copy_(sl)
                torch.cuda.synchronize()
                barrier.wait()
                al.copy_(shared)
                torch.cuda.synchronize()
                barrier.wait()
                return 0
            h.set_exchange(sl.data_ptr(), al.data_ptr(), chunk, exchange)
            xd = torch.from_numpy(X).cuda()
            h.set_seeds_device(xd.data_ptr(), S)
            h.lloyd_device(4)
            info = h.newton_device(2, 7)
            results[rank] = (h.get_seeds(), info)
            h.close()
        except Exception as ex:   # pragma: no cover
            errors.append(ex)
            barrier.abort()
    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors
    h = handle_for(V, F)
    x1 = h.lloyd(X, 4)
    x1, info = h.newton(x1, 2, 7)
    h.close()
    assert np.array_equal(results[0][0], results[1][0])
    assert np.array_equal(results[0][0], x1)          # sharding does not change a single bit
    assert results[0][1] == info


# ---------------------------------------------------------------------------------------
# volumetric mode (C5): tetrahedra clipped as convex cells, SURVEY.md §8 rows a7 / a11
# ---------------------------------------------------------------------------------------
def volume_handle(V, T):
    h = capi.Handle(3, volumetric=True)
    h.set_mesh(V, T)
    return h


def untruncated(V, T, x):
    """Seeds whose Lloyd-mode cell (20 stored neighbours, no enlargement) IS their restricted Voronoi cell: the oracle's
    check_SR = false and check_SR = true results agree. Elsewhere the reference's result depends on which (tet, seed)
    pairs its flood fill happens to reach (flagged near-degenerate configurations)."""
    eL = port.surface_eval(V, T, x, 0, False)
    eX = port.surface_eval(V, T, x, 0, True, kcap=min(256, x.shape[0] - 1))
    same = np.abs(eL.m - eX.m) <= 1e-12 * eX.m.max()
    return same, eL, eX


@pytest.mark.parametrize("path", GOLDEN_VOLUME, ids=[os.path.basename(p) for p in GOLDEN_VOLUME])
def test_volume_against_reference_golden(built, path):
    G = load(path)
    V, T, X = G["V"], G["F"], G["X"]
    h = volume_handle(V, T)
    # exact cells: no exemption
    h.set_seeds(X)
    mg, m = h.centroids(True)
    assert_close(m, G["m_exact"], what="volume")
    assert_close(mg, G["mg_exact"], what="volume*centroid")
    assert abs(m.sum() - 1.0) <= 1e-12
    h.set_seeds(X)
    f, g = h.funcgrad(True)
    assert abs(f - float(G["f"])) <= RTOL * abs(float(G["f"]))
    assert_close(g, G["g"], what="gradient")
    assert (h.flags() & (capi.FLAG_EXHAUSTED | capi.FLAG_POLY_OVERFLOW | capi.FLAG_KMAX)).sum() == 0
    e = port.surface_eval(V, T, X, 1, True, kcap=X.shape[0] - 1)
    assert_close(h.seed_energy(), e.f_seed, what="per-seed energy")
    # Lloyd mode: parity wherever the 20-neighbour truncation does not change the cell
    h.set_seeds(X)
    mg, m = h.centroids(False)
    ok, eL, eX = untruncated(V, T, X)
    assert ok.sum() > 0
    assert_close(m, G["m"], ok, "volume (Lloyd mode)")
    assert_close(mg, G["mg"], ok, "volume*centroid (Lloyd mode)")
    # Newton from the reference's Lloyd result
    xn, info = h.newton(G["x_lloyd"], int(G["newton_iters"]), 7)
    assert info["iters"] == int(G["newton_iters"]) + 1
    assert np.abs(xn - G["x_newton"]).max() <= 1e-8
    h.close()


def test_volume_parity_with_oracle_t_over_s_10(built):
    V, T = shapes.kuhn_cube(14)                      # 16 464 tets
    X = np.random.default_rng(21).random((1600, 3))
    x0, _ = port.lloyd(V, T, X, 4)                   # relaxed sampling, as after the first iterations of a run
    h = volume_handle(V, T)
    h.set_seeds(x0)
    f, g = h.funcgrad(True)
    e = port.surface_eval(V, T, x0, 1, True, kcap=256)
    assert abs(f - e.f) <= RTOL * abs(e.f)
    assert_close(g, e.g, what="gradient")
    assert_close(h.seed_energy(), e.f_seed, what="per-seed energy")
    h.set_seeds(x0)
    mg, m = h.centroids(True)
    ex = port.surface_eval(V, T, x0, 0, True, kcap=256)
    assert_close(m, ex.m, what="volume")
    assert_close(mg, ex.mg, what="volume*centroid")
    assert np.abs(g - 2.0 * (m[:, None] * x0 - mg)).max() <= 1e-12      # g = 2 m (x - c) links rows a7 Lloyd and Newton
    # Lloyd mode and Lloyd iterates
    h.set_seeds(x0)
    mgL, mL = h.centroids(False)
    ok, eL, eX = untruncated(V, T, x0)
    assert ok.mean() >= 0.9                          # measured 0.952: the truncated cells sit on the cube's faces
    assert_close(mL, eL.m, ok, "volume (Lloyd mode)")
    assert_close(mgL, eL.mg, ok, "volume*centroid (Lloyd mode)")
    if ok.all():
        xg = h.lloyd(x0, 2)
        xo, _ = port.lloyd(V, T, x0, 2)
        assert np.abs(xg - xo).max() <= 1e-9
    # Newton trajectory against the oracle: same iteration and evaluation counts
    xn, info = h.newton(x0, 4, 7)
    xo, io = port.newton(V, T, x0, 4, 7)
    assert info["iters"] == io["iters"] and info["nfev"] == io["nfev"]
    assert np.abs(xn - xo).max() <= 1e-8
    h.close()


def test_volume_properties_large(built):
    # size-independent properties at a size the oracle does not finish in seconds: 196 608 tets, 20 000 seeds
    V, T = shapes.kuhn_cube(32)
    X = 0.02 + 0.96 * np.random.default_rng(8).random((20000, 3))
    h = volume_handle(V, T)
    x = h.lloyd(X, 3)
    h.set_seeds(x)
    mg, m = h.centroids(True)
    assert (h.flags() & (capi.FLAG_POLY_OVERFLOW | capi.FLAG_KMAX)).sum() == 0
    assert abs(m.sum() - 1.0) <= 1e-11                                   # the cells tile the unit cube
    assert np.abs(mg.sum(0) - 0.5).max() <= 1e-11
    h.set_seeds(x)
    f, g = h.funcgrad(True)
    assert np.abs(g - 2.0 * (m[:, None] * x - mg)).max() <= 1e-12
    assert f > 0 and abs(h.seed_energy().sum() - f) <= 1e-12 * f
    # Lloyd decreases the CVT energy
    x2 = h.lloyd(x, 2)
    h.set_seeds(x2)
    f2, _ = h.funcgrad(True)
    assert f2 < f
    h.close()


# ---------------------------------------------------------------------------------------
# restricted Delaunay triangulation, simple mode: b200cvt_rdt
# ---------------------------------------------------------------------------------------
def rows(tri):
    t = np.asarray(tri, dtype=np.int64).reshape(-1, 3)
    return t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_rdt_against_reference_golden(built, path):
    # the reference's compute_RDT(RDTMode(0)) at its own Lloyd result: identical triangle lists (duplicates included)
    G = load(path)
    h = handle_for(G["V"], G["F"], G.get("weights"))
    h.set_seeds(G["x_lloyd"])
    tri = h.rdt()
    assert (h.flags() & (capi.FLAG_POLY_OVERFLOW | capi.FLAG_KMAX)).sum() == 0
    ref = rows(G["rdt_tri"])
    assert (ref[:, 0] < ref[:, 2]).all() and (ref[:, 2] < ref[:, 1]).all()     # (s, bisector(0), bisector(1)): s < b1 < b0
    got = rows(tri)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    # the call is cached until the seeds change, and leaves the evaluation path usable
    assert np.array_equal(h.rdt(), tri)
    h.set_seeds(G["x_lloyd"])
    mg, m = h.centroids(True)
    assert abs(m.sum() - port.surface_eval(G["V"], G["F"], G["x_lloyd"], 0, True, weights=G.get("weights")).m.sum()) <= 1e-9 * m.sum()
    h.close()


def test_rdt_closed_surface_properties_c1(built):
    # C1 size: after Lloyd the simple-mode RDT of a sphere is a closed genus-0 triangulation of the seeds
    V, F = shapes.icosphere(45)
    X = shapes.sample_surface(V, F, 10000, 1)
    h = handle_for(V, F)
    x = h.lloyd(X, 10)
    h.set_seeds(x)
    tri = np.unique(np.sort(rows(h.rdt()), axis=1), axis=0)
    assert tri.shape[0] == 2 * 10000 - 4                                         # Euler characteristic 2
    e = np.sort(np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [0, 2]]]), axis=1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    assert (cnt == 2).all()                                                      # every edge shared by two triangles
    assert np.unique(tri).size == 10000                                          # every seed is a vertex
    # every RDT edge joins kNN neighbours (the Delaunay graph is a subgraph of the 20-NN graph after Lloyd)
    h.set_seeds(x)
    idx, cnt20, _, _ = h.knn(20)
    nbr = [set(idx[i, :cnt20[i]].tolist()) for i in range(10000)]
    ue = np.unique(e, axis=0)
    assert all(int(b) in nbr[int(a)] for a, b in ue[:: max(1, ue.shape[0] // 2000)])
    h.close()


def test_anisotropic_6d_crease_facets_parity(built):
    # C4-like: CAD-like surface with sharp creases lifted to 6D (normals x 0.04 x bbox diagonal). Facets next to a crease are
    # much longer in 6D than the seed spacing and span many cells: their candidates come from facet_big_kernel (adaptive
    # bisection). Exact cells (check_SR = true) against the oracle, no exemption.
    V3, F = shapes.cad_like(12)
    V = shapes.lift_anisotropic(V3, F, 0.04)
    X = shapes.sample_surface(V, F, 6000, 2)
    h = handle_for(V, F)
    x = h.lloyd(X, 3)
    h.stats()
    h.set_seeds(x)
    f, g = h.funcgrad(True)
    st = h.stats()
    assert st["facets_subdivided"] > 0 and st["facets_subdivision_gave_up"] == 0
    assert (h.flags() & (capi.FLAG_POLY_OVERFLOW | capi.FLAG_KMAX)).sum() == 0
    e = port.surface_eval(V, F, x, 1, True, kcap=256)
    assert abs(f - e.f) <= RTOL * abs(e.f)
    assert_close(g, e.g, what="gradient")
    assert_close(h.seed_energy(), e.f_seed, what="per-seed energy")
    h.set_seeds(x)
    mg, m = h.centroids(True)
    e0 = port.surface_eval(V, F, x, 0, True, kcap=256)
    assert_close(m, e0.m, what="mass")
    assert_close(mg, e0.mg, what="mass*centroid")
    h.close()


def test_rdt_against_oracle_trefoil(built):
    # larger than the golden cases, genus 1: the device RDT against the oracle's restatement (pinned to the reference by
    # tests/test_oracle_cpu.py), on a raw sampling (enlarged neighbourhoods) and after Lloyd
    V, F = shapes.trefoil_tube(300, 30)
    X = shapes.sample_surface(V, F, 3000, 5)
    h = handle_for(V, F)
    for x in (X, h.lloyd(X, 5)):
        h.set_seeds(x)
        got = rows(h.rdt())
        want = rows(port.rdt(V, F, x))
        assert got.shape == want.shape and np.array_equal(got, want)
    t = np.unique(np.sort(got, axis=1), axis=0)
    assert t.shape[0] == 2 * 3000                       # torus: Euler characteristic 0
    h.close()


def test_volume_cell_first_path_equals_tet_path(built, monkeypatch):
    """vcell.cuh builds the Voronoi cell of a seed once (one warp, cooperative) and integrates it directly when the
    inside / boundary grid proves it inside the domain; everything else takes the (tet, seed) path (B200CVT_VCELL=0 forces it
    for all seeds). Same per-seed results on a convex and on a concave domain with seeds outside of it."""
    V, T = shapes.kuhn_cube(20)                                  # 48 000 tets
    cen = V[T].mean(axis=1)
    T_L = np.ascontiguousarray(T[~((cen[:, 0] > 0.5) & (cen[:, 1] > 0.5) & (cen[:, 2] > 0.5))])   # cube minus an octant
    rng = np.random.default_rng(33)
    for Tm, vol in ((T, 1.0), (T_L, 0.875)):
        X = 0.01 + 0.98 * rng.random((Tm.shape[0] // 10, 3))     # on the concave domain some seeds lie outside
        monkeypatch.setenv("B200CVT_VCELL", "0")
        ht = volume_handle(V, Tm)
        monkeypatch.setenv("B200CVT_VCELL", "1")
        hc = volume_handle(V, Tm)
        x = ht.lloyd(X, 2)
        res = {}
        for name, h in (("tet", ht), ("cell", hc)):
            h.stats()
            h.set_seeds(x); mg, m = h.centroids(True); fl = h.flags().copy(); st = h.stats()
            h.set_seeds(x); f, g = h.funcgrad(True); fs = h.seed_energy().copy()
            h.set_seeds(x); mgL, mL = h.centroids(False); flL = h.flags().copy()
            res[name] = (m, mg, f, g, fl, st, mL, mgL, flL, fs)
        a, b = res["tet"], res["cell"]
        S = x.shape[0]
        assert b[5]["volumetric_cells"]["direct"] > 0.4 * S and b[5]["volumetric_cells"]["tet_path"] > 0
        assert a[5]["volumetric_cells"]["direct"] == 0
        assert ((a[4] | b[4]) & (capi.FLAG_EXHAUSTED | capi.FLAG_POLY_OVERFLOW | capi.FLAG_KMAX)).sum() == 0
        assert abs(b[0].sum() - vol) <= 1e-12
        assert np.abs(a[0] - b[0]).max() <= 1e-12 * a[0].max()
        assert np.abs(a[1] - b[1]).max() <= 1e-12 * np.abs(a[1]).max()
        assert abs(a[2] - b[2]) <= 1e-12 * abs(a[2])
        assert np.abs(a[3] - b[3]).max() <= 1e-12 * np.abs(a[3]).max()
        assert np.abs(a[9] - b[9]).max() <= 1e-12 * np.abs(a[9]).max()
        # Lloyd mode (20 stored neighbours): wherever neither path used up a list the cell is exact in both
        ok = ((a[8] | b[8]) & capi.FLAG_EXHAUSTED) == 0
        assert ok.sum() > 0
        assert np.abs(a[6] - b[6])[ok].max() <= 1e-12 * a[6].max()
        assert np.abs(a[7] - b[7])[ok].max() <= 1e-12 * np.abs(a[7]).max()
        # trajectories: exact cells, so the two paths walk the same Newton iterates
        xa, ia = ht.newton(x, 5, 7)
        xb, ib = hc.newton(x, 5, 7)
        assert ia == ib and np.abs(xa - xb).max() <= 1e-9
        ht.close(); hc.close()


def test_volume_cell_with_more_vertices_than_slots_takes_the_tet_path(built, monkeypatch):
    """A seed surrounded by 160 seeds on a sphere: its exact cell has ~160 faces and ~300 vertices, more than the 64 slots of
    vcell_kernel; the kernel gives up on it (and only on it) and the (tet, seed) path integrates it. Same result as the run
    that sends every seed through the tet path, and the cells still tile the cube."""
    V, T = shapes.kuhn_cube(10)
    rng = np.random.default_rng(12)
    u = rng.standard_normal((160, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    ring = 0.5 + 0.3 * u * (1.0 + 1e-3 * rng.random((160, 1)))
    outer = rng.random((4000, 3))
    outer = outer[np.linalg.norm(outer - 0.5, axis=1) > 0.34][:500]
    X = np.concatenate([[[0.5, 0.5, 0.5]], ring, outer])
    res = []
    for env in ("0", "1"):
        monkeypatch.setenv("B200CVT_VCELL", env)
        h = volume_handle(V, T)
        h.stats()
        h.set_seeds(X); mg, m = h.centroids(True); st = h.stats(); fl = h.flags().copy()
        h.set_seeds(X); f, g = h.funcgrad(True)
        h.close()
        res.append((m, mg, f, g, fl, st))
    a, b = res
    assert ((a[4] | b[4]) & (capi.FLAG_EXHAUSTED | capi.FLAG_POLY_OVERFLOW | capi.FLAG_KMAX)).sum() == 0
    assert b[5]["volumetric_cells"]["direct"] > 0 and b[5]["volumetric_cells"]["tet_path"] > 0
    assert abs(b[0].sum() - 1.0) <= 1e-12
    assert 0.012 < b[0][0] < 0.018                         # the big cell: roughly a ball of radius 0.15 (volume 0.0141)
    assert np.abs(a[0] - b[0]).max() <= 1e-12 * a[0].max()
    assert np.abs(a[1] - b[1]).max() <= 1e-12 * np.abs(a[1]).max()
    assert abs(a[2] - b[2]) <= 1e-12 * abs(a[2])
    assert np.abs(a[3] - b[3]).max() <= 1e-12 * np.abs(a[3]).max()


RDTVOL = os.path.join(os.path.dirname(ALL_GOLDEN[0]), "rdtvol_cube_s160.npz")


def lex_rows(a):
    return a[np.lexsort(a.T[::-1])] if len(a) else a


def assert_same_tets(gpu, ref_rows, x):
    """same Delaunay tets (as sets of four seeds), every GPU row positively oriented like the reference's (orient_3d > 0)"""
    assert np.array_equal(lex_rows(np.sort(gpu, axis=1)), lex_rows(np.sort(ref_rows, axis=1)))
    P = x[gpu.astype(np.int64)]
    det = np.linalg.det(np.stack([P[:, 1] - P[:, 0], P[:, 2] - P[:, 0], P[:, 3] - P[:, 0]], axis=1))
    assert (det > 0).all()
    # canonical form: ascending except for the swap of the first two, rows sorted and unique
    srt = np.sort(gpu, axis=1)
    assert np.array_equal(srt, lex_rows(srt)) and len(np.unique(srt, axis=0)) == len(srt)


def test_volume_rdt_against_reference_golden(built, monkeypatch):
    """compute_RDT of a volumetric diagram (CentroidalVoronoiTesselation::compute_volume, RVD.cpp:2308-2335): the Delaunay tets
    whose Voronoi vertex lies inside the domain, from the cell-first path and from the (tet, seed) path alone."""
    G = load(RDTVOL)
    V, T = G["V"], G["F"]
    for env in ("1", "0"):
        monkeypatch.setenv("B200CVT_VCELL", env)
        h = volume_handle(V, T)
        for x, key in ((G["X"], "rdt_tets_raw"), (G["x_lloyd"], "rdt_tets_lloyd")):
            h.set_seeds(x)
            assert_same_tets(h.rdt(), G[key], x)
        # the cached rows follow the seeds: after device-side iterations the tets are those of the moved seeds
        h.set_seeds(G["x_lloyd"])
        h.rdt()
        x = h.newton(G["x_lloyd"], 2, 5)[0]
        assert_same_tets(h.rdt(), port.rdt_volume(V, T, x)[0], x)
        h.close()


def test_volume_rdt_against_oracle_and_delaunay(built):
    """larger case: against the oracle's restatement (pinned to the reference row by row on the CPU) and against the Delaunay
    triangulation of the seeds (scipy / Qhull): the tets whose circumcentre lies in the unit cube"""
    V, T = shapes.kuhn_cube(16)
    X = 0.01 + 0.98 * np.random.default_rng(41).random((3000, 3))
    h = volume_handle(V, T)
    x = h.lloyd(X, 3)
    h.set_seeds(x)
    tets = h.rdt()
    h.close()
    ot, unc = port.rdt_volume(V, T, x)
    assert unc == 0
    assert_same_tets(tets, ot, x)
    from scipy.spatial import Delaunay
    D = Delaunay(x)
    P = x[D.simplices]
    cc = np.linalg.solve(2 * (P[:, 1:] - P[:, :1]), ((P[:, 1:] ** 2).sum(2) - (P[:, :1] ** 2).sum(2))[..., None])[..., 0]
    margin = 1e-9
    surely_in = ((cc > margin) & (cc < 1 - margin)).all(1)
    maybe_in = ((cc > -margin) & (cc < 1 + margin)).all(1)
    mine = set(map(tuple, np.sort(tets, axis=1).tolist()))
    assert set(map(tuple, np.sort(D.simplices[surely_in], axis=1).tolist())) <= mine
    assert mine <= set(map(tuple, np.sort(D.simplices[maybe_in], axis=1).tolist()))


def test_sharded_volumetric_two_partitions_one_gpu(built):
    """Volumetric mode through the partition + exchange path: two handles on one GPU stand in for two ranks."""
    import torch
    V, T = shapes.kuhn_cube(12)
    X = 0.02 + 0.96 * np.random.default_rng(9).random((1001, 3))
    S, dim, world = X.shape[0], 3, 2
    chunk = sharding.chunk_doubles(dim, S, world)
    shared = torch.zeros(chunk * world, dtype=torch.float64, device="cuda")
    barrier = threading.Barrier(world)
    results, errors = [None] * world, []

    def worker(rank):
        try:
            h = volume_handle(V, T)
            h.set_partition(rank, world)
            sl = torch.zeros(chunk, dtype=torch.float64, device="cuda")
            al = torch.zeros(chunk * world, dtype=torch.float64, device="cuda")

            def exchange():
                shared[rank * chunk:(rank + 1) * chunk].copy_(sl)
                torch.cuda.synchronize()
                barrier.wait()
                al.copy_(shared)
                torch.cuda.synchronize()
                barrier.wait()
                return 0
            h.set_exchange(sl.data_ptr(), al.data_ptr(), chunk, exchange)
            xd = torch.from_numpy(X).cuda()
            h.set_seeds_device(xd.data_ptr(), S)
            h.lloyd_device(3)
            info = h.newton_device(2, 7)
            results[rank] = (h.get_seeds(), info)
            h.close()
        except Exception as ex:   # pragma: no cover
            errors.append(ex)
            barrier.abort()
    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors
    h = volume_handle(V, T)
    x1 = h.lloyd(X, 3)
    x1, info = h.newton(x1, 2, 7)
    h.close()
    assert np.array_equal(results[0][0], results[1][0])          # both ranks hold the same seeds
    assert np.abs(results[0][0] - x1).max() <= 1e-12             # and the unsharded run's
    assert results[0][1] == info


# ---------------------------------------------------------------------------------------
# multinerve RDT (the default mode of remesh_smooth): rdt_mn.cuh against the reference's golden rows and the oracle
# ---------------------------------------------------------------------------------------
def assert_same_multinerve(gpu, ref_rows, S, centroids):
    """Order-independent comparison: the reference numbers components and orders / orients rows by its traversal."""
    tg, vg, sg = gpu
    tr, vr, sr = ref_rows
    assert len(vg) == len(vr)
    assert np.array_equal(np.bincount(sg, minlength=S), np.bincount(sr, minlength=S))      # components per seed
    cg, cr = port.canonical_multinerve(tg, vg, sg), port.canonical_multinerve(tr, vr, sr)
    assert np.array_equal(cg[2], cr[2])
    assert np.abs(cg[1] - cr[1]).max() <= 1e-12 * max(1.0, np.abs(cr[1]).max())              # positions of the components
    # components of one seed that share a position (no centroids, locked seeds, border) cannot be told apart by position:
    # triangles are then compared through the seeds of their vertices
    q = np.round(cr[1] * 1e6)
    key = np.concatenate([cr[2][:, None].astype(np.float64), q], 1)
    coincident = len(np.unique(key, axis=0)) != len(key)
    if centroids and not coincident:
        a = np.unique(np.sort(cg[0], axis=1), axis=0)
        b = np.unique(np.sort(cr[0], axis=1), axis=0)
    else:
        a = np.unique(np.sort(sg[tg.astype(np.int64)], axis=1), axis=0)
        b = np.unique(np.sort(sr[tr.astype(np.int64)], axis=1), axis=0)
    assert np.array_equal(a, b)                                                              # identical triangle sets


MN_GOLDEN = [p for p in GOLDEN if "rdt_mn_tri" in np.load(p).files]


@pytest.mark.parametrize("path", MN_GOLDEN, ids=[os.path.basename(p) for p in MN_GOLDEN])
def test_rdt_multinerve_against_reference_golden(built, path):
    d = load(path)
    V, F, x = d["V"], d["F"], d["x_lloyd"]
    # the oracle returns the golden rows bit for bit (tests/test_oracle_cpu.py) and, in addition, the seed of every vertex
    to, vo, so = port.rdt_multinerve(V, F, x, True, True)
    assert np.array_equal(to, d["rdt_mn_tri"]) and np.array_equal(vo, d["rdt_mn_vert"])
    h = capi.Handle(x.shape[1])
    h.set_mesh(V, F)
    h.set_seeds(x)
    got = h.rdt_multinerve(True, True)
    assert (h.flags() & (capi.FLAG_POLY_OVERFLOW | capi.FLAG_KMAX)).sum() == 0
    h.close()
    assert_same_multinerve(got, (to, vo, so), x.shape[0], True)


def test_rdt_multinerve_thin_plate_components(built):
    d = load(os.path.join(os.path.dirname(GOLDEN[0]), "thinbox_multinerve_s150.npz"))
    V, F, x = d["V"], d["F"], d["x_lloyd"]
    S = x.shape[0]
    locked = np.zeros(S, dtype=np.uint8)
    locked[::7] = 1
    for mode, (uc, ps) in {1: (False, False), 3: (True, False), 7: (True, True)}.items():
        to, vo, so = port.rdt_multinerve(V, F, x, uc, ps)
        assert np.array_equal(to, d["rdt_mn%d_tri" % mode]) and np.array_equal(vo, d["rdt_mn%d_vert" % mode])
        h = handle_for(V, F)
        h.set_seeds(x)
        got = h.rdt_multinerve(uc, ps)
        assert (np.bincount(got[2], minlength=S) > 1).sum() > 50        # most cells of the thin plate have two components
        assert_same_multinerve(got, (to, vo, so), S, uc)
        if uc:
            # locked seeds keep their position (RVD.cpp:2199-2203)
            gotl = h.rdt_multinerve(uc, ps, locked=locked)
            assert_same_multinerve(gotl, port.rdt_multinerve(V, F, x, uc, ps, locked=locked), S, True)
            assert np.array_equal(gotl[1][locked[gotl[2]] == 1], x[gotl[2][locked[gotl[2]] == 1]])
        h.close()


def test_rdt_multinerve_against_oracle_raw_and_relaxed(built):
    # a raw sampling (enlarged neighbourhoods, check_SR = true) and a relaxed one; closed genus-0 surface: 2 S - 4 triangles
    V, F = shapes.noise_sphere(60)
    X = shapes.sample_surface(V, F, 8000, 2)
    xl, _ = port.lloyd(V, F, X, 2)
    for x in (X, xl):
        h = handle_for(V, F)
        h.set_seeds(x)
        got = h.rdt_multinerve(True, True)
        h.close()
        assert_same_multinerve(got, port.rdt_multinerve(V, F, x, True, True), 8000, True)
    assert np.unique(np.sort(got[0], axis=1), axis=0).shape[0] == 2 * 8000 - 4


# ---------------------------------------------------------------------------------------
# initial sampling (sampling.cuh): the reference's points bit for bit
# ---------------------------------------------------------------------------------------
def test_initial_sampling_bit_exact(built):
    d = np.load(os.path.join(os.path.dirname(ALL_GOLDEN[0]), "sampling.npz"))
    for name in ("noise3d", "sphere6d", "boxw", "kuhn"):
        V, E, x = d[name + "_V"], d[name + "_E"], d[name + "_x"]
        w = d[name + "_w"] if name + "_w" in d.files else None
        h = capi.Handle(V.shape[1], volumetric=(E.shape[1] == 4))
        h.set_mesh(V, E, weights=w)
        got, ok = h.initial_sampling(x.shape[0])
        assert ok and np.array_equal(got, x), name                # the reference's own output (tests/golden/make_golden.py)
        assert np.array_equal(h.get_seeds(), x)                   # ... and they are the handle's current seeds
        h.close()
    # larger, against the oracle restatement; then straight into the optimisation
    V, F = shapes.noise_sphere(120)
    h = handle_for(V, F)
    got, ok = h.initial_sampling(30000)
    assert ok and np.array_equal(got, port.initial_sampling(V, F, 30000)[0])
    h.lloyd_device(2)
    x1 = h.get_seeds()
    h.close()
    h = handle_for(V, F)
    assert np.array_equal(h.lloyd(got, 2), x1)
    h.close()
    # every sample in one element: the reference reports failure
    V1 = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [5.0, 5.0, 0.0], [5.0 + 1e-9, 5.0, 0.0], [5.0, 5.0 + 1e-9, 0.0]])
    F1 = np.array([[0, 1, 2], [3, 4, 5]], dtype=np.uint32)
    h = handle_for(V1, F1)
    got, ok = h.initial_sampling(5)
    xo, _, oko = port.initial_sampling(V1, F1, 5)
    assert ok == oko and not ok and np.array_equal(got, xo)
    h.close()


# ---------------------------------------------------------------------------------------
# full-size configurations, seed by seed against the UNMODIFIED reference (oracle/_ref, all host threads)
# ---------------------------------------------------------------------------------------
def per_seed_vs_reference(V, E, x, volumetric=False):
    """exact cells (check_SR = true: the reference's result does not depend on its thread count there, only the order of
    its per-seed sums does), masses / centroids / energy / gradient of EVERY seed at 1e-9"""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    r = ref.RefCVT(V, E, volumetric=volumetric, multithread=True)
    try:
        r.set_points(x)
        r.update_delaunay()
        mg_r, m_r, _ = r.centroids(True)
        f_r, g_r, _ = r.funcgrad(True)
    finally:
        r.close()
    h = capi.Handle(x.shape[1], volumetric=volumetric)
    h.set_mesh(V, E)
    h.set_seeds(x)
    mg, m = h.centroids(True)
    fl = h.flags()
    h.set_seeds(x)
    f, g = h.funcgrad(True)
    h.close()
    assert (fl & (capi.FLAG_POLY_OVERFLOW | capi.FLAG_KMAX)).sum() == 0
    assert_close(m, m_r, what="mass")
    assert_close(mg, mg_r, what="mass*centroid")
    assert abs(f - f_r) <= RTOL * abs(f_r)
    assert_close(g, g_r, what="gradient")


def test_full_size_c2_per_seed_against_reference(built):
    V, F = shapes.noise_sphere(316)                                   # 2.0 M triangles
    h = handle_for(V, F)
    x = h.lloyd(shapes.sample_surface(V, F, 200000, 1), 3)
    h.close()
    per_seed_vs_reference(V, F, x)


def test_full_size_c4_6d_per_seed_against_reference(built):
    V3, F = shapes.cad_like(268)                                      # 2.0 M triangles, sharp creases
    V = shapes.lift_anisotropic(V3, F, 0.04)
    h = capi.Handle(6)
    h.set_mesh(V, F)
    x = h.lloyd(shapes.sample_surface(V, F, 500000, 1), 3)
    h.close()
    per_seed_vs_reference(V, F, x)


def test_c5_volumetric_62k_per_seed_against_reference(built):
    V, T = shapes.kuhn_cube(47)                                       # 623 k tetrahedra, T / S = 10 as in C5
    X = 0.01 + 0.98 * np.random.default_rng(5).random((62000, 3))
    h = volume_handle(V, T)
    x = h.lloyd(X, 2)
    h.close()
    per_seed_vs_reference(V, T, x, volumetric=True)


def test_sticky_list_sizes_only_matter_where_flagged(built):
    """The reference keeps the size a neighbour list was enlarged to (Delaunay::update_neighbors stores array_size(v)
    neighbours, delaunay.cpp:260-268, delaunay_nn.cpp:73-86): a Lloyd-mode evaluation AFTER a Newton evaluation clips some cells
    by more than 20 neighbours. The library restarts every evaluation from 20 and enlarges on demand. Provoked
    here on a raw sampling: the oracle run with the sticky sizes differs from the oracle run with fresh sizes on some seeds (it
    computes their exact cells); every such seed lies in the flagged set of the library's result (truncated cells and the cells
    that share a facet with one), and all other seeds agree with the sticky run at 1e-9."""
    V, F = shapes.icosphere(12)
    X = shapes.sample_surface(V, F, 700, 5)
    ks = np.full(X.shape[0], 20, dtype=np.uint32)
    port.surface_eval(V, F, X, 1, True, ksize=ks)                 # a Newton-mode evaluation: ks now holds the enlarged sizes
    assert (ks > 20).sum() > 0
    x2 = X + 1e-4 * np.random.default_rng(1).standard_normal(X.shape)
    e_sticky = port.surface_eval(V, F, x2, 0, False, ksize=ks.copy(), kcap=int(ks.max()))
    e_fresh = port.surface_eval(V, F, x2, 0, False)
    differs = np.abs(e_sticky.m - e_fresh.m) > 1e-12 * e_fresh.m.max()
    assert differs.sum() > 0                                       # the case is provoked (measured: 25 of 700 seeds)
    h = handle_for(V, F)
    h.set_seeds(X); h.funcgrad(True)
    h.set_seeds(x2)
    mg, m = h.centroids(False)
    fl = h.flags()
    h.close()
    t, _ = tainted(V, F, x2, gpu_flags=fl)
    assert np.all(t[differs])                                      # sticky sizes change nothing outside the flagged configurations
    assert np.all(((fl & capi.FLAG_EXHAUSTED) != 0)[(e_fresh.flags & port.FLAG_EXHAUSTED) != 0])
    ok = ~t
    assert ok.sum() > 0.3 * len(ok)
    assert_close(m, e_sticky.m, ok, "mass against the sticky-size run")
    assert_close(mg, e_sticky.mg, ok, "mass*centroid against the sticky-size run")


def test_neighbour_cap_is_flagged(built):
    # Seeds on a line next to a facet much longer than their spacing: every cell is a strip whose security radius reaches all
    # the other seeds, so check_SR asks for all S - 1 neighbours (the reference grows the list without bound,
    # generic_RVD.h:2183-2197). Up to B200CVT_KMAX neighbours the library follows (exact cells, Newton trajectory included);
    # beyond the cap every seed that ran into it is flagged. The strips are bounded by the two adjacent bisectors only, so
    # the values of a single evaluation are still the exact ones.
    V = np.array([[-1.0, -5.0, 0.0], [2.0, -5.0, 0.0], [2.0, 5.0, 0.0], [-1.0, 5.0, 0.0]])
    F = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
    for S in (200, 400):
        x = np.column_stack([(np.arange(S) + 0.5) / S, 1e-3 * np.sin(np.arange(S)), np.zeros(S)])
        h = handle_for(V, F)
        h.set_seeds(x)
        f, g = h.funcgrad(True)
        fl = h.flags()
        e = port.surface_eval(V, F, x, 1, True, kcap=S - 1)
        assert abs(f - e.f) <= RTOL * abs(e.f)
        assert_close(g, e.g, what="gradient")
        assert_close(h.seed_energy(), e.f_seed, what="per-seed energy")
        if S - 1 <= capi.KMAX:
            assert (fl & capi.FLAG_KMAX).sum() == 0                      # 199 neighbours: below the cap, nothing to flag
            x1, info = h.newton(x, 3, 7)
            x1o, infoo = port.newton(V, F, x, 3, 7, kcap=S - 1)
            assert info["iters"] == infoo["iters"] and info["nfev"] == infoo["nfev"] and np.abs(x1 - x1o).max() <= 1e-8
        else:
            assert (fl & capi.FLAG_KMAX).sum() >= S - 2 * capi.KMAX      # flagged: the cap was reached
        h.close()
