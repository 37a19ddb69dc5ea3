"""GPU drop-in test (pytest -m gpu): the reference's own C++ API with the adapter classes of integration/geogram_b200.h
(CentroidalVoronoiTesselationB200 / RestrictedVoronoiDiagramB200 / Delaunay "B200NN") against the stock classes, through
integration/_build/dropin_check (built in the container by __graft_entry__.build(), travels to the GPU box).

BASELINE.json north_star: after N Lloyd iterations the restricted-Delaunay triangle sets must be identical except at flagged
near-degenerate configurations, and the Hausdorff distance to the reference remesh must be within 1e-6 of the bounding-box
diagonal. Both remeshes are extracted by the reference's own compute_RDT / compute_surface from each run's seeds.
"""
import json
import os
import subprocess

import numpy as np
import pytest

from graphitethree_b200 import shapes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "integration", "_build", "dropin_check")


def write_inputs(tmp_path, V, F, X):
    mp, sp = str(tmp_path / "mesh.bin"), str(tmp_path / "seeds.bin")
    with open(mp, "wb") as f:
        f.write(np.array([V.shape[0], F.shape[0], V.shape[1]], dtype=np.uint32).tobytes())
        f.write(np.ascontiguousarray(V, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(F, dtype=np.uint32).tobytes())
    with open(sp, "wb") as f:
        f.write(np.array([X.shape[0], X.shape[1]], dtype=np.uint32).tobytes())
        f.write(np.ascontiguousarray(X, dtype=np.float64).tobytes())
    return mp, sp


def run_check(tmp_path, V, F, X, nl, nn, m=7, pre=1, volumetric=False, sizing_samples=0):
    """pre = stock Lloyd iterations that make the common start: on a raw random sampling the reference's own first
    iteration depends on its thread count (cells that need more than 20 neighbours, check_SR = false; measured 7e-3
    between 1, 3 and 8 threads, 5e-15 afterwards) — the flagged configurations of the parity statement."""
    if not os.path.exists(EXE):
        pytest.fail("integration/_build/dropin_check is missing: run __graft_entry__.build() where /root/reference exists")
    mp, sp = write_inputs(tmp_path, V, F, X)
    out = subprocess.run([EXE, mp, sp, str(nl), str(nn), str(m), str(pre), "1" if volumetric else "0", str(sizing_samples)],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    print(r)
    return r


def test_dropin_lloyd_rdt_identical_c1_like(tmp_path):
    # C1-like: icosphere 20 480 triangles (Graphite's create_sphere at 5 splits), 5 000 seeds, 30 Lloyd iterations
    V, F = shapes.icosphere_split(5)
    X = shapes.sample_surface(V, F, 5000, 1)
    r = run_check(tmp_path, V, F, X, 30, 0)
    assert r["on_gpu"]
    assert r["nn_mismatch"] == 0                         # "B200NN" Delaunay backend == "NN", list by list
    assert r["max_abs_dx_lloyd"] <= 1e-9
    assert r["ref_triangles"] == r["b200_triangles"] and r["only_ref"] == 0 and r["only_b200"] == 0
    assert r["ref_triangles"] == 2 * 5000 - 4            # closed genus-0 remesh: Euler characteristic 2
    tol = 1e-6 * r["bbox_diagonal"]
    assert r["hausdorff_ref_to_b200"] <= tol and r["hausdorff_b200_to_ref"] <= tol


def test_dropin_lloyd_newton_trefoil(tmp_path):
    V, F = shapes.trefoil_tube(400, 40)
    X = shapes.sample_surface(V, F, 4000, 3)
    r = run_check(tmp_path, V, F, X, 5, 10)
    assert r["on_gpu"]
    assert r["max_abs_dx_lloyd"] <= 1e-9
    assert r["max_abs_dx_final"] <= 1e-7                 # 11 L-BFGS iterations of accumulated round-off
    assert r["only_ref"] == 0 and r["only_b200"] == 0
    tol = 1e-6 * r["bbox_diagonal"]
    assert r["hausdorff_ref_to_b200"] <= tol and r["hausdorff_b200_to_ref"] <= tol


def test_dropin_sizing_field(tmp_path):
    """GEO::compute_sizing_field against compute_sizing_field_b200. The device part — the local feature size of
    every mesh vertex = squared distance to its nearest pole, LocalFeatureSize::squared_lfs — is compared on a SHARED set of poles:
    the "weight" attribute is bit-equal. The pole construction itself (parallel 3D Delaunay, delaunay/LFS.cpp) stays on the reference
    and is not reproducible from run to run on degenerate inputs, so the whole function (with a CVT pre-sampling on both sides) is
    run and its deviation reported, not asserted."""
    V, F = shapes.noise_sphere(40)
    X = shapes.sample_surface(V, F, 500, 3)
    r = run_check(tmp_path, V, F, X, 1, 0, sizing_samples=3000)
    assert r["sizing_on_gpu"] and r["sizing_vertices"] == V.shape[0]
    assert r["sizing_max_rel_diff"] == 0.0
    # whole function, pre-sampled on both sides: reported only. weight = lfs^-4 of an unstable quantity (the medial axis of a
    # noisy sphere) behind a pole construction that two runs of the reference itself do not reproduce (measured: 10 081 against
    # 10 086 poles and weights apart by a factor 5 on the same tube mesh, /tmp check during development)
    assert r["sizing_sampled_median_rel_diff"] >= 0.0


def test_dropin_volumetric_lloyd_newton(tmp_path):
    # CentroidalVoronoiTesselation::set_volumetric(true) on a tetrahedralised cube: the adapter's compute_centroids_in_volume /
    # compute_CVT_func_grad_in_volume and the device-resident loops against the stock classes
    V, T = shapes.kuhn_cube(10)                      # 6 000 tets
    X = 0.02 + 0.96 * np.random.default_rng(4).random((600, 3))
    # Newton only (check_SR = true: exact cells). In volumetric Lloyd mode ~5 % of the cells stay truncated to their 20
    # stored neighbours even on a relaxed sampling (tests/test_gpu_parity.py::untruncated), where the reference integrates
    # what its flood fill reaches: flagged configurations, compared seed by seed in the parity suite instead.
    r = run_check(tmp_path, V, T, X, 0, 5, pre=3, volumetric=True)
    assert r["volumetric"] and r["on_gpu"]
    assert r["max_abs_dx_final"] <= 1e-7
    # CentroidalVoronoiTesselation::compute_volume: the adapter's compute_RDT in volumetric mode (b200cvt_rdt on tets) gives
    # the same Delaunay tets as the stock class, all positively oriented, and compute_volume builds the same number of cells
    assert r["volume_tets_ref"] == r["volume_tets_b200"] > 0 and r["only_ref"] == 0 and r["only_b200"] == 0
    assert r["volume_bad_orientation"] == 0
    assert r["compute_volume_cells_ref"] == r["compute_volume_cells_b200"] > 0


def test_dropin_c2_full_size(tmp_path):
    """BASELINE.json configs[1] through the reference's C++ API: 2 M triangles, 200 k seeds, 10 Lloyd + 30 Newton (m = 7), both
    arms from the sampling after two stock Lloyd iterations (no truncated cell is left there). north_star acceptance:
    bit-exact neighbour lists, identical restricted-Delaunay triangle sets, Hausdorff <= 1e-6 of the bounding-box diagonal —
    on the raw RDT, on the simple-mode remesh and on the remesh remesh_smooth produces by default (multinerve + RVC
    centroids). Seeds without any RDT triangle (isolated vertices of BOTH remeshes) are the flagged configurations."""
    import bench
    V, F, X = bench.workload(1, False)
    r = run_check(tmp_path, V, F, X, 10, 30, 7, pre=2)
    assert r["on_gpu"] and r["nn_mismatch"] == 0
    assert r["max_abs_dx_lloyd"] <= 1e-9 and r["max_abs_dx_final"] <= 1e-8
    assert r["only_ref"] == 0 and r["only_b200"] == 0 and r["ref_triangles"] == r["b200_triangles"]
    assert r["isolated_seeds_ref"] == r["isolated_seeds_b200"] <= 10
    tol = 1e-6 * r["bbox_diagonal"]
    for k in ("hausdorff_raw_rdt_ref_to_b200", "hausdorff_raw_rdt_b200_to_ref", "hausdorff_ref_to_b200", "hausdorff_b200_to_ref",
              "hausdorff_multinerve_ref_to_b200", "hausdorff_multinerve_b200_to_ref"):
        assert r[k] <= tol, (k, r[k], tol)
    assert r["multinerve_ref_vertices"] == r["multinerve_b200_vertices"] and r["multinerve_ref_facets"] == r["multinerve_b200_facets"]
    # the whole job through the adapter (mesh upload, Delaunay hand-over included) against the stock classes on this box's cores:
    # reported, not asserted (measured 35x, profiles/r2_dropin_c2.log; one run on a box whose first minute was slow showed 1.8x)
    print("drop-in C2: reference %.2f s, adapter %.3f s" % (r["t_ref_lloyd"] + r["t_ref_newton"], r["t_b200_lloyd"] + r["t_b200_newton"]))
