/*
 * dropin_check.cpp — drives the SAME remeshing job twice through the reference's own C++ API:
 * once with the stock GEO::CentroidalVoronoiTesselation (CPU) and once with
 * GEO::CentroidalVoronoiTesselationB200 (geogram_b200.h), then compares what BASELINE.json's north_star
 * asks for: seeds after the Lloyd / Newton iterations, the restricted-Delaunay triangle sets
 * (CentroidalVoronoiTesselation::compute_surface, i.e. RestrictedVoronoiDiagram::compute_RDT on each run's seeds)
 * and the two-sided Hausdorff distance between the two remeshes relative to the bounding-box diagonal.
 *
 * usage: dropin_check mesh.bin seeds.bin nb_Lloyd nb_Newton m [rvd_only]
 *   mesh.bin : u32 nv, u32 nf, u32 dim, f64 vertices[nv*dim], u32 triangles[nf*3]
 *   seeds.bin: u32 S, u32 dim, f64 x[S*dim]
 * prints one JSON line. Used by tests/test_gpu_dropin.py (GPU box) — test infrastructure around the adapter.
 */
#include "geogram_b200.h"

#include <geogram/basic/command_line.h>
#include <geogram/basic/command_line_args.h>
#include <geogram/basic/logger.h>
#include <geogram/basic/process.h>
#include <geogram/basic/stopwatch.h>
#include <geogram/mesh/mesh_distance.h>
#include <geogram/mesh/mesh_geometry.h>
#include <geogram/mesh/mesh_repair.h>
#include <geogram/delaunay/LFS.h>
#include <geogram/numerics/predicates.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <type_traits>
#include <vector>

using namespace GEO;

namespace {

    bool read_all(const char* path, std::vector<unsigned char>& buf) {
        FILE* f = fopen(path, "rb");
        if(f == nullptr) {
            return false;
        }
        fseek(f, 0, SEEK_END);
        long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        buf.resize(size_t(n));
        bool ok = fread(buf.data(), 1, size_t(n), f) == size_t(n);
        fclose(f);
        return ok;
    }

    bool g_volumetric = false;   /* the element table holds tetrahedra; CVT in volumetric mode */

    void load_mesh(const std::vector<unsigned char>& buf, Mesh& M) {
        const uint32_t* hdr = reinterpret_cast<const uint32_t*>(buf.data());
        uint32_t nv = hdr[0], nf = hdr[1], dim = hdr[2];
        const double* V = reinterpret_cast<const double*>(buf.data() + 12);
        const uint32_t* T = reinterpret_cast<const uint32_t*>(buf.data() + 12 + sizeof(double) * size_t(nv) * dim);
        M.clear();
        M.vertices.set_dimension(dim);
        M.vertices.create_vertices(nv);
        for(uint32_t v = 0; v < nv; ++v) {
            for(uint32_t c = 0; c < dim; ++c) {
                M.vertices.point_ptr(v)[c] = V[size_t(v) * dim + c];
            }
        }
        if(g_volumetric) {
            M.cells.create_tets(nf);
            for(uint32_t t = 0; t < nf; ++t) {
                for(uint32_t lv = 0; lv < 4; ++lv) {
                    M.cells.set_vertex(t, lv, T[size_t(t) * 4 + lv]);
                }
            }
            M.cells.connect();
            M.cells.compute_borders();
            return;
        }
        M.facets.create_triangles(nf);
        for(uint32_t f = 0; f < nf; ++f) {
            for(uint32_t lv = 0; lv < 3; ++lv) {
                M.facets.set_vertex(f, lv, T[size_t(f) * 3 + lv]);
            }
        }
        M.facets.connect();
    }

    typedef std::array<index_t, 3> Tri;

    /* triangles of RestrictedVoronoiDiagram::compute_RDT in simple mode: indices ARE seed indices (RVD.cpp:2352-2370) */
    std::set<Tri> triangle_set(const vector<index_t>& tris) {
        std::set<Tri> s;
        for(index_t f = 0; f + 2 < tris.size(); f += 3) {
            Tri t = {{tris[f], tris[f + 1], tris[f + 2]}};
            /* rotate so that the smallest vertex comes first (orientation kept) */
            index_t k = index_t(std::min_element(t.begin(), t.end()) - t.begin());
            Tri r = {{t[k], t[(k + 1) % 3], t[(k + 2) % 3]}};
            s.insert(r);
        }
        return s;
    }

    /* distinct triangles as unordered vertex triples */
    std::set<Tri> ta_all(const vector<index_t>& tris) {
        std::set<Tri> s;
        for(index_t f = 0; f + 2 < tris.size(); f += 3) {
            Tri t = {{tris[f], tris[f + 1], tris[f + 2]}};
            std::sort(t.begin(), t.end());
            s.insert(t);
        }
        return s;
    }

    struct Run {
        std::vector<double> x_lloyd, x_final;
        Mesh surface;          /* compute_surface(simple mode, seeds as vertices) */
        Mesh surface_mn;       /* compute_surface as remesh_smooth calls it: multinerve + RVC centroids (CVT.cpp:180-229) */
        vector<index_t> rdt;
        double t_lloyd = 0.0, t_newton = 0.0;
        bool on_gpu = false;
        index_t nb_volume_tets = 0;
    };

    template <class CVT_T>
    void run(Mesh& M, index_t S, index_t dim, const double* seeds, index_t nl, index_t nn, index_t m, Run& out) {
        /* the reference with its kd-tree backend; the adapter with its default ("default" -> algo:delaunay = NN -> B200NN) */
        CVT_T cvt(&M, coord_index_t(dim), std::is_same<CVT_T, CentroidalVoronoiTesselationB200>::value ? "default" : "NN");
        cvt.set_volumetric(g_volumetric);
        cvt.set_points(S, seeds);
        double t0 = Stopwatch::now();
        cvt.Lloyd_iterations(nl);
        out.t_lloyd = Stopwatch::now() - t0;
        out.x_lloyd.assign(cvt.embedding(0), cvt.embedding(0) + size_t(S) * dim);
        if(nn > 0) {
            t0 = Stopwatch::now();
            cvt.Newton_iterations(nn, m);
            out.t_newton = Stopwatch::now() - t0;
        }
        out.x_final.assign(cvt.embedding(0), cvt.embedding(0) + size_t(S) * dim);
        if constexpr (std::is_same<CVT_T, CentroidalVoronoiTesselationB200>::value) {
            out.on_gpu = cvt.last_call_on_gpu();
        }
        if(g_volumetric) {
            /* CentroidalVoronoiTesselation::compute_volume (CVT.cpp:234-270): set_vertices, check_SR = true, compute_RDT —
             * the Delaunay tets whose Voronoi vertex lies inside the domain */
            cvt.RVD()->delete_threads();
            cvt.compute_volume(&out.surface);
            out.nb_volume_tets = out.surface.cells.nb();
            vector<double> emb;
            cvt.RVD()->compute_RDT(out.rdt, emb, RestrictedVoronoiDiagram::RDTMode(0));
            if constexpr (std::is_same<CVT_T, CentroidalVoronoiTesselationB200>::value) {
                out.on_gpu = cvt.last_call_on_gpu();
            }
            return;
        }
        cvt.RVD()->delete_threads();
        cvt.set_use_RVC_centroids(false);   /* vertices of the remesh = the seeds, so that the triangle sets are comparable */
        cvt.compute_surface(&out.surface, false);
        vector<double> emb;
        cvt.RVD()->compute_RDT(out.rdt, emb, RestrictedVoronoiDiagram::RDTMode(0));
        cvt.set_use_RVC_centroids(true);
        cvt.compute_surface(&out.surface_mn, true);
        if constexpr (std::is_same<CVT_T, CentroidalVoronoiTesselationB200>::value) {
            out.on_gpu = cvt.last_call_on_gpu();
        }
    }

    double max_abs_diff(const std::vector<double>& a, const std::vector<double>& b) {
        double d = 0.0;
        for(size_t i = 0; i < a.size(); ++i) {
            d = std::max(d, std::fabs(a[i] - b[i]));
        }
        return d;
    }
}

int main(int argc, char** argv) {
    if(argc < 6) {
        fprintf(stderr, "usage: %s mesh.bin seeds.bin nb_Lloyd nb_Newton m [nb_pre_Lloyd] [volumetric] [sizing_samples]\n", argv[0]);
        return 2;
    }
    GEO::initialize(GEO::GEOGRAM_INSTALL_NONE);
    CmdLine::import_arg_group("standard");
    CmdLine::import_arg_group("algo");
    CmdLine::import_arg_group("opt");
    CmdLine::import_arg_group("remesh");
    CmdLine::set_arg("log:quiet", "true");
    Logger::instance()->set_quiet(true);
    b200_register();

    std::vector<unsigned char> mb, sb;
    if(!read_all(argv[1], mb) || !read_all(argv[2], sb)) {
        fprintf(stderr, "cannot read inputs\n");
        return 2;
    }
    const index_t nl = index_t(atoi(argv[3])), nn = index_t(atoi(argv[4])), m = index_t(atoi(argv[5]));
    const uint32_t* sh = reinterpret_cast<const uint32_t*>(sb.data());
    const index_t S = sh[0], dim = sh[1];
    const double* seeds = reinterpret_cast<const double*>(sb.data() + 8);

    /* Common start = the given sampling after nb_pre stock Lloyd iterations. On a raw random sampling some cells need more
     * than the 20 stored neighbours; with check_SR = false (Lloyd mode) the reference then integrates whatever its facet
     * flood fill reaches, which depends on its own thread count (measured: 7e-3 between 1, 3 and 8 threads on the first
     * iteration of the C1-like case, 5e-15 from the second on). Those are the "flagged configurations" of the parity
     * statement; one stock iteration removes them. */
    const index_t npre = (argc > 6) ? index_t(atoi(argv[6])) : 0;
    g_volumetric = (argc > 7) && atoi(argv[7]) != 0;
    std::vector<double> start(seeds, seeds + size_t(S) * dim);
    if(npre > 0) {
        Mesh M;
        load_mesh(mb, M);
        CentroidalVoronoiTesselation pre(&M, coord_index_t(dim), "NN");
        pre.set_volumetric(g_volumetric);
        pre.set_points(S, seeds);
        pre.Lloyd_iterations(npre);
        start.assign(pre.embedding(0), pre.embedding(0) + size_t(S) * dim);
    }
    seeds = start.data();

    /* optional (argv[8] != 0, surfaces in dimension 3): the sizing field, GEO::compute_sizing_field against
     * compute_sizing_field_b200 — without pre-sampling (bit-equal weights expected: same poles, nearest pole by the device) and
     * with a CVT pre-sampling of argv[8] points (the samplings differ by the optimiser's round-off, so do the poles) */
    const index_t sizing_samples = (argc > 8) ? index_t(atoi(argv[8])) : 0;
    double sizing_max_rel = -1.0, sizing_sampled_median_rel = -1.0;
    index_t sizing_vertices = 0;
    bool sizing_on_gpu = false;
    if(sizing_samples != 0 && !g_volumetric && dim == 3) {
        /* (1) the device part on a SHARED set of poles: weights from LFS.squared_lfs (the reference's loop,
         * mesh_geometry.cpp:57-73) against compute_sizing_field_lfs_b200 — bit-equal expected */
        {
            Mesh M;
            load_mesh(mb, M);
            LocalFeatureSize LFS(M.vertices.nb(), M.vertices.point_ptr(0));
            double min_distance2 = 0.1 * surface_average_edge_length(M);
            min_distance2 = min_distance2 * min_distance2;
            std::vector<double> wr(M.vertices.nb());
            for(index_t v = 0; v < M.vertices.nb(); ++v) {
                double lfs2 = std::max(LFS.squared_lfs(M.vertices.point_ptr(v)), min_distance2);
                wr[v] = pow(lfs2, -2.0);
            }
            sizing_on_gpu = compute_sizing_field_lfs_b200(M, LFS, 1.0);
            Attribute<double> w(M.vertices.attributes(), "weight");
            sizing_vertices = M.vertices.nb();
            sizing_max_rel = 0.0;
            for(index_t v = 0; v < M.vertices.nb(); ++v) {
                sizing_max_rel = std::max(sizing_max_rel, std::fabs(wr[v] - w[v]) / std::fabs(wr[v]));
            }
        }
        /* (2) the whole function with a CVT pre-sampling on both sides: different samples (device loop against CPU loop) and the
         * reference's pole construction is not reproducible run to run, so only the distribution is compared */
        auto weights = [&](bool gpu, index_t samples, std::vector<double>& out) {
            Mesh M;
            load_mesh(mb, M);
            if(gpu) {
                compute_sizing_field_b200(M, 1.0, samples);
            } else {
                compute_sizing_field(M, 1.0, samples);
            }
            Attribute<double> w(M.vertices.attributes(), "weight");
            out.resize(M.vertices.nb());
            for(index_t v = 0; v < M.vertices.nb(); ++v) {
                out[v] = w[v];
            }
        };
        std::vector<double> wr, wg;
        weights(false, sizing_samples, wr);
        weights(true, sizing_samples, wg);
        std::vector<double> rel(wr.size());
        for(size_t v = 0; v < wr.size(); ++v) {
            rel[v] = std::fabs(wr[v] - wg[v]) / std::fabs(wr[v]);
        }
        std::nth_element(rel.begin(), rel.begin() + rel.size() / 2, rel.end());
        sizing_sampled_median_rel = rel.empty() ? 0.0 : rel[rel.size() / 2];
    }

    Run ref, b200;
    {
        Mesh M;
        load_mesh(mb, M);
        run<CentroidalVoronoiTesselation>(M, S, dim, seeds, nl, nn, m, ref);
    }
    {
        Mesh M;
        load_mesh(mb, M);
        try {
            run<CentroidalVoronoiTesselationB200>(M, S, dim, seeds, nl, nn, m, b200);
        } catch(const std::exception& e) {
            printf("{\"error\": \"%s\"}\n", e.what());
            return 1;
        }
    }

    /* the "B200NN" Delaunay backend through the reference factory, compared list by list with "NN" */
    index_t nn_mismatch = 0, nn_rows = 0;
    {
        Delaunay_var d_ref = Delaunay::create(coord_index_t(dim), "NN");
        Delaunay_var d_gpu = Delaunay::create(coord_index_t(dim), "B200NN");
        d_ref->set_stores_neighbors(true);
        d_gpu->set_stores_neighbors(true);
        d_ref->set_vertices(S, ref.x_lloyd.data());
        d_gpu->set_vertices(S, ref.x_lloyd.data());
        vector<index_t> a, b;
        for(index_t i = 0; i < S; ++i) {
            d_ref->get_neighbors(i, a);
            d_gpu->get_neighbors(i, b);
            ++nn_rows;
            if(a.size() != b.size() || !std::equal(a.begin(), a.end(), b.begin())) {
                ++nn_mismatch;
            }
        }
    }

    std::set<Tri> ta, tb;
    size_t only_ref = 0, only_b200 = 0, vol_tets_ref = 0, vol_tets_b200 = 0, vol_bad_orientation = 0;
    if(g_volumetric) {
        /* tets as sorted quadruples; every row of both sides positively oriented (RVD.cpp:2318-2329) */
        typedef std::array<index_t, 4> Tet;
        auto tet_set = [&](const vector<index_t>& rows, const std::vector<double>& x) {
            std::set<Tet> out;
            for(index_t f = 0; f + 3 < rows.size(); f += 4) {
                Tet t = {{rows[f], rows[f + 1], rows[f + 2], rows[f + 3]}};
                if(PCK::orient_3d(&x[size_t(t[0]) * 3], &x[size_t(t[1]) * 3], &x[size_t(t[2]) * 3], &x[size_t(t[3]) * 3]) <= 0) {
                    ++vol_bad_orientation;
                }
                std::sort(t.begin(), t.end());
                out.insert(t);
            }
            return out;
        };
        std::set<Tet> va = tet_set(ref.rdt, ref.x_final), vb = tet_set(b200.rdt, ref.x_final);
        vol_tets_ref = va.size(); vol_tets_b200 = vb.size();
        for(const Tet& t : va) { if(vb.find(t) == vb.end()) { ++only_ref; } }
        for(const Tet& t : vb) { if(va.find(t) == va.end()) { ++only_b200; } }
    } else {
        ta = triangle_set(ref.rdt); tb = triangle_set(b200.rdt);
    }
    for(const Tri& t : ta) {
        if(tb.find(t) == tb.end()) {
            ++only_ref;
        }
    }
    for(const Tri& t : tb) {
        if(ta.find(t) == ta.end()) {
            ++only_b200;
        }
    }
    double diag = (ref.surface.vertices.nb() > 0) ? bbox_diagonal(ref.surface) : 0.0;
    double h_ab = 0.0, h_ba = 0.0, h_ctrl = 0.0, h_raw_ab = 0.0, h_raw_ba = 0.0, h_mn_ab = 0.0, h_mn_ba = 0.0;
    size_t nonmanifold_edges = 0;
    /* A seed without any restricted Delaunay triangle stays in the remesh as an isolated vertex (assign_triangle_mesh keeps
     * all seeds, mesh_repair does not remove them), and mesh_one_sided_Hausdorff_distance measures every vertex of a mesh
     * without cells: one such seed puts a seed spacing into the "distance" of two identical surfaces — of the reference to
     * itself too. They are counted (flagged) and removed before the surfaces are compared. */
    index_t isolated_ref = 0, isolated_b200 = 0;
    auto drop_isolated = [](Mesh& M) -> index_t {
        const index_t before = M.vertices.nb();
        if(before != 0 && M.facets.nb() != 0) {
            M.vertices.remove_isolated();
        }
        return before - M.vertices.nb();
    };
    isolated_ref = drop_isolated(ref.surface);
    isolated_b200 = drop_isolated(b200.surface);
    drop_isolated(ref.surface_mn);
    drop_isolated(b200.surface_mn);
    if(!g_volumetric && !ref.rdt.empty() && !b200.rdt.empty()) {
        /* the RAW restricted Delaunay triangulations (no mesh_repair): the north-star criterion on what the path itself
         * produces. mesh_repair, which compute_surface runs afterwards in simple mode, resolves non-manifold edges by
         * dropping triangles in an order that depends on the ROW order of its input (the reference's is its traversal
         * order); those edges are counted here so that the configurations are flagged explicitly. */
        auto soup = [&](const vector<index_t>& tris, const std::vector<double>& x, Mesh& M) {
            vector<double> v3(size_t(S) * 3);
            for(index_t i = 0; i < S; ++i) {
                for(index_t c = 0; c < 3; ++c) {
                    v3[index_t(i * 3 + c)] = x[size_t(i) * dim + c];
                }
            }
            vector<index_t> t = tris;
            M.facets.assign_triangle_mesh(3, v3, t, true);
        };
        Mesh A, B;
        soup(ref.rdt, ref.x_final, A);
        soup(b200.rdt, b200.x_final, B);
        drop_isolated(A);
        drop_isolated(B);
        const double sampling = 0.01 * bbox_diagonal(A);
        h_raw_ab = mesh_one_sided_Hausdorff_distance(A, B, sampling);
        h_raw_ba = mesh_one_sided_Hausdorff_distance(B, A, sampling);
        std::map<std::pair<index_t, index_t>, int> edges;
        for(const Tri& t : ta_all(ref.rdt)) {
            for(int e = 0; e < 3; ++e) {
                index_t a = t[e], b = t[(e + 1) % 3];
                ++edges[std::make_pair(std::min(a, b), std::max(a, b))];
            }
        }
        for(const auto& kv : edges) {
            nonmanifold_edges += kv.second > 2 ? 1 : 0;
        }
    }
    if(ref.surface_mn.facets.nb() > 0 && b200.surface_mn.facets.nb() > 0) {
        const double sampling = 0.01 * bbox_diagonal(ref.surface_mn);
        h_mn_ab = mesh_one_sided_Hausdorff_distance(ref.surface_mn, b200.surface_mn, sampling);
        h_mn_ba = mesh_one_sided_Hausdorff_distance(b200.surface_mn, ref.surface_mn, sampling);
    }
    if(ref.surface.facets.nb() > 0 && b200.surface.facets.nb() > 0) {
        double sampling = 0.01 * diag;
        /* control: the REFERENCE's own triangles in lexicographic order instead of traversal order, through the same
         * post-process as compute_surface (CVT.cpp:218-229: assign_triangle_mesh + mesh_repair). A non-zero distance to the
         * reference's surface shows how much of h_ab / h_ba is the order sensitivity of that post-process. */
        {
            std::vector<Tri> sorted;
            for(index_t f = 0; f + 2 < ref.rdt.size(); f += 3) {
                sorted.push_back(Tri{{ref.rdt[f], ref.rdt[f + 1], ref.rdt[f + 2]}});
            }
            std::sort(sorted.begin(), sorted.end());
            vector<index_t> tris;
            for(const Tri& t : sorted) {
                tris.push_back(t[0]); tris.push_back(t[1]); tris.push_back(t[2]);
            }
            vector<double> v3(size_t(S) * 3);
            for(index_t i = 0; i < S; ++i) {
                for(index_t c = 0; c < 3; ++c) {
                    v3[index_t(i * 3 + c)] = ref.x_final[size_t(i) * dim + c];
                }
            }
            Mesh ctrl;
            ctrl.facets.assign_triangle_mesh(3, v3, tris, true);
            mesh_repair(ctrl, MESH_REPAIR_DEFAULT, 1e-6 * bbox_diagonal(ctrl));
            drop_isolated(ctrl);
            h_ctrl = mesh_one_sided_Hausdorff_distance(ref.surface, ctrl, sampling);
        }
        h_ab = mesh_one_sided_Hausdorff_distance(ref.surface, b200.surface, sampling);
        h_ba = mesh_one_sided_Hausdorff_distance(b200.surface, ref.surface, sampling);
    }
    printf(
        "{\"volumetric\": %s, \"seeds\": %u, \"dim\": %u, \"pre_lloyd\": %u, \"lloyd\": %u, \"newton\": %u, \"on_gpu\": %s, "
        "\"max_abs_dx_lloyd\": %.3e, \"max_abs_dx_final\": %.3e, "
        "\"ref_triangles\": %zu, \"b200_triangles\": %zu, \"only_ref\": %zu, \"only_b200\": %zu, "
        "\"sizing_on_gpu\": %s, \"sizing_vertices\": %u, \"sizing_max_rel_diff\": %.3e, \"sizing_sampled_median_rel_diff\": %.3e, "
        "\"volume_tets_ref\": %zu, \"volume_tets_b200\": %zu, \"volume_bad_orientation\": %zu, \"compute_volume_cells_ref\": %u, \"compute_volume_cells_b200\": %u, "
        "\"ref_vertices\": %u, \"b200_vertices\": %u, "
        "\"hausdorff_ref_to_b200\": %.3e, \"hausdorff_b200_to_ref\": %.3e, \"hausdorff_ref_to_ref_sorted\": %.3e, \"bbox_diagonal\": %.6e, "
        "\"hausdorff_raw_rdt_ref_to_b200\": %.3e, \"hausdorff_raw_rdt_b200_to_ref\": %.3e, \"nonmanifold_edges\": %zu, \"isolated_seeds_ref\": %u, \"isolated_seeds_b200\": %u, "
        "\"multinerve_ref_vertices\": %u, \"multinerve_b200_vertices\": %u, \"multinerve_ref_facets\": %u, \"multinerve_b200_facets\": %u, "
        "\"hausdorff_multinerve_ref_to_b200\": %.3e, \"hausdorff_multinerve_b200_to_ref\": %.3e, "
        "\"nn_rows\": %u, \"nn_mismatch\": %u, "
        "\"t_ref_lloyd\": %.4f, \"t_ref_newton\": %.4f, \"t_b200_lloyd\": %.4f, \"t_b200_newton\": %.4f, \"ref_threads\": %u}\n",
        g_volumetric ? "true" : "false", S, dim, npre, nl, nn, b200.on_gpu ? "true" : "false",
        max_abs_diff(ref.x_lloyd, b200.x_lloyd), max_abs_diff(ref.x_final, b200.x_final),
        ta.size(), tb.size(), only_ref, only_b200,
        sizing_on_gpu ? "true" : "false", unsigned(sizing_vertices), sizing_max_rel, sizing_sampled_median_rel,
        vol_tets_ref, vol_tets_b200, vol_bad_orientation, unsigned(ref.nb_volume_tets), unsigned(b200.nb_volume_tets),
        ref.surface.vertices.nb(), b200.surface.vertices.nb(),
        h_ab, h_ba, h_ctrl, diag,
        h_raw_ab, h_raw_ba, nonmanifold_edges, isolated_ref, isolated_b200,
        ref.surface_mn.vertices.nb(), b200.surface_mn.vertices.nb(), ref.surface_mn.facets.nb(), b200.surface_mn.facets.nb(),
        h_mn_ab, h_mn_ba, nn_rows, nn_mismatch,
        ref.t_lloyd, ref.t_newton, b200.t_lloyd, b200.t_newton, unsigned(Process::maximum_concurrent_threads())
    );
    return 0;
}
