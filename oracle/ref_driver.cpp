/*
 * oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" entry points over the UNMODIFIED reference classes, compiled against
 * the headers where they lie under /root/reference and linked with
 * oracle/_ref/libgeogram_ref.so (see oracle/Makefile.ref). Used by:
 *   - tests/ (to pin the oracle restatement and to generate tests/golden/*),
 *   - bench.py --impl reference and bench.py's cpu_baseline leg.
 * Never loaded by the product (graphitethree_b200/).
 *
 * Reference classes driven here:
 *   GEO::Delaunay::create(dim,"NN")            geogram/delaunay/delaunay.cpp:152-185
 *   GEO::Delaunay_NearestNeighbors             geogram/delaunay/delaunay_nn.cpp:44-149
 *   GEO::RestrictedVoronoiDiagram::create      geogram/voronoi/RVD.cpp:2540-2600
 *   GEO::CentroidalVoronoiTesselation          geogram/voronoi/CVT.cpp:56-338
 */
#include <geogram/basic/common.h>
#include <geogram/basic/command_line.h>
#include <geogram/basic/command_line_args.h>
#include <geogram/basic/logger.h>
#include <geogram/basic/process.h>
#include <geogram/basic/stopwatch.h>
#include <geogram/mesh/mesh.h>
#include <geogram/delaunay/delaunay.h>
#include <geogram/voronoi/CVT.h>
#include <geogram/voronoi/RVD.h>
#include <geogram/voronoi/RVD_callback.h>
#include <geogram/voronoi/generic_RVD_polygon.h>
#include <geogram/basic/geometry_nd.h>

#include <cstring>
#include <string>
#include <vector>

using namespace GEO;

namespace {

    bool g_initialized = false;

    void ensure_init(int multithread, int max_threads) {
        if(!g_initialized) {
            GEO::initialize(GEO::GEOGRAM_INSTALL_NONE);
            CmdLine::import_arg_group("standard");
            CmdLine::import_arg_group("algo");
            CmdLine::import_arg_group("opt");
            CmdLine::import_arg_group("remesh");
            CmdLine::set_arg("log:quiet", "true");
            Logger::instance()->set_quiet(true);
            g_initialized = true;
        }
        Process::enable_multithreading(multithread != 0);
        if(max_threads > 0) {
            Process::set_max_threads(index_t(max_threads));
        }
    }

    /* counts evaluations and iterations; everything else is the stock class */
    class CountingCVT : public CentroidalVoronoiTesselation {
    public:
        CountingCVT(Mesh* mesh, coord_index_t dim, const std::string& delaunay)
            : CentroidalVoronoiTesselation(mesh, dim, delaunay) {}
        void funcgrad(index_t n, double* x, double& f, double* g) override {
            ++nb_funcgrad;
            CentroidalVoronoiTesselation::funcgrad(n, x, f, g);
            last_f = f;
        }
        void newiteration() override {
            ++nb_newiteration;
            CentroidalVoronoiTesselation::newiteration();
        }
        unsigned nb_funcgrad = 0, nb_newiteration = 0;
        double last_f = 0.0;
    };

    struct Ctx {
        Mesh mesh;
        coord_index_t dim = 3;
        bool volumetric = false;
        CountingCVT* cvt = nullptr;
        ~Ctx() { delete cvt; }
    };

    /* Collects every (seed, facet) polygon of the surfacic traversal
     * (GEOGen::RestrictedVoronoiDiagram::compute_surfacic_with_seeds_priority,
     * geogram/voronoi/generic_RVD.h:1318-1424). */
    class PolygonDump : public RVDPolygonCallback {
    public:
        PolygonDump(index_t dim, const Delaunay* del, double* f_per_seed,
                    double* m_per_seed, unsigned long long* counters,
                    std::vector<unsigned>* pairs)
            : dim_(dim), del_(del), f_(f_per_seed), m_(m_per_seed),
              cnt_(counters), pairs_(pairs) {}
        void operator()(index_t v, index_t t, const GEOGen::Polygon& P) const override {
            if(cnt_) {
                cnt_[0] += 1;                       /* pairs */
                cnt_[1] += P.nb_vertices();         /* output vertices */
                if(P.nb_vertices() >= 3) {
                    cnt_[2] += P.nb_vertices() - 2; /* fan triangles */
                    cnt_[3] += 1;                   /* non-empty pairs */
                }
            }
            if(pairs_ && P.nb_vertices() >= 3) {
                pairs_->push_back(v);
                pairs_->push_back(t);
            }
            const double* p0 = del_->vertex_ptr(v);
            for(index_t i = 1; i + 1 < P.nb_vertices(); ++i) {
                const double* p1 = P.vertex(0).point();
                const double* p2 = P.vertex(i).point();
                const double* p3 = P.vertex(i + 1).point();
                double a = Geom::triangle_area(p1, p2, p3, coord_index_t(dim_));
                if(m_) m_[v] += a;
                if(f_) {
                    /* same expression as ComputeCVTFuncGrad, RVD.cpp:586-596 */
                    double cur_f = 0.0;
                    for(index_t c = 0; c < dim_; c++) {
                        double u0 = p0[c] - p1[c];
                        double u1 = p0[c] - p2[c];
                        double u2 = p0[c] - p3[c];
                        cur_f += u0 * u0;
                        cur_f += u1 * (u0 + u1);
                        cur_f += u2 * (u0 + u1 + u2);
                    }
                    f_[v] += a * cur_f / 6.0;
                }
            }
        }
    private:
        index_t dim_;
        const Delaunay* del_;
        double* f_;
        double* m_;
        unsigned long long* cnt_;
        std::vector<unsigned>* pairs_;
    };
}

extern "C" {

    int ref_nb_threads() {
        return int(Process::maximum_concurrent_threads());
    }

    /* elems: triangles (3 ids) when volumetric==0, tetrahedra (4 ids) otherwise.
     * weights: optional per-vertex "weight" attribute (RVD.cpp:143-146). */
    void* ref_create(
        int dim, unsigned nv, const double* vertices,
        unsigned ne, const unsigned* elems, int volumetric,
        const double* weights, int multithread, int max_threads
    ) {
        ensure_init(multithread, max_threads);
        Ctx* c = new Ctx;
        c->dim = coord_index_t(dim);
        c->volumetric = volumetric != 0;
        c->mesh.vertices.set_dimension(index_t(dim));
        c->mesh.vertices.create_vertices(nv);
        for(index_t v = 0; v < nv; ++v) {
            std::memcpy(c->mesh.vertices.point_ptr(v), vertices + size_t(v) * dim,
                        sizeof(double) * size_t(dim));
        }
        if(volumetric) {
            for(index_t t = 0; t < ne; ++t) {
                c->mesh.cells.create_tet(elems[4*t], elems[4*t+1], elems[4*t+2], elems[4*t+3]);
            }
            c->mesh.cells.connect();
            c->mesh.cells.compute_borders();
        } else {
            for(index_t t = 0; t < ne; ++t) {
                c->mesh.facets.create_triangle(elems[3*t], elems[3*t+1], elems[3*t+2]);
            }
            c->mesh.facets.connect();
        }
        if(weights != nullptr) {
            Attribute<double> w(c->mesh.vertices.attributes(), "weight");
            for(index_t v = 0; v < nv; ++v) w[v] = weights[v];
        }
        c->cvt = new CountingCVT(&c->mesh, coord_index_t(dim), "NN");
        c->cvt->set_volumetric(c->volumetric);
        return c;
    }

    void ref_destroy(void* h) { delete static_cast<Ctx*>(h); }

    void ref_set_points(void* h, unsigned S, const double* x) {
        Ctx* c = static_cast<Ctx*>(h);
        c->cvt->set_points(S, x);
    }

    void ref_get_points(void* h, double* x) {
        Ctx* c = static_cast<Ctx*>(h);
        std::memcpy(x, c->cvt->embedding(0),
                    sizeof(double) * size_t(c->cvt->nb_points()) * c->dim);
    }

    int ref_initial_sampling(void* h, unsigned S) {
        Ctx* c = static_cast<Ctx*>(h);
        return c->cvt->compute_initial_sampling(S) ? 0 : 1;
    }

    void ref_lock_point(void* h, unsigned i) { static_cast<Ctx*>(h)->cvt->lock_point(i); }

    /* returns wall seconds spent inside Lloyd_iterations (CVT.cpp:133-167) */
    double ref_lloyd(void* h, unsigned nb_iter) {
        Ctx* c = static_cast<Ctx*>(h);
        double t0 = Stopwatch::now();
        c->cvt->Lloyd_iterations(nb_iter);
        return Stopwatch::now() - t0;
    }

    /* returns wall seconds spent inside Newton_iterations (CVT.cpp:272-307) */
    double ref_newton(void* h, unsigned nb_iter, unsigned m) {
        Ctx* c = static_cast<Ctx*>(h);
        double t0 = Stopwatch::now();
        c->cvt->Newton_iterations(nb_iter, m);
        return Stopwatch::now() - t0;
    }

    /* evaluations (funcgrad calls) and newiteration callbacks since creation */
    void ref_counters(void* h, unsigned* nb_funcgrad, unsigned* nb_newiteration) {
        Ctx* c = static_cast<Ctx*>(h);
        *nb_funcgrad = c->cvt->nb_funcgrad;
        *nb_newiteration = c->cvt->nb_newiteration;
    }

    /* delaunay->set_vertices on the current points; returns seconds */
    double ref_update_delaunay(void* h) {
        Ctx* c = static_cast<Ctx*>(h);
        double t0 = Stopwatch::now();
        c->cvt->delaunay()->set_vertices(c->cvt->nb_points(), c->cvt->embedding(0));
        return Stopwatch::now() - t0;
    }

    /* neighbour lists as stored by Delaunay_NearestNeighbors (delaunay_nn.cpp:73-145).
     * idx: S*kmax (padded with 0xffffffff), count: S. */
    void ref_get_neighbors(void* h, unsigned kmax, unsigned* idx, unsigned* count) {
        Ctx* c = static_cast<Ctx*>(h);
        Delaunay* d = c->cvt->delaunay();
        vector<index_t> N;
        for(index_t i = 0; i < d->nb_vertices(); ++i) {
            d->get_neighbors(i, N);
            count[i] = N.size();
            for(index_t j = 0; j < kmax; ++j) {
                idx[size_t(i) * kmax + j] = (j < N.size()) ? N[j] : 0xffffffffu;
            }
        }
    }

    unsigned ref_nearest_vertex(void* h, const double* p) {
        Ctx* c = static_cast<Ctx*>(h);
        return c->cvt->delaunay()->nearest_vertex(p);
    }

    /* RVD->compute_centroids (RVD.h:284); mg,m accumulated into (caller zeroes).
     * Precondition: ref_update_delaunay() called after the last point change. */
    double ref_centroids(void* h, int check_SR, double* mg, double* m) {
        Ctx* c = static_cast<Ctx*>(h);
        c->cvt->RVD()->set_check_SR(check_SR != 0);
        double t0 = Stopwatch::now();
        c->cvt->RVD()->compute_centroids(mg, m);
        return Stopwatch::now() - t0;
    }

    /* RVD->compute_CVT_func_grad (RVD.h:335) */
    double ref_funcgrad(void* h, int check_SR, double* f, double* g) {
        Ctx* c = static_cast<Ctx*>(h);
        c->cvt->RVD()->set_check_SR(check_SR != 0);
        double t0 = Stopwatch::now();
        c->cvt->RVD()->compute_CVT_func_grad(*f, g);
        return Stopwatch::now() - t0;
    }

    /* Surface only: serial traversal with per-seed energy / mass and event counters.
     * counters[0..3] = pairs, output vertices, fan triangles, non-empty pairs.
     * pairs_out (optional, capacity pairs_cap entries of (seed, facet)); returns the
     * number of non-empty pairs. */
    unsigned long long ref_polygons(
        void* h, int check_SR, double* f_per_seed, double* m_per_seed,
        unsigned long long* counters, unsigned* pairs_out, unsigned long long pairs_cap
    ) {
        Ctx* c = static_cast<Ctx*>(h);
        c->cvt->RVD()->set_check_SR(check_SR != 0);
        std::vector<unsigned> pairs;
        PolygonDump cb(c->dim, c->cvt->delaunay(), f_per_seed, m_per_seed, counters,
                       pairs_out ? &pairs : nullptr);
        c->cvt->RVD()->for_each_polygon(cb, false, false, false);
        unsigned long long n = pairs.size() / 2;
        if(pairs_out) {
            unsigned long long ncopy = n < pairs_cap ? n : pairs_cap;
            std::memcpy(pairs_out, pairs.data(), sizeof(unsigned) * 2 * ncopy);
        }
        return n;
    }

    /* Restricted Delaunay triangulation, simple mode (RVD.cpp:2302-2372) with the
     * seeds as vertex geometry. tri_out capacity tri_cap triangles; returns count. */
    unsigned ref_rdt(void* h, int mode, unsigned* tri_out, unsigned tri_cap,
                     double* vtx_out, unsigned vtx_cap, unsigned* nvtx_out) {
        Ctx* c = static_cast<Ctx*>(h);
        vector<index_t> simplices;
        vector<double> embedding;
        c->cvt->RVD()->set_check_SR(true);
        c->cvt->RVD()->compute_RDT(
            simplices, embedding, RestrictedVoronoiDiagram::RDTMode(mode)
        );
        index_t per = c->volumetric ? 4 : 3;
        unsigned n = unsigned(simplices.size() / per);
        unsigned ncopy = n < tri_cap ? n : tri_cap;
        if(tri_out) std::memcpy(tri_out, simplices.data(), sizeof(unsigned) * per * ncopy);
        unsigned nvtx = unsigned(embedding.size() / c->dim);
        if(nvtx_out) *nvtx_out = nvtx;
        if(vtx_out) {
            unsigned nv = nvtx < vtx_cap ? nvtx : vtx_cap;
            std::memcpy(vtx_out, embedding.data(), sizeof(double) * size_t(nv) * c->dim);
        }
        return n;
    }
}
