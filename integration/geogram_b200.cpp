/*
 * geogram_b200.cpp — adapter classes of geogram_b200.h over the C-ABI (include/b200cvt.h).
 * See INTEGRATION.md. Host-side glue only: no geometry is computed here.
 */
#include "geogram_b200.h"
#include <geogram/delaunay/LFS.h>
#include <geogram/mesh/mesh_geometry.h>

#include <geogram/basic/command_line.h>
#include <geogram/basic/logger.h>
#include <geogram/basic/progress.h>
#include <geogram/basic/stopwatch.h>
#include <geogram/mesh/mesh_remesh.h>

#include <algorithm>
#include <cstdlib>
#include <stdexcept>
#include <vector>

namespace {

    using namespace GEO;

    /* precondition failures of the reference are geo_assert(); here the C-ABI reports them as a status */
    void check(int status, const char* what) {
        if(status == B200CVT_OK) {
            return;
        }
        if(status == B200CVT_ERR_CANCELED) {
            throw TaskCanceled();
        }
        std::string msg = std::string(what) + ": " + b200cvt_last_error();
        Logger::err("B200") << msg << std::endl;
        throw std::runtime_error(msg);
    }

    unsigned long long hash_bytes(unsigned long long h, const void* p, size_t n);

    /* the same hash over large arrays: 1 MiB chunks hashed in parallel (OpenMP), chunk hashes combined in order */
    unsigned long long hash_bytes_parallel(unsigned long long h, const void* p, size_t n) {
        const size_t chunk = size_t(1) << 20;
        const size_t nchunks = (n + chunk - 1) / chunk;
        if(nchunks <= 4) {
            return hash_bytes(h, p, n);
        }
        std::vector<unsigned long long> part(nchunks);
        const unsigned char* b = static_cast<const unsigned char*>(p);
        #pragma omp parallel for schedule(static)
        for(long long c = 0; c < (long long)nchunks; ++c) {
            const size_t off = size_t(c) * chunk;
            part[size_t(c)] = hash_bytes(0x9E3779B97F4A7C15ull + (unsigned long long)c, b + off, std::min(chunk, n - off));
        }
        return hash_bytes(h, part.data(), sizeof(unsigned long long) * nchunks);
    }

    unsigned long long hash_bytes(unsigned long long h, const void* p, size_t n) {
        /* word-wise multiply-xorshift: only used to notice that the caller edited the mesh */
        const unsigned long long* w = static_cast<const unsigned long long*>(p);
        size_t nw = n / 8;
        for(size_t i = 0; i < nw; ++i) {
            h = (h ^ w[i]) * 0x9E3779B97F4A7C15ull;
            h ^= h >> 29;
        }
        const unsigned char* b = static_cast<const unsigned char*>(p) + nw * 8;
        for(size_t i = 0; i < n % 8; ++i) {
            h = (h ^ b[i]) * 0x100000001B3ull;
        }
        return h;
    }
}

namespace GEO {

    /************************ Delaunay backend ************************/

    Delaunay_B200NN::Delaunay_B200NN(coord_index_t dimension) :
        Delaunay_NearestNeighbors(dimension), h_(nullptr), tree_valid_(false) {
        check(b200cvt_create(-1, int(dimension), 0, &h_), "b200cvt_create");
    }

    Delaunay_B200NN::~Delaunay_B200NN() {
        b200cvt_destroy(h_);
    }

    void Delaunay_B200NN::set_vertices(index_t nb_vertices, const double* vertices) {
        Delaunay::set_vertices(nb_vertices, vertices);   /* borrows the pointer (delaunay.cpp:210-213) */
        tree_valid_ = false;
        /* update_neighbors() (delaunay.cpp:260-274): sizes are reset to the default when the number of seeds changes and are
         * sticky otherwise */
        if(nb_vertices != neighbors_.nb_arrays()) {
            neighbors_.init(nb_vertices, default_nb_neighbors_);
            for(index_t i = 0; i < nb_vertices; ++i) {
                neighbors_.resize_array(i, default_nb_neighbors_, false);
            }
        }
        index_t kmax = 0;
        for(index_t i = 0; i < nb_vertices; ++i) {
            kmax = std::max(kmax, neighbors_.array_size(i));
        }
        kmax = std::max(kmax, default_nb_neighbors_);
        if(nb_vertices < 2) {
            return;
        }
        index_t k = std::min(kmax, nb_vertices - 1);
        if(k > B200CVT_KMAX) {
            /* lists larger than the GPU cap: the reference path */
            ensure_tree();
            Delaunay::update_neighbors();
            return;
        }
        check(b200cvt_set_seeds(h_, vertices, nb_vertices), "b200cvt_set_seeds");
        std::vector<uint32_t> idx(size_t(nb_vertices) * k), cnt(nb_vertices);
        check(b200cvt_knn(h_, k, idx.data(), cnt.data(), nullptr, nullptr), "b200cvt_knn");
        for(index_t i = 0; i < nb_vertices; ++i) {
            index_t want = std::min(neighbors_.array_size(i), nb_vertices - 1);
            /* a later duplicate is disconnected whatever the list size (delaunay_nn.cpp:123-134) */
            index_t n = (cnt[i] == 0) ? 0 : std::min<index_t>(want, cnt[i]);
            neighbors_.set_array(i, n, idx.data() + size_t(i) * k, false);
        }
    }

    void Delaunay_B200NN::ensure_tree() const {
        std::lock_guard<std::mutex> lock(tree_mutex_);
        if(!tree_valid_) {
            const_cast<Delaunay_B200NN*>(this)->nn_search()->set_points(nb_vertices(), vertex_ptr(0));
            tree_valid_ = true;
        }
    }

    void Delaunay_B200NN::enlarge_neighborhood(index_t i, index_t nb) {
        ensure_tree();
        Delaunay_NearestNeighbors::enlarge_neighborhood(i, nb);
    }

    index_t Delaunay_B200NN::nearest_vertex(const double* p) const {
        ensure_tree();
        return Delaunay_NearestNeighbors::nearest_vertex(p);
    }

    index_t Delaunay_B200NN::get_neighbors_internal(index_t v, index_t nb_neighbors, index_t* neighbors) const {
        ensure_tree();
        return Delaunay_NearestNeighbors::get_neighbors_internal(v, nb_neighbors, neighbors);
    }

    void b200_register() {
        static bool done = false;
        if(!done) {
            geo_register_Delaunay_creator(Delaunay_B200NN, "B200NN");
            done = true;
        }
    }

    /************************ RVD ************************/

    RestrictedVoronoiDiagramB200::RestrictedVoronoiDiagramB200(Delaunay* delaunay, Mesh* mesh) :
        RestrictedVoronoiDiagram(
            delaunay, mesh, (mesh->vertices.nb() > 0) ? mesh->vertices.point_ptr(0) : nullptr, mesh->vertices.dimension()
        ),
        h_(nullptr), h_vol_(nullptr), check_SR_(false), mesh_uploaded_version_(0), vol_uploaded_version_(0),
        nb_gpu_calls_(0) {
        ref_ = RestrictedVoronoiDiagram::create(delaunay, mesh);   /* also forces set_stores_neighbors(true), RVD.cpp:2552 */
        has_weights_ = mesh->vertices.attributes().is_defined("weight");
        if(has_weights_) {
            vertex_weight_.bind(mesh->vertices.attributes(), "weight");
        }
        if(dimension_ == 3 || dimension_ == 6) {
            check(b200cvt_create(-1, int(dimension_), 0, &h_), "b200cvt_create");
        }
    }

    RestrictedVoronoiDiagramB200::~RestrictedVoronoiDiagramB200() {
        if(h_ != nullptr) {
            b200cvt_destroy(h_);
        }
        if(h_vol_ != nullptr) {
            b200cvt_destroy(h_vol_);
        }
    }

    bool RestrictedVoronoiDiagramB200::gpu_eligible() const {
        /* the range is NO_INDEX/NO_INDEX until set_facets_range is called (RVD.cpp:2623-2624) = the whole mesh */
        const bool whole_mesh = (facets_begin_ == NO_INDEX && facets_end_ == NO_INDEX) ||
            (facets_begin_ == 0 && facets_end_ == mesh_->facets.nb());
        if(h_ == nullptr || ref_->exact_predicates() || delaunay_ == nullptr || delaunay_->nb_vertices() == 0) {
            return false;
        }
        /* The GPU path clips every cell by its nearest neighbours in distance order (and truncates at the stored list in
         * Lloyd mode), i.e. the semantics of the "NN" backends. With a true Delaunay backend (BDEL, PDEL, ...) the reference
         * clips by the full Delaunay neighbourhood: that stays on the reference implementation. */
        if(dynamic_cast<Delaunay_NearestNeighbors*>(delaunay_) == nullptr) {
            return false;
        }
        if(volumetric_) {
            /* tets only (the reference walks any cell type through its own tet decomposition), dimension 3 */
            const bool whole_cells = (tets_begin_ == NO_INDEX && tets_end_ == NO_INDEX) ||
                (tets_begin_ == 0 && tets_end_ == mesh_->cells.nb());
            return dimension_ == 3 && mesh_->cells.nb() > 0 && mesh_->cells.are_simplices() && whole_cells;
        }
        return mesh_->facets.nb() > 0 && mesh_->facets.are_simplices() && whole_mesh;
    }

    /* Cheap signature of the borrowed mesh: addresses, counts and a strided sample of the coordinates and corners. It is
     * checked on every compute_* call; the element tables are rebuilt and everything is hashed only when it changes or when
     * the caller asks for a full check (once per Lloyd_iterations / Newton_iterations / compute_RDT call, mesh_modified()). */
    unsigned long long RestrictedVoronoiDiagramB200::quick_signature() const {
        const index_t nv = mesh_->vertices.nb(), stride = mesh_->vertices.dimension();
        const index_t ne = volumetric_ ? mesh_->cells.nb() : mesh_->facets.nb();
        unsigned long long h = 0x452821E638D01377ull ^ (static_cast<unsigned long long>(nv) << 32) ^ ne ^
            (static_cast<unsigned long long>(stride) << 58) ^ (volumetric_ ? 1ull << 57 : 0ull);
        if(nv == 0 || ne == 0) {
            return h;
        }
        const double* P = mesh_->vertices.point_ptr(0);
        h = hash_bytes(h, &P, sizeof(P));
        const size_t nd = size_t(nv) * stride, step = std::max<size_t>(1, nd / 2048);
        for(size_t i = 0; i < nd; i += step) {
            h = hash_bytes(h, P + i, sizeof(double));
        }
        h = hash_bytes(h, P + nd - 1, sizeof(double));
        const index_t estep = std::max<index_t>(1, ne / 2048);
        for(index_t e = 0; e < ne; e += estep) {
            const index_t v = volumetric_ ? mesh_->cells.vertex(e, 0) : mesh_->facets.vertex(e, 0);
            const index_t w = volumetric_ ? mesh_->cells.vertex(e, 3) : mesh_->facets.vertex(e, 2);
            const unsigned long long vw = (static_cast<unsigned long long>(v) << 32) | w;
            h = hash_bytes(h, &vw, sizeof(vw));
        }
        return h;
    }

    const RestrictedVoronoiDiagramB200::MeshArrays& RestrictedVoronoiDiagramB200::arrays(bool full_check) {
        MeshArrays& A = volumetric_ ? vol_arrays_ : surf_arrays_;
        const unsigned long long q = quick_signature();
        /* the full hash is taken once per quick signature: a caller that edits coordinates in place without changing any
         * sampled value must call mesh_modified() */
        if(A.valid && q == A.quick && (!full_check || A.full_checked)) {
            return A;
        }
        const index_t nv = mesh_->vertices.nb(), stride = mesh_->vertices.dimension();
        const index_t per = volumetric_ ? 4 : 3;
        const index_t ne = volumetric_ ? mesh_->cells.nb() : mesh_->facets.nb();
        A.elems.resize(size_t(ne) * per);
        A.adj.resize(size_t(ne) * per);
        #pragma omp parallel for schedule(static)
        for(long long ee = 0; ee < (long long)ne; ++ee) {
            const index_t e = index_t(ee);
            for(index_t lv = 0; lv < per; ++lv) {
                A.elems[size_t(e) * per + lv] = volumetric_ ? mesh_->cells.vertex(e, lv) : mesh_->facets.vertex(e, lv);
                const index_t a = volumetric_ ? mesh_->cells.adjacent(e, lv) : mesh_->facets.adjacent(e, lv);
                A.adj[size_t(e) * per + lv] = (a == NO_INDEX) ? -1 : int32_t(a);
            }
        }
        /* the volumetric actions ignore the "weight" attribute (RVD.cpp:420,783) */
        A.weights.clear();
        if(has_weights_ && !volumetric_) {
            A.weights.resize(nv);
            for(index_t v = 0; v < nv; ++v) {
                A.weights[v] = vertex_weight_[v];
            }
        }
        unsigned long long hsh = hash_bytes_parallel(0x243F6A8885A308D3ull + nv + (volumetric_ ? 7 : 0), mesh_->vertices.point_ptr(0),
                                                     sizeof(double) * size_t(nv) * stride);
        hsh = hash_bytes_parallel(hsh, A.elems.data(), sizeof(uint32_t) * A.elems.size());
        hsh = hash_bytes_parallel(hsh, A.adj.data(), sizeof(int32_t) * A.adj.size());
        if(!A.weights.empty()) {
            hsh = hash_bytes(hsh, A.weights.data(), sizeof(double) * A.weights.size());
        }
        if(!A.valid || hsh != A.full) {
            ++A.version;
        }
        A.full = hsh;
        A.quick = q;
        A.valid = true;
        A.full_checked = true;
        return A;
    }

    void RestrictedVoronoiDiagramB200::mesh_modified() {
        surf_arrays_.valid = false; surf_arrays_.full_checked = false;
        vol_arrays_.valid = false; vol_arrays_.full_checked = false;
    }

    b200cvt_handle RestrictedVoronoiDiagramB200::handle(bool full_check) {
        const MeshArrays& A = arrays(full_check);
        const index_t nv = mesh_->vertices.nb(), stride = mesh_->vertices.dimension();
        if(volumetric_) {
            if(h_vol_ == nullptr) {
                check(b200cvt_create(-1, 3, 1, &h_vol_), "b200cvt_create");
            }
            if(vol_uploaded_version_ != A.version) {
                check(
                    b200cvt_set_mesh(h_vol_, mesh_->vertices.point_ptr(0), nv, stride, A.elems.data(), A.adj.data(),
                                     index_t(A.elems.size() / 4), nullptr),
                    "b200cvt_set_mesh"
                );
                vol_uploaded_version_ = A.version;
            }
            return h_vol_;
        }
        if(mesh_uploaded_version_ != A.version) {
            check(
                b200cvt_set_mesh(h_, mesh_->vertices.point_ptr(0), nv, stride, A.elems.data(), A.adj.data(),
                                 index_t(A.elems.size() / 3), A.weights.empty() ? nullptr : A.weights.data()),
                "b200cvt_set_mesh"
            );
            mesh_uploaded_version_ = A.version;
        }
        return h_;
    }

    /* status bits of the last evaluation that the caller should hear about (b200cvt.h: B200CVT_FLAG_*) */
    void RestrictedVoronoiDiagramB200::report_flags(b200cvt_handle h, index_t nb_seeds) {
        std::vector<uint8_t> fl(nb_seeds);
        if(nb_seeds == 0 || b200cvt_get_flags(h, fl.data()) != B200CVT_OK) {
            return;
        }
        index_t kmax = 0, overflow = 0;
        for(index_t i = 0; i < nb_seeds; ++i) {
            kmax += (fl[i] & B200CVT_FLAG_KMAX) ? 1 : 0;
            overflow += (fl[i] & B200CVT_FLAG_POLY_OVERFLOW) ? 1 : 0;
        }
        if(kmax != 0) {
            Logger::warn("B200") << kmax << " seed(s) needed more than " << B200CVT_KMAX
                                 << " neighbours: their cells are truncated (the reference grows the list without bound)" << std::endl;
        }
        if(overflow != 0) {
            Logger::warn("B200") << overflow << " seed(s) with a clipped polygon over the vertex budget" << std::endl;
        }
    }

    void RestrictedVoronoiDiagramB200::upload_seeds() {
        check(b200cvt_set_seeds(handle(), delaunay_->vertex_ptr(0), delaunay_->nb_vertices()), "b200cvt_set_seeds");
    }

    void RestrictedVoronoiDiagramB200::compute_centroids_on_surface(double* mg, double* m) {
        if(volumetric_ || !gpu_eligible()) {
            ref_->compute_centroids_on_surface(mg, m);
            return;
        }
        upload_seeds();
        check(b200cvt_centroids(h_, check_SR_ ? 1 : 0, mg, m), "b200cvt_centroids");
        report_flags(h_, delaunay_->nb_vertices());
        ++nb_gpu_calls_;
    }

    void RestrictedVoronoiDiagramB200::compute_CVT_func_grad_on_surface(double& f, double* g) {
        if(volumetric_ || !gpu_eligible()) {
            ref_->compute_CVT_func_grad_on_surface(f, g);
            return;
        }
        upload_seeds();
        check(b200cvt_funcgrad(h_, check_SR_ ? 1 : 0, &f, g), "b200cvt_funcgrad");
        report_flags(h_, delaunay_->nb_vertices());
        ++nb_gpu_calls_;
    }

    /* ---- state setters: kept in sync with the delegate ---- */

    void RestrictedVoronoiDiagramB200::set_delaunay(Delaunay* delaunay) {
        RestrictedVoronoiDiagram::set_delaunay(delaunay);
        ref_->set_delaunay(delaunay);
    }

    void RestrictedVoronoiDiagramB200::set_volumetric(bool x) {
        volumetric_ = x;
        ref_->set_volumetric(x);
    }

    void RestrictedVoronoiDiagramB200::set_check_SR(bool x) {
        check_SR_ = x;
        ref_->set_check_SR(x);
    }

    void RestrictedVoronoiDiagramB200::set_exact_predicates(bool x) {
        ref_->set_exact_predicates(x);
    }

    bool RestrictedVoronoiDiagramB200::exact_predicates() const {
        return ref_->exact_predicates();
    }

    void RestrictedVoronoiDiagramB200::set_facets_range(index_t facets_begin, index_t facets_end) {
        facets_begin_ = facets_begin;
        facets_end_ = facets_end;
        ref_->set_facets_range(facets_begin, facets_end);
    }

    void RestrictedVoronoiDiagramB200::set_tetrahedra_range(index_t tets_begin, index_t tets_end) {
        tets_begin_ = tets_begin;
        tets_end_ = tets_end;
        ref_->set_tetrahedra_range(tets_begin, tets_end);
    }

    /* ---- everything off the hot path: the unmodified reference ---- */

    /* the initial sampling runs on the device when the optimisation loops would: whole element range, simplices, dimension 3
     * or 6 (volumes: dimension 3, no vertex weights). It does not need the Delaunay object. */
    bool RestrictedVoronoiDiagramB200::sampling_eligible() const {
        if(h_ == nullptr) {
            return false;
        }
        if(volumetric_) {
            const bool whole_cells = (tets_begin_ == NO_INDEX && tets_end_ == NO_INDEX) ||
                (tets_begin_ == 0 && tets_end_ == mesh_->cells.nb());
            return dimension_ == 3 && mesh_->cells.nb() > 0 && mesh_->cells.are_simplices() && whole_cells && !has_weights_;
        }
        const bool whole_mesh = (facets_begin_ == NO_INDEX && facets_end_ == NO_INDEX) ||
            (facets_begin_ == 0 && facets_end_ == mesh_->facets.nb());
        return mesh_->facets.nb() > 0 && mesh_->facets.are_simplices() && whole_mesh;
    }

    bool RestrictedVoronoiDiagramB200::compute_initial_sampling_on_surface(double* p, index_t nb_points, bool verbose) {
        if(volumetric_ || !sampling_eligible() || nb_points == 0) {
            return ref_->compute_initial_sampling_on_surface(p, nb_points, verbose);
        }
        if(verbose) {
            Logger::out("RVD") << "Computing initial sampling on surface (B200), using dimension=" << index_t(dimension_) << std::endl;
        }
        int ok = 1;
        check(b200cvt_initial_sampling(handle(true), nb_points, p, &ok), "b200cvt_initial_sampling");
        if(ok == 0) {
            Logger::warn("Sampler") << "Did put all the points in the same triangle" << std::endl;
        }
        ++nb_gpu_calls_;
        return ok != 0;
    }

    bool RestrictedVoronoiDiagramB200::compute_initial_sampling_in_volume(double* p, index_t nb_points, bool verbose) {
        if(!volumetric_ || !sampling_eligible() || nb_points == 0) {
            return ref_->compute_initial_sampling_in_volume(p, nb_points, verbose);
        }
        if(verbose) {
            Logger::out("RVD") << "Computing initial sampling in volume (B200), using dimension=" << index_t(dimension_) << std::endl;
        }
        int ok = 1;
        check(b200cvt_initial_sampling(handle(true), nb_points, p, &ok), "b200cvt_initial_sampling");
        if(ok == 0) {
            Logger::warn("Sampler") << "Did put all the points in the same tetrahedron" << std::endl;
        }
        ++nb_gpu_calls_;
        return ok != 0;
    }

    void RestrictedVoronoiDiagramB200::compute_centroids_in_volume(double* mg, double* m) {
        if(!volumetric_ || !gpu_eligible()) {
            ref_->compute_centroids_in_volume(mg, m);
            return;
        }
        upload_seeds();
        check(b200cvt_centroids(handle(), check_SR_ ? 1 : 0, mg, m), "b200cvt_centroids");
        ++nb_gpu_calls_;
    }

    void RestrictedVoronoiDiagramB200::compute_CVT_func_grad_in_volume(double& f, double* g) {
        if(!volumetric_ || !gpu_eligible()) {
            ref_->compute_CVT_func_grad_in_volume(f, g);
            return;
        }
        upload_seeds();
        check(b200cvt_funcgrad(handle(), check_SR_ ? 1 : 0, &f, g), "b200cvt_funcgrad");
        ++nb_gpu_calls_;
    }

    void RestrictedVoronoiDiagramB200::compute_integration_simplex_func_grad(double& f, double* g, IntegrationSimplex* F) {
        ref_->compute_integration_simplex_func_grad(f, g, F);
    }

    void RestrictedVoronoiDiagramB200::project_points_on_surface(index_t nb_points, double* points, vec3* nearest, bool do_project) {
        ref_->project_points_on_surface(nb_points, points, nearest, do_project);
    }

    void RestrictedVoronoiDiagramB200::compute_RDT(
        vector<index_t>& simplices, vector<double>& embedding, RDTMode mode, const vector<bool>& seed_is_locked, MeshFacetsAABB* AABB
    ) {
        /* On the device: the simple mode (RVD.cpp:2353-2370) and the multinerve mode with or without RVC centroids / seed
         * preference (RVD.cpp:1901-2264) — the one CentroidalVoronoiTesselation::compute_surface takes by default — both with
         * check_SR = true as compute_surface sets it (CVT.cpp:194). RDT_SELECT_NEAREST / RDT_PROJECT_ON_SURFACE need the AABB
         * tree of the input surface: reference implementation. */
        const bool multinerve = (mode & RDT_MULTINERVE) != 0;
        const bool needs_aabb = (mode & (RDT_SELECT_NEAREST | RDT_PROJECT_ON_SURFACE)) != 0;
        if(volumetric_ && gpu_eligible() && check_SR_) {
            /* volumes: "only simple mode is supported" (RVD.cpp:2309), whatever the mode bits: the Delaunay tets whose Voronoi
             * vertex lies inside the domain, reoriented, with the seeds as embedding (RVD.cpp:2308-2335) */
            b200cvt_handle hv = handle(true);
            upload_seeds();
            uint64_t n = 0;
            check(b200cvt_rdt(hv, nullptr, 0, &n), "b200cvt_rdt");
            std::vector<uint32_t> tets(size_t(n) * 4);
            check(b200cvt_rdt(hv, tets.data(), n, &n), "b200cvt_rdt");
            simplices.assign(tets.begin(), tets.end());
            const index_t nb = delaunay_->nb_vertices();
            embedding.assign(delaunay_->vertex_ptr(0), delaunay_->vertex_ptr(0) + size_t(nb) * dimension_);
            report_flags(hv, nb);
            ++nb_gpu_calls_;
            return;
        }
        if(volumetric_ || !gpu_eligible() || !check_SR_ || needs_aabb || (!multinerve && mode != RDTMode(0))) {
            ref_->compute_RDT(simplices, embedding, mode, seed_is_locked, AABB);
            return;
        }
        if(multinerve) {
            handle(true);
            upload_seeds();
            const index_t nb = delaunay_->nb_vertices();
            std::vector<uint8_t> locked;
            if(seed_is_locked.size() != 0) {
                locked.resize(nb);
                for(index_t i = 0; i < nb; ++i) {
                    locked[i] = seed_is_locked[i] ? 1 : 0;
                }
            }
            const int centroids = (mode & RDT_RVC_CENTROIDS) ? 1 : 0, prefer = (mode & RDT_PREFER_SEEDS) ? 1 : 0;
            uint64_t nt = 0, nv = 0;
            check(b200cvt_rdt_multinerve(h_, centroids, prefer, locked.empty() ? nullptr : locked.data(), nullptr, 0, &nt,
                                         nullptr, nullptr, 0, &nv), "b200cvt_rdt_multinerve");
            std::vector<uint32_t> tri(size_t(nt) * 3);
            embedding.resize(size_t(nv) * dimension_);
            check(b200cvt_rdt_multinerve(h_, centroids, prefer, locked.empty() ? nullptr : locked.data(), tri.data(), nt, &nt,
                                         embedding.data(), nullptr, nv, &nv), "b200cvt_rdt_multinerve");
            simplices.assign(tri.begin(), tri.end());
            report_flags(h_, nb);
            ++nb_gpu_calls_;
            return;
        }
        handle(true);
        upload_seeds();
        uint64_t n = 0;
        check(b200cvt_rdt(h_, nullptr, 0, &n), "b200cvt_rdt");
        std::vector<uint32_t> tri(size_t(n) * 3);
        check(b200cvt_rdt(h_, tri.data(), n, &n), "b200cvt_rdt");
        simplices.assign(tri.begin(), tri.end());
        const index_t nb = delaunay_->nb_vertices();
        embedding.assign(delaunay_->vertex_ptr(0), delaunay_->vertex_ptr(0) + size_t(nb) * dimension_);
        ++nb_gpu_calls_;
    }

    void RestrictedVoronoiDiagramB200::compute_RVD(Mesh& M, coord_index_t dim, bool cell_borders_only, bool integration_simplices) {
        ref_->compute_RVD(M, dim, cell_borders_only, integration_simplices);
    }

    void RestrictedVoronoiDiagramB200::compute_RVC(index_t i, Mesh& M, Mesh& result, bool copy_symbolic_info) {
        ref_->compute_RVC(i, M, result, copy_symbolic_info);
    }

    void RestrictedVoronoiDiagramB200::for_each_polyhedron(RVDPolyhedronCallback& callback, bool symbolic, bool connected_comp_priority,
                                                           bool parallel) {
        ref_->for_each_polyhedron(callback, symbolic, connected_comp_priority, parallel);
    }

    void RestrictedVoronoiDiagramB200::for_each_polygon(RVDPolygonCallback& callback, bool symbolic, bool connected_comp_priority,
                                                        bool parallel) {
        ref_->for_each_polygon(callback, symbolic, connected_comp_priority, parallel);
    }

    void RestrictedVoronoiDiagramB200::create_threads() {
        /* the GPU path has no host threads and never reorders the caller's mesh (RVD.cpp:2390-2395 does);
         * the delegate creates its own on demand */
    }

    void RestrictedVoronoiDiagramB200::delete_threads() {
        ref_->delete_threads();
    }

    GEOGen::PointAllocator* RestrictedVoronoiDiagramB200::point_allocator() {
        return ref_->point_allocator();
    }

    /************************ CVT ************************/

    namespace {
        /* "default" resolves to algo:delaunay (Delaunay::create); when that is the kd-tree backend "NN", the B200NN backend
         * takes its place: same lists bit for bit, computed on the device. An explicit name is left alone. */
        std::string resolve_delaunay(const std::string& name) {
            b200_register();
            if(name == "default") {
                const std::string algo = CmdLine::arg_is_declared("algo:delaunay") ? CmdLine::get_arg("algo:delaunay") : std::string("NN");
                if(algo == "NN" || algo == "default") {
                    return "B200NN";
                }
            }
            return name;
        }
    }

    CentroidalVoronoiTesselationB200::CentroidalVoronoiTesselationB200(Mesh* mesh, coord_index_t dimension, const std::string& delaunay) :
        CentroidalVoronoiTesselation(mesh, dimension, resolve_delaunay(delaunay)), canceled_(false), last_on_gpu_(false),
        nb_gpus_(1), group_(nullptr), group_mesh_version_(0) {
        for(int i = 0; i < 4; ++i) {
            newton_info_[i] = 0;
        }
        /* the base constructor created the reference RVD; swap in the adapter (which keeps its own reference delegate) */
        RVD_ = new RestrictedVoronoiDiagramB200(delaunay_, mesh);
        const char* env = getenv("B200CVT_GPUS");
        if(env != nullptr && atoi(env) > 1) {
            nb_gpus_ = index_t(atoi(env));
        }
    }

    CentroidalVoronoiTesselationB200::~CentroidalVoronoiTesselationB200() {
        if(group_ != nullptr) {
            b200cvt_group_destroy(group_);
        }
    }

    void CentroidalVoronoiTesselationB200::set_nb_gpus(index_t n) {
        nb_gpus_ = std::max<index_t>(n, 1);
    }

    RestrictedVoronoiDiagramB200* CentroidalVoronoiTesselationB200::rvd_b200() {
        return dynamic_cast<RestrictedVoronoiDiagramB200*>(RVD_.get());
    }

    int CentroidalVoronoiTesselationB200::progress_trampoline(void* user, uint32_t, double, double) {
        CentroidalVoronoiTesselationB200* self = static_cast<CentroidalVoronoiTesselationB200*>(user);
        try {
            self->newiteration();      /* progress_->next() may throw TaskCanceled; it must not cross the C boundary */
        } catch(const TaskCanceled&) {
            self->canceled_ = true;
            return 1;
        }
        return 0;
    }

    /* the multi-GPU group with the current surface mesh (surfacic mode only), or nullptr: one GPU */
    b200cvt_group_handle CentroidalVoronoiTesselationB200::group(RestrictedVoronoiDiagramB200* rvd) {
        if(nb_gpus_ <= 1 || volumetric() || (dimension_ != 3 && dimension_ != 6)) {
            return nullptr;
        }
        if(group_ == nullptr) {
            check(b200cvt_group_create(int(nb_gpus_), int(dimension_), 0, &group_), "b200cvt_group_create");
        }
        const RestrictedVoronoiDiagramB200::MeshArrays& A = rvd->arrays(true);
        if(group_mesh_version_ != A.version) {
            check(
                b200cvt_group_set_mesh(group_, mesh_->vertices.point_ptr(0), mesh_->vertices.nb(), mesh_->vertices.dimension(),
                                       A.elems.data(), A.adj.data(), index_t(A.elems.size() / 3),
                                       A.weights.empty() ? nullptr : A.weights.data()),
                "b200cvt_group_set_mesh"
            );
            group_mesh_version_ = A.version;
        }
        return group_;
    }

    bool CentroidalVoronoiTesselationB200::begin_gpu_loop(index_t nb_iter, std::vector<uint8_t>& locked) {
        RestrictedVoronoiDiagramB200* rvd = rvd_b200();
        const index_t nb = nb_points();
        last_on_gpu_ = false;
        /* The device loops never read the Delaunay object; it only has to be attached to the points for gpu_eligible() and
         * for whoever reads it afterwards. (The reference calls set_vertices — a kd-tree build plus S queries — at every
         * iteration, CVT.cpp:147.) */
        if(rvd != nullptr && nb > 0 && (delaunay_->nb_vertices() != nb || delaunay_->vertices_ptr() != points_.data())) {
            delaunay_->set_vertices(nb, points_.data());
        }
        if(rvd == nullptr || nb == 0 || !rvd->gpu_eligible()) {
            return false;
        }
        if(progress_ != nullptr) {
            progress_->reset(nb_iter);
        }
        cur_iter_ = 0;
        nb_iter_ = nb_iter;
        locked.clear();
        if(point_is_locked_.size() != 0) {
            locked.resize(nb);
            for(index_t i = 0; i < nb; ++i) {
                locked[i] = point_is_locked_[i] ? 1 : 0;
            }
        }
        canceled_ = false;
        return true;
    }

    void CentroidalVoronoiTesselationB200::end_gpu_loop(b200cvt_handle h, int status, const char* what) {
        const index_t nb = nb_points();
        last_on_gpu_ = true;
        /* leave the Delaunay object as the reference loop does: attached to the final points, lists rebuilt */
        delaunay_->set_vertices(nb, points_.data());
        progress_ = nullptr;
        check(status, what);
        if(h != nullptr) {
            RestrictedVoronoiDiagramB200::report_flags(h, nb);
        }
    }

    void CentroidalVoronoiTesselationB200::Lloyd_iterations(index_t nb_iter) {
        std::vector<uint8_t> locked;
        if(!begin_gpu_loop(nb_iter, locked)) {
            CentroidalVoronoiTesselation::Lloyd_iterations(nb_iter);
            return;
        }
        RestrictedVoronoiDiagramB200* rvd = rvd_b200();
        const index_t nb = nb_points();
        RVD_->set_check_SR(false);
        b200cvt_group_handle g = group(rvd);
        if(g != nullptr) {
            int status = b200cvt_group_lloyd(g, nb_iter, locked.empty() ? nullptr : locked.data(), points_.data(), nb, progress_trampoline, this);
            end_gpu_loop(nullptr, status, "b200cvt_group_lloyd");
            return;
        }
        b200cvt_handle h = rvd->handle(true);
        int status = b200cvt_lloyd(h, nb_iter, locked.empty() ? nullptr : locked.data(), points_.data(), nb, progress_trampoline, this);
        end_gpu_loop(h, status, "b200cvt_lloyd");
    }

    void CentroidalVoronoiTesselationB200::Newton_iterations(index_t nb_iter, index_t m) {
        std::vector<uint8_t> locked;
        if(!simplex_func_.is_null() || !begin_gpu_loop(nb_iter, locked)) {
            CentroidalVoronoiTesselation::Newton_iterations(nb_iter, m);
            return;
        }
        RestrictedVoronoiDiagramB200* rvd = rvd_b200();
        const index_t nb = nb_points();
        RVD_->set_check_SR(true);
        b200cvt_group_handle g = group(rvd);
        if(g != nullptr) {
            int status = b200cvt_group_newton(
                g, nb_iter, m, locked.empty() ? nullptr : locked.data(), points_.data(), nb, progress_trampoline, this, newton_info_
            );
            end_gpu_loop(nullptr, status, "b200cvt_group_newton");
            return;
        }
        b200cvt_handle h = rvd->handle(true);
        int status = b200cvt_newton(
            h, nb_iter, m, locked.empty() ? nullptr : locked.data(), points_.data(), nb, progress_trampoline, this, newton_info_
        );
        end_gpu_loop(h, status, "b200cvt_newton");
    }

    /************************ sizing field ************************/

    void compute_sizing_field_b200(Mesh& M, double gradation, index_t nb_lfs_samples) {
        /* the points the medial axis is estimated from (mesh_geometry.cpp:266-292): a CVT sampling of the surface, or the
         * mesh vertices themselves */
        std::vector<double> pts;
        index_t nb_pts = 0;
        if(nb_lfs_samples != 0) {
            Logger::out("LFS") << "Sampling surface (B200)" << std::endl;
            CentroidalVoronoiTesselationB200 CVT(&M, 3);
            CVT.compute_initial_sampling(nb_lfs_samples);
            CVT.Lloyd_iterations(5);
            CVT.Newton_iterations(10);
            nb_pts = CVT.nb_points();
            pts.assign(CVT.embedding(0), CVT.embedding(0) + size_t(nb_pts) * 3);
        } else {
            nb_pts = M.vertices.nb();
            pts.reserve(size_t(nb_pts) * 3);
            for(index_t v: M.vertices) {
                for(index_t c = 0; c < 3; ++c) {
                    pts.push_back(M.vertices.point_ptr(v)[c]);
                }
            }
        }
        Logger::out("LFS") << "Computing medial axis" << std::endl;
        LocalFeatureSize LFS(nb_pts, pts.data());
        compute_sizing_field_lfs_b200(M, LFS, gradation);
    }

    bool compute_sizing_field_lfs_b200(Mesh& M, const LocalFeatureSize& LFS, double gradation) {
        /* compute_sizing_field_lfs (mesh_geometry.cpp:57-73): weight = max(lfs^2, (0.1 average edge length)^2)^(-2 gradation),
         * lfs^2 = squared distance to the nearest pole: one nearest-neighbour query per mesh vertex, on the device */
        double min_distance2 = 0.1 * surface_average_edge_length(M);
        min_distance2 = min_distance2 * min_distance2;
        const index_t nv = M.vertices.nb(), np = LFS.nb_poles();
        std::vector<double> q(size_t(nv) * 3);
        for(index_t v = 0; v < nv; ++v) {
            for(index_t c = 0; c < 3; ++c) {
                q[size_t(v) * 3 + c] = M.vertices.point_ptr(v)[c];
            }
        }
        std::vector<uint32_t> nearest(nv);
        bool on_gpu = false;
        if(np > 0 && nv > 0) {
            b200cvt_handle h = nullptr;
            if(b200cvt_create(-1, 3, 0, &h) == B200CVT_OK) {
                on_gpu = b200cvt_set_seeds(h, LFS.pole(0), np) == B200CVT_OK &&
                    b200cvt_nearest(h, q.data(), nv, nearest.data()) == B200CVT_OK;
                b200cvt_destroy(h);
            }
        }
        Attribute<double> weight(M.vertices.attributes(), "weight");
        for(index_t v = 0; v < nv; ++v) {
            const double* p = M.vertices.point_ptr(v);
            double lfs2;
            if(on_gpu) {
                const double* pole = LFS.pole(nearest[v]);
                lfs2 = geo_sqr(p[0] - pole[0]) + geo_sqr(p[1] - pole[1]) + geo_sqr(p[2] - pole[2]);
            } else {
                lfs2 = LFS.squared_lfs(p);      /* no device, or no pole: the reference's own query */
            }
            lfs2 = std::max(lfs2, min_distance2);
            weight[v] = pow(lfs2, -2.0 * gradation);
        }
        return on_gpu;
    }

    /************************ remesh_smooth ************************/

    void remesh_smooth_b200(
        Mesh& M_in, Mesh& M_out, index_t nb_points, coord_index_t dim, index_t nb_Lloyd_iter, index_t nb_Newton_iter,
        index_t Newton_m, bool adjust, double adjust_max_edge_distance, double adjust_border_importance
    ) {
        geo_argused(dim);   /* as in the reference: the dimension is the mesh's (mesh_remesh.cpp:78-82) */
        Stopwatch W("Remesh(B200)");
        CentroidalVoronoiTesselationB200 CVT(&M_in);
        if(nb_points == 0) {
            nb_points = M_in.vertices.nb();
        }
        CVT.compute_initial_sampling(nb_points, true);
        try {
            ProgressTask progress("Lloyd", 100);
            CVT.set_progress_logger(&progress);
            CVT.Lloyd_iterations(nb_Lloyd_iter);
        } catch(const TaskCanceled&) {
        }
        if(nb_Newton_iter != 0) {
            try {
                ProgressTask progress("Newton", 100);
                CVT.set_progress_logger(&progress);
                CVT.Newton_iterations(nb_Newton_iter, Newton_m);
            } catch(const TaskCanceled&) {
            }
        }
        CVT.RVD()->delete_threads();
        CVT.set_use_RVC_centroids(CmdLine::get_arg_bool("remesh:RVC_centroids"));
        CVT.compute_surface(&M_out, CmdLine::get_arg_bool("remesh:multi_nerve"));
        if(adjust) {
            mesh_adjust_surface(M_out, M_in, adjust_max_edge_distance, false, adjust_border_importance);
        }
    }
}
