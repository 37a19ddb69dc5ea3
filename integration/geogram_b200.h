/*
 * geogram_b200.h — the reference-side binding of the B200 CVT path (what a geogram/Graphite
 * maintainer adds to their tree).
 *
 * Three adapter classes keep the reference's own C++ interfaces and route the hot path through
 * the C-ABI of include/b200cvt.h (libb200cvt.so); everything that is not on the hot path is
 * delegated to the unmodified reference implementation.
 *
 *   GEO::Delaunay_B200NN                    Delaunay factory backend "B200NN"
 *                                           (replaces Delaunay_NearestNeighbors,
 *                                            geogram/delaunay/delaunay_nn.cpp:44-149)
 *   GEO::RestrictedVoronoiDiagramB200       subclass of the abstract RestrictedVoronoiDiagram
 *                                           (geogram/voronoi/RVD.h:103-703); compute_centroids /
 *                                           compute_CVT_func_grad on the GPU, the other pure
 *                                           virtuals delegated to RestrictedVoronoiDiagram::create
 *   GEO::CentroidalVoronoiTesselationB200   subclass of CentroidalVoronoiTesselation
 *                                           (geogram/voronoi/CVT.h:68); Lloyd_iterations and
 *                                           Newton_iterations run device-resident
 *   GEO::remesh_smooth_b200                 same signature as GEO::remesh_smooth
 *                                           (geogram/mesh/mesh_remesh.h:92)
 *   GEO::compute_sizing_field_b200          same signature as GEO::compute_sizing_field
 *                                           (geogram/mesh/mesh_geometry.h:222)
 *
 * Compiled against the reference's headers where they lie; never copied into this repository.
 */
#ifndef GEOGRAM_B200_ADAPTER_H
#define GEOGRAM_B200_ADAPTER_H

#include <geogram/basic/common.h>
#include <geogram/delaunay/delaunay_nn.h>
#include <geogram/voronoi/RVD.h>
#include <geogram/voronoi/CVT.h>
#include <geogram/mesh/mesh.h>
#include <geogram/delaunay/LFS.h>

#include "../include/b200cvt.h"

#include <mutex>
#include <string>
#include <vector>

namespace GEO {

    /**
     * Delaunay backend "B200NN": neighbour lists computed by b200cvt_knn instead of the kd-tree.
     * Registered with geo_register_Delaunay_creator by b200_register() so that
     * Delaunay::create(dim, "B200NN") and `algo:delaunay=B200NN` select it.
     * enlarge_neighborhood() and nearest_vertex() (called by the CPU RVD when it is combined with this
     * backend) fall back to the reference kd-tree, which is built lazily on first use.
     */
    class Delaunay_B200NN : public Delaunay_NearestNeighbors {
    public:
        explicit Delaunay_B200NN(coord_index_t dimension);
        void set_vertices(index_t nb_vertices, const double* vertices) override;
        void enlarge_neighborhood(index_t i, index_t nb) override;
        index_t nearest_vertex(const double* p) const override;
    protected:
        ~Delaunay_B200NN() override;
        index_t get_neighbors_internal(index_t v, index_t nb_neighbors, index_t* neighbors) const override;
    private:
        void ensure_tree() const;
        b200cvt_handle h_;
        mutable bool tree_valid_;
        mutable std::mutex tree_mutex_;
    };

    /**
     * RestrictedVoronoiDiagram whose surfacic integrals run on the GPU. Not created by
     * RestrictedVoronoiDiagram::create (which is not a factory, RVD.cpp:2540-2600): construct explicitly.
     */
    class RestrictedVoronoiDiagramB200 : public RestrictedVoronoiDiagram {
    public:
        RestrictedVoronoiDiagramB200(Delaunay* delaunay, Mesh* mesh);

        /** true if the GPU path takes the next compute_* call: fast predicates, and either surfacic mode (dimension 3 or 6,
         *  triangulated surface, whole facet range) or volumetric mode (dimension 3, tetrahedral cells, whole tet range);
         *  otherwise the call is delegated to the reference. */
        bool gpu_eligible() const;
        /** true if the GPU path takes the next compute_initial_sampling call */
        bool sampling_eligible() const;
        /** the element tables the C-ABI takes, rebuilt only when the borrowed mesh changed */
        struct MeshArrays {
            std::vector<uint32_t> elems;
            std::vector<int32_t> adj;
            std::vector<double> weights;
            unsigned long long quick = 0, full = 0;
            unsigned version = 0;     /* bumped when the content changed */
            bool valid = false, full_checked = false;
        };
        /** the C-ABI handle with the current mesh uploaded. Every call checks a cheap signature of the borrowed mesh (addresses,
         *  counts, a strided sample); full_check also hashes all of it (done once per Lloyd_iterations / Newton_iterations /
         *  compute_RDT call). A caller that edits coordinates in place between two compute_* calls says so with mesh_modified(). */
        b200cvt_handle handle(bool full_check = false);
        const MeshArrays& arrays(bool full_check = false);
        void mesh_modified();
        /** logs the status bits a caller should hear about (neighbour cap reached, polygon budget exceeded) */
        static void report_flags(b200cvt_handle h, index_t nb_seeds);
        /** counts the compute_* calls served by the GPU (tests) */
        index_t nb_gpu_calls() const { return nb_gpu_calls_; }

        void set_delaunay(Delaunay* delaunay) override;
        void set_volumetric(bool x) override;
        bool compute_initial_sampling_on_surface(double* p, index_t nb_points, bool verbose) override;
        bool compute_initial_sampling_in_volume(double* p, index_t nb_points, bool verbose) override;
        void compute_centroids_on_surface(double* mg, double* m) override;
        void compute_centroids_in_volume(double* mg, double* m) override;
        void compute_CVT_func_grad_on_surface(double& f, double* g) override;
        void compute_CVT_func_grad_in_volume(double& f, double* g) override;
        void compute_integration_simplex_func_grad(double& f, double* g, IntegrationSimplex* F) override;
        void project_points_on_surface(index_t nb_points, double* points, vec3* nearest, bool do_project = false) override;
        void compute_RDT(
            vector<index_t>& simplices, vector<double>& embedding,
            RDTMode mode = RDTMode(RDT_RVC_CENTROIDS | RDT_PREFER_SEEDS),
            const vector<bool>& seed_is_locked = vector<bool>(), MeshFacetsAABB* AABB = nullptr
        ) override;
        void compute_RVD(Mesh& M, coord_index_t dim = 0, bool cell_borders_only = false, bool integration_simplices = false) override;
        void compute_RVC(index_t i, Mesh& M, Mesh& result, bool copy_symbolic_info = false) override;
        void for_each_polyhedron(RVDPolyhedronCallback& callback, bool symbolic = true, bool connected_comp_priority = true,
                                 bool parallel = false) override;
        void for_each_polygon(RVDPolygonCallback& callback, bool symbolic = true, bool connected_comp_priority = true,
                              bool parallel = false) override;
        void set_check_SR(bool x) override;
        void set_exact_predicates(bool x) override;
        bool exact_predicates() const override;
        void create_threads() override;
        void delete_threads() override;
        void set_facets_range(index_t facets_begin, index_t facets_end) override;
        void set_tetrahedra_range(index_t tets_begin, index_t tets_end) override;
        GEOGen::PointAllocator* point_allocator() override;

    protected:
        ~RestrictedVoronoiDiagramB200() override;
    private:
        void upload_seeds();
        RestrictedVoronoiDiagram_var ref_;   /* the unmodified reference implementation, for everything off the hot path */
        b200cvt_handle h_;                   /* surfacic handle */
        b200cvt_handle h_vol_;               /* volumetric handle (created on first use) */
        bool check_SR_;
        unsigned long long quick_signature() const;
        MeshArrays surf_arrays_, vol_arrays_;
        unsigned mesh_uploaded_version_, vol_uploaded_version_;
        index_t nb_gpu_calls_;
    };

    /**
     * CentroidalVoronoiTesselation whose optimisation loops stay on the device.
     * Same constructor arguments as the reference class; everything else (sampling, compute_surface,
     * locking, progress logger) is inherited.
     */
    class CentroidalVoronoiTesselationB200 : public CentroidalVoronoiTesselation {
    public:
        CentroidalVoronoiTesselationB200(Mesh* mesh, coord_index_t dimension = 0, const std::string& delaunay = "default");
        ~CentroidalVoronoiTesselationB200() override;
        void Lloyd_iterations(index_t nb_iter) override;
        void Newton_iterations(index_t nb_iter, index_t m = 7) override;
        /** statistics of the last Newton_iterations: iterations, function evaluations, line-search info */
        const unsigned* last_newton_info() const { return newton_info_; }
        /** true if the last Lloyd/Newton call ran on the GPU */
        bool last_call_on_gpu() const { return last_on_gpu_; }
        /** number of GPUs of this process the optimisation loops use (default 1, or the environment variable B200CVT_GPUS):
         *  more than one runs them through b200cvt_group_* — seeds sharded by Morton range, mesh replicated */
        void set_nb_gpus(index_t n);
        index_t nb_gpus() const { return nb_gpus_; }
    private:
        RestrictedVoronoiDiagramB200* rvd_b200();
        b200cvt_group_handle group(RestrictedVoronoiDiagramB200* rvd);
        bool begin_gpu_loop(index_t nb_iter, std::vector<uint8_t>& locked);
        void end_gpu_loop(b200cvt_handle h, int status, const char* what);
        index_t nb_gpus_;
        b200cvt_group_handle group_;
        unsigned group_mesh_version_;
        static int progress_trampoline(void* user, uint32_t iter, double f, double gnorm);
        bool canceled_;
        bool last_on_gpu_;
        unsigned newton_info_[4];
    };

    /** Registers the "B200NN" Delaunay backend (call once after GEO::initialize()). */
    void b200_register();

    /** GEO::compute_sizing_field (geogram/mesh/mesh_geometry.h:222, impl mesh_geometry.cpp:57-73, 263-294) with its data-parallel
     *  parts on the device: the optional pre-sampling (initial sampling, 5 Lloyd and 10 Newton iterations) runs through
     *  CentroidalVoronoiTesselationB200, and the local feature size of every mesh vertex — the squared distance to its nearest
     *  pole, LocalFeatureSize::squared_lfs (geogram/delaunay/LFS.h:97-104) — through b200cvt_nearest on the pole set. The poles
     *  themselves (3D Delaunay triangulation and sliver filtering, delaunay/LFS.cpp) stay on the reference implementation.
     *  Writes the same "weight" vertex attribute, bit for bit when nb_lfs_samples == 0. */
    void compute_sizing_field_b200(Mesh& M, double gradation = 1.0, index_t nb_lfs_samples = 0);

    /** The device part of it alone (compute_sizing_field_lfs, mesh_geometry.cpp:57-73): the "weight" attribute from a given set
     *  of poles. Returns false if the nearest-pole queries had to be served by the reference (no device). The reference's pole
     *  construction is not reproducible from run to run on degenerate inputs (parallel Delaunay: measured 10 081 against 10 086
     *  poles on the same tube mesh), so parity of the sizing field is stated on a shared LocalFeatureSize. */
    bool compute_sizing_field_lfs_b200(Mesh& M, const LocalFeatureSize& LFS, double gradation);

    /** GEO::remesh_smooth (geogram/mesh/mesh_remesh.h:92) with the B200 CVT. */
    void remesh_smooth_b200(
        Mesh& M_in, Mesh& M_out, index_t nb_points, coord_index_t dim = 0,
        index_t nb_Lloyd_iter = 5, index_t nb_Newton_iter = 30, index_t Newton_m = 7,
        bool adjust = true, double adjust_max_edge_distance = 0.5, double adjust_border_importance = 2.0
    );
}

#endif
