/*
 * b200cvt.h — C-ABI of the B200-native CVT / restricted-Voronoi-diagram hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types. Each entry
 * point names the reference interface it replaces (paths relative to the GraphiteThree
 * tree; G/ = geogram/src/lib/geogram/). INTEGRATION.md shows the C++ adapter a geogram
 * maintainer adds on top (Delaunay factory backend, RVD subclass, CVT subclass).
 *
 * Conventions:
 *  - every function returns 0 on success, non-zero on error; b200cvt_last_error() then
 *    describes the failure. No exception crosses this boundary.
 *  - host pointers are borrowed for the duration of the call only.
 *  - one handle per host thread; a handle is bound to one CUDA device.
 *  - seeds are S x dim doubles (AoS, like CentroidalVoronoiTesselation::points_,
 *    G/voronoi/CVT.h:434-456); mesh vertices are nv x stride doubles of which the first
 *    `dim` are used (Mesh::vertices.point_ptr, G/mesh/mesh.h); elements are uint32
 *    vertex ids (index_t without GARGANTUA).
 *  - there is NO CPU fallback: every call fails with B200CVT_ERR_CUDA if no device.
 */
#ifndef B200CVT_H
#define B200CVT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200cvt_ctx* b200cvt_handle;

enum {
    B200CVT_OK = 0,
    B200CVT_ERR_ARG = 1,        /* precondition failure (geo_assert in the reference) */
    B200CVT_ERR_CUDA = 2,       /* CUDA runtime error or no device */
    B200CVT_ERR_STATE = 3,      /* call order (no mesh / no seeds) */
    B200CVT_ERR_CAPACITY = 4,   /* internal capacity exceeded after retries */
    B200CVT_ERR_CANCELED = 5    /* progress callback asked to stop (TaskCanceled) */
};

/* per-seed status bits (b200cvt_get_flags) */
enum {
    B200CVT_FLAG_EXHAUSTED = 1, /* neighbour list used up before the security-radius test passed
                                   (Lloyd mode, check_SR=0): cell truncated to k neighbours, G/voronoi/generic_RVD.h:2179-2181 */
    B200CVT_FLAG_TIE = 2,       /* exact distance tie in the kNN list: order among ties is traversal-defined
                                   in the reference (G/points/kd_tree.h:173-195) */
    B200CVT_FLAG_POLY_OVERFLOW = 4, /* clipped polygon exceeded the per-lane vertex budget */
    B200CVT_FLAG_KMAX = 8       /* check_SR=1 but the neighbourhood hit the implementation cap (B200CVT_KMAX) */
};

#define B200CVT_KMAX 252u       /* largest neighbour list (reference: unbounded, S-1); a seed that needs more is flagged */

/* progress callback: called after each Lloyd iteration / each L-BFGS iteration
 * (CentroidalVoronoiTesselation::newiteration, G/voronoi/CVT.cpp:340-345).
 * Return non-zero to cancel (ProgressTask::next throwing TaskCanceled). */
typedef int (*b200cvt_progress_cb)(void* user, uint32_t iter, double f, double gnorm);

/* Replaces: CentroidalVoronoiTesselation ctor = Delaunay::create(dim,"NN") +
 * RestrictedVoronoiDiagram::create (G/voronoi/CVT.cpp:56-75, G/voronoi/RVD.cpp:2540-2600).
 * device: CUDA ordinal, or -1 for the current device. dim: 3 or 6 (other dimensions stay on
 * the reference implementation). volumetric: RestrictedVoronoiDiagram::set_volumetric. */
int b200cvt_create(int device, int dim, int volumetric, b200cvt_handle* out);
void b200cvt_destroy(b200cvt_handle h);
const char* b200cvt_last_error(void);

/* Replaces: the borrowed GEO::Mesh* of RestrictedVoronoiDiagram (G/voronoi/RVD.h:103-147):
 * vertices (point_ptr / dimension stride), facets.vertex(f,lv) or cells.tet_vertex(t,lv),
 * facet_corners.adjacent_facet / cells.tet_adjacent (adjacency may be NULL: not needed by
 * compute_centroids / compute_CVT_func_grad), and the optional "weight" vertex attribute
 * (G/voronoi/RVD.cpp:143-146). The mesh is copied to the device (replicated on every GPU). */
int b200cvt_set_mesh(b200cvt_handle h, const double* vertices, uint32_t nv, uint32_t stride_doubles,
                     const uint32_t* elems, const int32_t* adjacency_or_null, uint32_t n_elems,
                     const double* weights_or_null);

/* Replaces: Delaunay_NearestNeighbors::set_vertices (G/delaunay/delaunay_nn.cpp:97-103):
 * uploads S seeds, Morton-sorts them into the uniform grid and rebuilds every neighbour list
 * at its current (sticky) size, default 20 (delaunay_nn.cpp:49, G/delaunay/delaunay.cpp:259-274). */
int b200cvt_set_seeds(b200cvt_handle h, const double* x, uint32_t S);

/* Replaces: RestrictedVoronoiDiagram::compute_initial_sampling (G/voronoi/RVD.h:193-207) = compute_initial_sampling_on_surface /
 * _in_volume (G/voronoi/RVD.cpp:1658-1698) = mesh_generate_random_samples_on_surface / _in_volume
 * (G/mesh/mesh_sampling.h:119-199, 280-360) on the mesh of the handle: nb_points seeds, the reference's own points bit for
 * bit (same mt19937_64 stream, same element order, same arithmetic; the reference's multithreaded create_threads() reorders
 * the caller's mesh first — this path never does, so the match is with the reference on the mesh as given). The seeds become
 * the handle's current seeds (as b200cvt_set_seeds_device) and are copied to x_out (nb_points x dim, may be NULL).
 * *ok_out (optional) = 0 when every sample fell into one element (the reference returns false and warns), 1 otherwise. */
int b200cvt_initial_sampling(b200cvt_handle h, uint32_t nb_points, double* x_out, int* ok_out);

/* Replaces: Delaunay::get_neighbors over all seeds (G/delaunay/delaunay.h:477-484) after a
 * set_vertices with k stored neighbours: idx_out is S x k (original seed indices, ascending
 * distance, padded with 0xffffffff), count_out is S, sqdist_out (optional) S x k squared
 * distances bit-equal to Geom::distance2 (G/basic/geometry_nd.h:65-74), flags_out (optional) S. */
int b200cvt_knn(b200cvt_handle h, uint32_t k, uint32_t* idx_out, uint32_t* count_out,
                double* sqdist_out, uint8_t* flags_out);

/* Replaces: Delaunay_NearestNeighbors::nearest_vertex (G/delaunay/delaunay_nn.cpp:147-149)
 * for nq query points (nq x dim doubles). */
int b200cvt_nearest(b200cvt_handle h, const double* q, uint32_t nq, uint32_t* out);

/* Replaces: RestrictedVoronoiDiagram::compute_centroids(double* mg, double* m)
 * (G/voronoi/RVD.h:284, impl G/voronoi/RVD.cpp:371-412 / 500-527). mg (S x dim) and m (S)
 * are ACCUMULATED INTO, as in the reference (caller zeroes, G/voronoi/CVT.cpp:149-150).
 * check_SR: RestrictedVoronoiDiagram::set_check_SR. Uses the seeds of the last set_seeds. */
int b200cvt_centroids(b200cvt_handle h, int check_SR, double* mg_accum, double* m_accum);

/* Replaces: RestrictedVoronoiDiagram::compute_CVT_func_grad(double& f, double* g)
 * (G/voronoi/RVD.h:335, impl G/voronoi/RVD.cpp:726-774 / 878-910). *f and g are accumulated into. */
int b200cvt_funcgrad(b200cvt_handle h, int check_SR, double* f_accum, double* g_accum);

/* Per-seed results of the last centroids/funcgrad call, original seed order:
 * flags (S), per-seed energy (S; funcgrad only), candidate (facet,seed) pair counts (S). */
/* Replaces: RestrictedVoronoiDiagram::compute_RDT(simplices, embedding, RDTMode(0)) for surfaces
 * (G/voronoi/RVD.cpp:2302-2372, simple mode: for_each_primal_triangle, G/voronoi/generic_RVD.h:575-619) with
 * check_SR = true as CentroidalVoronoiTesselation::compute_surface sets it (G/voronoi/CVT.cpp:194): one triangle
 * (s, b0, b1), s < b1 < b0 original seed indices (SymbolicVertex::bisector(0) is the larger one), per polygon vertex
 * that lies on two bisectors. Rows sorted
 * lexicographically (the reference's order is its traversal order); duplicates are kept as the reference keeps them.
 * The embedding the reference returns with the triangles is the seed array itself. Call with tri_out = NULL to get the
 * count, then with a buffer of cap_triangles rows. Partitioned handles return the triangles of their own seeds.
 * Needs the facet adjacency: the one given to b200cvt_set_mesh, else it is rebuilt from shared edges.
 * Volumetric handles (RVD.cpp:2308-2335, for_each_primal_tetrahedron, G/voronoi/generic_RVD.h:1017-1058, check_SR = true as
 * CentroidalVoronoiTesselation::compute_volume sets it, CVT.cpp:245): rows of FOUR original seed indices, one per Voronoi
 * vertex that lies inside the tetrahedralised domain (the Delaunay tet of its four seeds), indices ascending within a row
 * except that the first two are swapped where PCK::orient_3d of the four seeds is negative (the reference's reorientation;
 * exact arithmetic where the floating-point determinant is not conclusive), rows sorted and unique. */
int b200cvt_rdt(b200cvt_handle h, uint32_t* tri_out, uint64_t cap_triangles, uint64_t* n_out);

/* Replaces: RestrictedVoronoiDiagram::compute_RDT with RDT_MULTINERVE (| RDT_RVC_CENTROIDS | RDT_PREFER_SEEDS), the mode
 * CentroidalVoronoiTesselation::compute_surface asks for by default (G/voronoi/CVT.cpp:180-197; remesh:multi_nerve,
 * remesh:RVC_centroids) — GetConnectedComponentsPrimalTriangles, G/voronoi/RVD.cpp:1901-2264, over
 * compute_surfacic_with_cnx_priority, G/voronoi/generic_RVD.h:1856-2001; check_SR = true. One vertex per connected
 * component of every restricted Voronoi cell, one triangle (3 vertex indices) per restricted Voronoi vertex.
 * vertices_out: n_vertices x dim positions (seed, or centroid of the component: RVD.cpp:2195-2237, 2123-2146);
 * vertex_seed_out (optional): the original seed index of every vertex. The reference numbers vertices and orders rows by
 * its sequential traversal; here vertices are numbered by (seed, smallest facet of the component) and rows are sorted,
 * duplicates removed (the reference leaves that to mesh_postprocess_RDT). RDT_SELECT_NEAREST / RDT_PROJECT_ON_SURFACE
 * stay on the reference. Call with tri_out = NULL to compute and get the counts, then with buffers. */
int b200cvt_rdt_multinerve(b200cvt_handle h, int use_rvc_centroids, int prefer_seeds, const uint8_t* locked_or_null,
                           uint32_t* tri_out, uint64_t cap_triangles, uint64_t* n_tri,
                           double* vertices_out, uint32_t* vertex_seed_out_or_null, uint64_t cap_vertices, uint64_t* n_vertices);

int b200cvt_get_flags(b200cvt_handle h, uint8_t* flags_out);
int b200cvt_get_seed_energy(b200cvt_handle h, double* f_seed_out);
int b200cvt_get_stats(b200cvt_handle h, uint64_t* stats_out /* 16 entries, see b200cvt.cu */);

/* Replaces: CentroidalVoronoiTesselation::Lloyd_iterations (G/voronoi/CVT.cpp:133-167).
 * x_inout: S x dim seeds, updated in place. locked_or_null: S bytes (point_is_locked_).
 * Seeds stay on the device between iterations. */
int b200cvt_lloyd(b200cvt_handle h, uint32_t nb_iter, const uint8_t* locked_or_null,
                  double* x_inout, uint32_t S, b200cvt_progress_cb cb, void* user);

/* Replaces: CentroidalVoronoiTesselation::Newton_iterations (G/voronoi/CVT.cpp:272-338) =
 * HLBFGS (G/third_party/HLBFGS/HLBFGS.cpp:281-587) with M=m, eps=0, max_iter=nb_iter, each
 * evaluation = set_vertices + compute_CVT_func_grad(check_SR=true) + constrain_points.
 * Vectors, dot products and the More-Thuente line search stay on the device.
 * info_out (optional, 4 entries): iterations, evaluations, line-search info, reserved. */
int b200cvt_newton(b200cvt_handle h, uint32_t nb_iter, uint32_t m, const uint8_t* locked_or_null,
                   double* x_inout, uint32_t S, b200cvt_progress_cb cb, void* user, uint32_t* info_out);

/* ---- device-resident / multi-GPU entry points (bench.py, torch.distributed harness) ---- */

/* Shards the seeds by contiguous Morton range: this handle evaluates sorted positions
 * [rank*ceil(S/nranks), (rank+1)*ceil(S/nranks)) only. Default 0/1. */
int b200cvt_set_partition(b200cvt_handle h, uint32_t rank, uint32_t nranks);

/* Seeds already in device memory (S x dim doubles); same effect as b200cvt_set_seeds. */
int b200cvt_set_seeds_device(b200cvt_handle h, const double* d_x, uint32_t S);

/* The one exchange of the sharded path: an all-gather of each rank's updated
 * slice. The library packs its slice into d_slice (chunk_doubles doubles: ceil(S/nranks) x dim
 * values, then ceil(S/nranks) scalars), calls cb(user) — which must all-gather d_slice of every
 * rank into d_all (rank-major) on the device, e.g. torch.distributed.all_gather_into_tensor
 * over NCCL, and return 0 — and unpacks d_all. If the caller gave the handle its stream
 * (b200cvt_set_stream) the callback only has to ENQUEUE the collective so that it is ordered on
 * that stream (no host synchronisation); with the handle's private stream the library drains it
 * before the call and d_all must be complete when cb returns. Both buffers are owned by the
 * caller (torch tensors in bench.py). */
typedef int (*b200cvt_exchange_cb)(void* user);
uint64_t b200cvt_exchange_chunk_doubles(int dim, uint32_t S, uint32_t nranks);
int b200cvt_set_exchange(b200cvt_handle h, double* d_slice, double* d_all, uint64_t chunk_doubles,
                         b200cvt_exchange_cb cb, void* user);

/* ---- multi-GPU inside the library (replaces nothing in the reference: its RVD is one process / one mesh
 * copy, partitioned over host threads by RVD_Nd_Impl::create_threads, G/voronoi/RVD.cpp:2374-2453) ----
 * A communicator makes the handle one rank of N: seeds are evaluated by Morton range, the mesh is replicated, seed
 * positions travel with ONE NCCL all-gather per evaluation (Lloyd: the updated slices; Newton: the new trial point, in
 * place on the seed array), the Newton gradient reaches the rank that owns its L-BFGS slice with one reduce-scatter, and
 * L-BFGS vectors are sharded by ranges of original seed indices: every dot product is a local partial plus a sum of a few
 * doubles that the cooperative L-BFGS kernels exchange through peer-memory mailboxes (stores into every peer's HBM over
 * NVLink, CUDA IPC between processes). It supersedes b200cvt_set_partition + b200cvt_set_exchange.
 *   one process per GPU: rank 0 calls b200cvt_comm_unique_id, the 128 bytes are sent to every rank by any means (MPI,
 *                        torch.distributed, a file), every rank calls b200cvt_comm_init on its handle (collective).
 *   one process, N GPUs: b200cvt_group_create (below). */
#define B200CVT_COMM_ID_BYTES 128
int b200cvt_comm_unique_id(uint8_t* id_out /* B200CVT_COMM_ID_BYTES */);
int b200cvt_comm_init(b200cvt_handle h, const uint8_t* id, uint32_t rank, uint32_t nranks);
int b200cvt_comm_destroy(b200cvt_handle h);

/* One process, N GPUs (devices 0 .. n_gpus-1; n_gpus <= 0: all visible): the group owns one handle per GPU, a
 * communicator over them (ncclCommInitAll + peer access) and runs every call on one host thread per GPU. This is what
 * CentroidalVoronoiTesselationB200 uses when asked for more than one GPU: same signatures as
 * b200cvt_set_mesh / b200cvt_lloyd / b200cvt_newton, the progress callback is called by rank 0's thread, and a cancel
 * request stops all ranks at the same iteration. */
typedef struct b200cvt_group* b200cvt_group_handle;
int b200cvt_group_create(int n_gpus, int dim, int volumetric, b200cvt_group_handle* out);
void b200cvt_group_destroy(b200cvt_group_handle g);
uint32_t b200cvt_group_size(b200cvt_group_handle g);
b200cvt_handle b200cvt_group_member(b200cvt_group_handle g, uint32_t rank);
int b200cvt_group_set_mesh(b200cvt_group_handle g, const double* vertices, uint32_t nv, uint32_t stride_doubles,
                           const uint32_t* elems, const int32_t* adjacency_or_null, uint32_t n_elems,
                           const double* weights_or_null);
int b200cvt_group_lloyd(b200cvt_group_handle g, uint32_t nb_iter, const uint8_t* locked_or_null, double* x_inout, uint32_t S,
                        b200cvt_progress_cb cb, void* user);
int b200cvt_group_newton(b200cvt_group_handle g, uint32_t nb_iter, uint32_t m, const uint8_t* locked_or_null, double* x_inout,
                         uint32_t S, b200cvt_progress_cb cb, void* user, uint32_t* info_out);

/* point_is_locked_ (G/voronoi/CVT.h:375-410) for the device-resident loops; NULL unlocks all. */
int b200cvt_set_locked(b200cvt_handle h, const uint8_t* locked_or_null, uint32_t S);

/* Lloyd_iterations / Newton_iterations on the seeds already resident on the device
 * (b200cvt_set_seeds[_device]); with a partition every rank must call them collectively. */
int b200cvt_lloyd_device(b200cvt_handle h, uint32_t nb_iter, b200cvt_progress_cb cb, void* user);
int b200cvt_newton_device(b200cvt_handle h, uint32_t nb_iter, uint32_t m, b200cvt_progress_cb cb, void* user,
                          uint32_t* info_out);

/* Copies the current seeds (original order) to device or host memory. */
int b200cvt_get_seeds_device(b200cvt_handle h, double* d_x_out);
int b200cvt_get_seeds(b200cvt_handle h, double* x_out);

/* Device timing of the phases of the last evaluation, milliseconds (CUDA events):
 * [0] sort+grid, [1] kNN, [2] candidate pairs, [3] clip+integrate, [4] update/reduce, [5] total. */
int b200cvt_get_timings(b200cvt_handle h, float* ms_out /* 6 entries */);
/* Runs the handle's kernels on the caller's stream (cudaStream_t; e.g. torch's current stream),
 * so that the caller's CUDA events bracket them. */
int b200cvt_set_stream(b200cvt_handle h, void* stream);
/* Cumulative device time per phase over all evaluations since the last reset, milliseconds (CUDA events on the
 * handle's stream): [0] sort+grid, [1] kNN + bisector table, [2] candidate pairs, [3] clip phase (compaction, class
 * sort, clip kernel, per-seed reduce, enlarged re-clips), [4] the clip kernel alone, [5] volumetric handles: the cell stage (vcell_kernel and its enlarged passes), included in [3]; surface handles:
 * the L-BFGS direction kernel and the push of the trial point to the peers (once per Newton iteration, outside the phases). */
int b200cvt_get_cumulative(b200cvt_handle h, double* ms_out /* 6 */, uint64_t* evals_out, int reset);
/* Roofline denominators measured on the device: non-tensor FP32 and FP64 FMA
 * throughput (TFLOP/s) and a STREAM-style copy (GB/s, read+write). Any pointer may be NULL. */
int b200cvt_measure_peaks(int device, double* fp32_tflops, double* fp64_tflops, double* copy_gbs);
/* Kernel launches issued by this handle since creation (bench.py gpu_launches). */
uint64_t b200cvt_launch_count(b200cvt_handle h);

#ifdef __cplusplus
}
#endif
#endif
