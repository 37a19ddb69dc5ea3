// b200cvt.cu — handle, orchestration and C-ABI of the B200 CVT / restricted Voronoi diagram path.
// See include/b200cvt.h for the reference interfaces each entry point replaces.
//
// Per evaluation (one Lloyd iteration or one Newton function evaluation), on one GPU:
//   1. seed_keys + radix sort + gather      : Morton-sort the seeds into the uniform grid
//   2. knn_kernel                           : 20 nearest seeds of every seed (FP64 exact) + their bisector rows
//   3. facet_home / facet_big / facet_task  : candidate (facet, seed) pairs with classification masks, rows per seed
//   4. compact_pairs + clip_win + reduce_pairs : flat pair list, thread-per-pair clip + integrate (FP64), per-seed sums
//      (clip_kernel: warp per seed, for enlarged neighbourhoods / overflowing polygons; clip_tet_kernel: volumetric mode)
//   5. update / scatter / L-BFGS kernels    : x <- mg/m (Lloyd) or f, g, direction and line search (Newton)
// Everything stays in device memory between iterations. DESIGN.md §4 describes every kernel.
#include "common.cuh"
#include "knn.cuh"
#include "clip.cuh"
#include "clip_flat.cuh"
#include "facet_pairs.cuh"
#include "clip_tet.cuh"
#include "vcell.cuh"
#include "comm.cuh"
#include "lbfgs.cuh"
#include "rdt.cuh"
#include "rdt_mn.cuh"
#include "mesh_prep.cuh"
#include "sampling.cuh"
#include "../../include/b200cvt.h"

#include <cub/cub.cuh>
#include <nvtx3/nvToolsExt.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

static thread_local std::string g_last_error;

// ---------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------
template <int D>
__global__ void seed_keys_kernel(const double* x, u32 S, GridParams g, u32* keys, u32* vals) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const double* p = x + (size_t)i * D;
    keys[i] = morton_encode(g, grid_coord(g, p[0], 0), grid_coord(g, p[1], 1), grid_coord(g, p[2], 2));
    vals[i] = i;
}

template <int D>
__global__ void gather_kernel(const double* x, const u32* keys, const u32* vals, u32 S,
                              SeedRec<D>* xs, u32* rank_of, uint2* cell_range) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    u32 o = vals[i];
    SeedRec<D> r;
#pragma unroll
    for (int c = 0; c < D; ++c) r.p[c] = x[(size_t)o * D + c];
    r.orig = (long long)o;
    xs[i] = r;
    rank_of[o] = i;
    u32 k = keys[i];
    if (i == 0 || keys[i - 1] != k) cell_range[k].x = i;
    if (i == S - 1 || keys[i + 1] != k) cell_range[k].y = i + 1;
}

// CentroidalVoronoiTesselation::Lloyd_iterations update rule (geogram/voronoi/CVT.cpp:153-162)
template <int D>
__global__ void lloyd_update_kernel(const SeedRec<D>* xs, const double* m, const double* mg, const uint8_t* locked,
                                    u32 qbegin, u32 qend, double* x_orig, double* slice_out, u32 slice_len) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= slice_len) return;
    u32 s = qbegin + i;
    if (s >= qend) {
        if (slice_out) for (int c = 0; c < D; ++c) slice_out[(size_t)i * D + c] = 0.0;
        return;
    }
    u32 o = (u32)xs[s].orig;
    double p[D];
#pragma unroll
    for (int c = 0; c < D; ++c) p[c] = xs[s].p[c];
    double mm = m[s];
    if (mm > 1e-30 && !(locked && locked[o])) {
        double sc = 1.0 / mm;
#pragma unroll
        for (int c = 0; c < D; ++c) p[c] = sc * mg[(size_t)s * D + c];
    }
#pragma unroll
    for (int c = 0; c < D; ++c) {
        if (x_orig) x_orig[(size_t)o * D + c] = p[c];
        if (slice_out) slice_out[(size_t)i * D + c] = p[c];
    }
}

template <int D>
__global__ void commit_sorted_kernel(const SeedRec<D>* xs, const double* all_sorted, u32 S, double* x_orig) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    u32 o = (u32)xs[i].orig;
#pragma unroll
    for (int c = 0; c < D; ++c) x_orig[(size_t)o * D + c] = all_sorted[(size_t)i * D + c];
}

// exchange chunk of one rank: [slice_len][D] vector part, then [slice_len] scalar part (zero padded)
template <int D>
__global__ void pack_slice_kernel(const double* out_s, const double* out_v, u32 qbegin, u32 qend, u32 slice_len, double* chunk) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= slice_len) return;
    u32 s = qbegin + i;
    bool in = s < qend;
#pragma unroll
    for (int c = 0; c < D; ++c) chunk[(size_t)i * D + c] = in ? out_v[(size_t)s * D + c] : 0.0;
    chunk[(size_t)slice_len * D + i] = in ? out_s[s] : 0.0;
}

// gathered chunks (rank-major, sorted order) -> x (original order)           [Lloyd]
//                                            -> g (original order), f_seed  [Newton]
template <int D>
__global__ void unpack_all_kernel(const SeedRec<D>* xs, const double* all, u32 S, u32 slice_len, const uint8_t* locked,
                                  int zero_locked, double* v_orig, double* s_sorted) {
    u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    u32 r = s / slice_len, i = s - r * slice_len;
    const double* chunk = all + (size_t)r * slice_len * (D + 1);
    u32 o = (u32)xs[s].orig;
    bool z = zero_locked && locked && locked[o];
#pragma unroll
    for (int c = 0; c < D; ++c) v_orig[(size_t)o * D + c] = z ? 0.0 : chunk[(size_t)i * D + c];
    if (s_sorted) s_sorted[s] = chunk[(size_t)slice_len * D + i];
}

// sorted -> original order
template <int D>
__global__ void scatter_results_kernel(const SeedRec<D>* xs, u32 qbegin, u32 qend, const double* out_s, const double* out_v,
                                       const uint8_t* flags, const u32* pair_cnt, const uint8_t* locked, int zero_locked,
                                       double* s_orig, double* v_orig, uint8_t* flags_orig, u32* cnt_orig) {
    u32 s = qbegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= qend) return;
    u32 o = (u32)xs[s].orig;
    if (s_orig) s_orig[o] = out_s[s];
    if (v_orig) {
        bool z = zero_locked && locked && locked[o];   // constrain_points, CVT.cpp:309-321
#pragma unroll
        for (int c = 0; c < D; ++c) v_orig[(size_t)o * D + c] = z ? 0.0 : out_v[(size_t)s * D + c];
    }
    if (flags_orig) flags_orig[o] = flags[s];
    if (cnt_orig) cnt_orig[o] = pair_cnt[s];
}

// Sharded Newton: the gradient of every evaluated seed goes straight into the L-BFGS slice of the rank that owns the seed
// (rows of L = ceil(S / nranks) original indices per rank), over NVLink when that is another GPU; locked seeds get zeros
// (constrain_points, CVT.cpp:309-321). Flags and pair counts stay local.
template <int D>
__global__ void scatter_gradient_peer_kernel(const SeedRec<D>* xs, u32 qbegin, u32 qend, const double* out_v, const uint8_t* flags,
                                             const u32* pair_cnt, const uint8_t* locked, PeerBufs pb, u32 L,
                                             uint8_t* flags_orig, u32* cnt_orig) {
    const u32 s = qbegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= qend) return;
    const u32 o = (u32)xs[s].orig;
    const u32 owner = o / L;
    double* dst = pb.g[owner] + (size_t)(o - owner * L) * D;
    const bool z = locked && locked[o];
#pragma unroll
    for (int c = 0; c < D; ++c) dst[c] = z ? 0.0 : out_v[(size_t)s * D + c];
    flags_orig[o] = flags[s];
    cnt_orig[o] = pair_cnt[s];
}

template <int D>
__global__ void knn_export_kernel(const SeedRec<D>* xs, const u32* nbr, const u32* nbr_n, const double* sqd, const uint8_t* flags,
                                  u32 S, u32 k, u32* idx_out, u32* cnt_out, double* sqd_out, uint8_t* flags_out) {
    u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    u32 o = (u32)xs[s].orig;
    u32 n = nbr_n[s];
    cnt_out[o] = n;
    for (u32 j = 0; j < k; ++j) {
        u32 t = nbr[(size_t)s * k + j];
        idx_out[(size_t)o * k + j] = (j < n && t != B200_NONE) ? (u32)xs[t].orig : B200_NONE;
        if (sqd_out) sqd_out[(size_t)o * k + j] = sqd[(size_t)s * k + j];
    }
    if (flags_out) flags_out[o] = flags[s];
}

// Sharded runs. One byte per grid cell (CELLF_*): an owned seed lies within 1 / 2 / 3 cells of it, and whether any seed at
// all lies in it. Seeds within two cells of an owned one get neighbour lists and bisector rows; the facet filter
// (facet_pairs.cuh) reads all four bits. Bytes are OR-ed four to a word with atomicOr, after a plain load that finds most
// of them already set.
__device__ __forceinline__ void cellflag_or(u32* words, u32 cid, u32 bits) {
    const u32 sh = 8u * (cid & 3u);
    if (((__ldcg(words + (cid >> 2)) >> sh) & bits) != bits) atomicOr(words + (cid >> 2), bits << sh);
}

// One thread per owned seed that is the first of its cell.
template <int D>
__global__ void mark_cells_kernel(const u32* sorted_keys, const void* xs_, u32 qbegin, u32 qend, GridParams g, u32* cellflag_words) {
    // A warp takes 32 consecutive owned seeds; every first seed of a grid cell (~1 in 9) has the cells within three cells
    // of its own marked by the WHOLE warp, 11 of the 343 neighbours per lane. The kernel is bound by instruction issue:
    // the offsets of a lane's 11 neighbours are unpacked once, and the cell coordinates come from the seed record
    // instead of a bit-by-bit decode of the Morton id.
    const SeedRec<D>* xs = (const SeedRec<D>*)xs_;
    const int lane = threadIdx.x & 31;
    const u32 s = qbegin + blockIdx.x * blockDim.x + threadIdx.x;
    u32 off[11];            // (dx + 3) | (dy + 3) << 4 | (dz + 3) << 8 | bits << 12, 0xffffffff: none
#pragma unroll
    for (int i = 0; i < 11; ++i) {
        const int e = lane + 32 * i;
        const int dz = e / 49, dy = (e / 7) % 7, dx = e % 7;
        const int m = max(max(abs(dx - 3), abs(dy - 3)), abs(dz - 3));
        off[i] = e < 343 ? (u32)(dx | (dy << 4) | (dz << 8) | ((m <= 1 ? 7 : (m <= 2 ? 6 : 4)) << 12)) : 0xffffffffu;
    }
    u32 key = 0;
    int c0 = 0, c1 = 0, c2 = 0;
    bool first = false;
    if (s < qend) {
        key = sorted_keys[s];
        first = (s == qbegin) || (sorted_keys[s - 1] != key);
        if (first) { c0 = grid_coord(g, xs[s].p[0], 0); c1 = grid_coord(g, xs[s].p[1], 1); c2 = grid_coord(g, xs[s].p[2], 2); }
    }
    u32 todo = __ballot_sync(B200_FULL, first);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int bx = __shfl_sync(B200_FULL, c0, src) - 3, by = __shfl_sync(B200_FULL, c1, src) - 3, bz = __shfl_sync(B200_FULL, c2, src) - 3;
#pragma unroll
        for (int i = 0; i < 11; ++i) {
            const u32 o = off[i];
            if (o == 0xffffffffu) continue;
            const int cx = bx + (int)(o & 15u), cy = by + (int)((o >> 4) & 15u), cz = bz + (int)((o >> 8) & 15u);
            if ((unsigned)cx >= (unsigned)g.res[0] || (unsigned)cy >= (unsigned)g.res[1] || (unsigned)cz >= (unsigned)g.res[2]) continue;
            cellflag_or(cellflag_words, morton_encode(g, cx, cy, cz), o >> 12);
        }
    }
}

// every seed: does it need lists / rows (an owned seed within two cells)? The first seed of a cell also records that the
// cell is occupied.
__global__ void need_flags_kernel(const u32* sorted_keys, u32 S, u32* cellflag_words, uint8_t* has_planes) {
    const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const u32 cid = sorted_keys[s];
    has_planes[s] = (uint8_t)(((__ldcg(cellflag_words + (cid >> 2)) >> (8u * (cid & 3u))) >> 1) & 1u);
    if (s == 0 || sorted_keys[s - 1] != cid) cellflag_or(cellflag_words, cid, CELLF_OCCUPIED);
}

__global__ void iota_u32_kernel(u32* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = (u32)i;
}

__global__ void fill_u32_kernel(u32* p, size_t n, u32 v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------
template <class T> struct DevBuf {
    T* p = nullptr; size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }          // every member buffer of the handle is released by `delete h`
    void ensure(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 16;
        CUDA_CHECK(cudaMalloc((void**)&p, want * sizeof(T)));
        cap = want;
    }
    // like ensure(), but the first `keep` elements survive a reallocation
    void grow_keep(size_t n, size_t keep, cudaStream_t st) {
        if (n <= cap) return;
        T* np = nullptr;
        const size_t want = n + n / 8 + 16;
        CUDA_CHECK(cudaMalloc((void**)&np, want * sizeof(T)));
        if (p && keep) CUDA_CHECK(cudaMemcpyAsync(np, p, std::min(keep, cap) * sizeof(T), cudaMemcpyDeviceToDevice, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        if (p) cudaFree(p);
        p = np; cap = want;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// shared by the member handles of a single-process group (b200cvt_group_*): one host thread per GPU
struct b200cvt_ctx;
struct GroupShared {
    std::vector<b200cvt_ctx*> members;
    std::mutex mu; std::condition_variable cv;
    int n = 1, waiting = 0; unsigned long long generation = 0;
    int cancel = 0;
    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned long long gen = generation;
        if (++waiting == n) { waiting = 0; ++generation; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != gen; });
    }
};

struct b200cvt_ctx {
    b200cvt_ctx() { memset(&pc, 0, sizeof(pc)); pc.nranks = 1; }
    int device = 0, dim = 3, volumetric = 0, num_sms = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    u64 launches = 0;
    // mesh
    u32 nv = 0, T = 0;
    bool has_mesh = false, weighted = false;
    DevBuf<double> tri, triw;
    DevBuf<uint8_t> tet_inner;         // volumetric: bit lf = face opposite to corner lf is shared with another tet
    // volumetric cell-first path (vcell.cuh): inside / boundary grid of the mesh, lists of the cells that need the tet path
    DevBuf<u32> vgrid_cells, vgrid_need, facet_list2; DevBuf<uint2> tet_gbox; VGrid vg; bool vgrid_valid = false; u32 vgrid_R = 0;
    size_t vgrid_ncell = 0; bool vneed_active = false;   // the facet walk is restricted to the tets near the cells of the tet path
    DevBuf<u32> vc_bnd, vc_redo_a, vc_redo_b, vc_n;
    DevBuf<u32> vt_ovf, vt_redo_a, vt_redo_b, vt_n;   // cell-by-tet kernel: cells it gives up on, cells whose list was used up
    bool use_vcell_tet = false;                       // B200CVT_VCELL_TET=1: boundary cells through vcell_tet_kernel (measured slower in Lloyd mode)
    bool use_vcell = true;
    // Newton loop on one GPU: the enlargement check of an evaluation is read with the line-search record
    bool defer_redo = false, redo_deferred = false; ClipArgs pending_clip;
    bool use_lbfgs_gram = true;                    // B200CVT_LBFGS_GRAM=0: the level-by-level direction kernel
    DevBuf<uint4> rdtv; DevBuf<unsigned long long> rdtv_n;   // volumetric RDT rows (mode 3)
    cudaEvent_t evc[2] = {nullptr, nullptr};   // around the cell stage
    cudaEvent_t evl[2] = {nullptr, nullptr};   // around the L-BFGS direction kernel (+ the push of the trial point)
    bool evl_used = false;
    bool evc_used = false;
    DevBuf<u32> facet_guess;
    // surface meshes: host copies kept for the facet adjacency the RDT extraction needs (built on first use)
    std::vector<u32> host_elems, host_perm;
    std::vector<int32_t> host_adj;
    DevBuf<int> facet_adj; bool facet_adj_valid = false;
    DevBuf<u32> perm_dev, inv_perm_dev; bool perm_dev_valid = false, vol_weights_dropped = false;
    DevBuf<u32> rdt_dev, rdt_n;
    std::vector<u32> rdt_host; bool rdt_valid = false;
    // multinerve RDT (rdt_mn.cuh): device scratch and the cached host result
    DevBuf<uint8_t> mn_pair_comp, mn_comp_border;
    DevBuf<u32> mn_ncomp, mn_comp_base, mn_vert_n, mn_tri, mn_tri_n, mn_vert_seed;
    DevBuf<double> mn_comp_m, mn_comp_mg, mn_emb;
    DevBuf<uint4> mn_vert;
    std::vector<u32> mn_tri_host, mn_vseed_host; std::vector<double> mn_emb_host;
    int mn_mode = -1; bool rdt_valid_mn = false;
    double bb_lo[3], bb_hi[3], mesh_measure = 0.0, mesh_vmax2 = 0.0;
    // seeds
    u32 S = 0;
    bool has_seeds = false, grid_valid = false;
    DevBuf<double> x;                 // [S][D] original order
    DevBuf<u32> keys, vals, keys2, vals2;
    DevBuf<unsigned char> cub_tmp;
    DevBuf<unsigned char> xs;         // SeedRec<D>[S]
    DevBuf<u32> rank_of;
    DevBuf<uint2> cell_range;
    GridParams g, g_prev;
    DevBuf<u32> mtab; int mtab_bits[3] = {-1, -1, -1};
    // kNN
    u32 k = 20, kstride = 20;
    bool knn_valid = false, planes_valid = false;   // planes_valid: the bisector table matches nbr (written by the kNN kernel)
    DevBuf<u32> cellflag;                   // sharded runs only: one byte per cell, four to a word
    DevBuf<uint8_t> has_planes;
    DevBuf<float4> facet_ball; DevBuf<float> facet_rad; DevBuf<u32> facet_cell, facet_list, facet_list_n;
    bool facet_cell_valid = false;
    DevBuf<u32> need_list, need_n;   // knn_fb: [0] count, [1..] queries knn_tile_kernel left to knn_kernel
    DevBuf<u32> nbr, nbr_n, nbr_prev;   // nbr_prev: lists of the previous evaluation, original indices, rows by original index
    bool prev_valid = false;
    DevBuf<double> sqd;
    DevBuf<uint8_t> flags;            // sorted order
    // redo (check_SR) tables
    DevBuf<u32> redo_a, redo_b, redo_n, nbr_big, nbr_big_n;
    // pairs
    u32 pair_cap = 0;
    DevBuf<u32> pair_cnt, pair_facet, pair_mask, max_cnt;
    DevBuf<uint2> tasks, big_list;
    // flat pair list (seed-major, facets ascending) + per-pair contributions
    DevBuf<u32> pair_off, flat_seed, flat_facet, slow_list;
    DevBuf<double> contrib, facet_area, planes;
    DevBuf<float> planes32;
    DevBuf<uint8_t> pstat;
    DevBuf<u32> flat_mask, iota;
    DevBuf<unsigned char> sort_tmp;
    size_t iota_filled = 0;
    u32 npairs = 0;
    // outputs (sorted order) and original-order staging
    DevBuf<double> out_s, out_v, s_orig, v_orig;
    DevBuf<uint8_t> flags_orig, locked;
    DevBuf<u32> cnt_orig;
    DevBuf<unsigned long long> stats;
    bool want_stats = false;
    bool has_results = false, has_energy = false;
    // partition + exchange (all-gather done by the host harness, e.g. torch.distributed over NCCL)
    u32 rank = 0, nranks = 1;
    double* x_slice = nullptr; double* x_all = nullptr; u64 x_chunk = 0;
    b200cvt_exchange_cb xcb = nullptr; void* xuser = nullptr;
    u64 exchanges = 0;
    // in-library communicator (b200cvt_comm_init*): NCCL for the bulk exchanges, peer mailboxes for the L-BFGS scalars
    ncclComm_t nccl = nullptr;
    bool has_comm = false;
    PeerComm pc;                                  // device view (pc.nranks == 1 without a communicator)
    DevBuf<PeerBox> boxes;                        // this rank's mailboxes [2][B200_MAX_RANKS]
    DevBuf<unsigned long long> pc_seq; DevBuf<double> pc_gtot; DevBuf<unsigned int> pc_err;
    void* ipc_opened[B200_MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                        nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf<double> xch_slice, xch_all, g_full;    // exchange buffers owned by the library (communicator mode)
    GroupShared* group = nullptr;                 // member of a single-process group
    // peer-mapped seed arrays / gradient slices of all ranks (Newton loop); IPC mappings are cached by handle
    PeerBufs pb; bool pb_valid = false;
    void* pb_ipc_ptr[2][B200_MAX_RANKS] = {};
    cudaIpcMemHandle_t pb_ipc_handle[2][B200_MAX_RANKS] = {};
    bool use_peer_exchange = true;                // B200CVT_PEER_EXCHANGE=0: NCCL reduce-scatter / all-gather instead
    // L-BFGS
    DevBuf<double> lb_g, lb_q, lb_px, lb_pg, lb_wa, lb_s, lb_y, lb_part;
    DevBuf<LbfgsScalars> lb_sc;
    u32 lb_dir_blocks = 0;            // grid of the cooperative direction kernel (all blocks resident)
    // timing
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t evk[2] = {nullptr, nullptr};   // around the clip kernel alone (roofline of the dominant kernel)
    bool ev_valid = false;
    u64 host_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool ev_pending = false;          // phase events recorded but not yet accumulated
    double cum_ms[6] = {0, 0, 0, 0, 0, 0};  // sort+grid, kNN+bisectors, pairs, clip phase (+redo), clip kernel alone, volumetric cell stage
    u64 cum_evals = 0;

    u32 slice_len() const { return (S + nranks - 1) / nranks; }
    u32 qbegin() const { return std::min<u64>((u64)rank * slice_len(), S); }
    u32 qend() const { return std::min<u64>((u64)qbegin() + slice_len(), S); }
};

#define LAUNCH(h, kernel, grid, block, smem, ...)                                  \
    do {                                                                            \
        kernel<<<(grid), (block), (smem), (h)->stream>>>(__VA_ARGS__);             \
        (h)->launches++;                                                            \
        CUDA_CHECK(cudaGetLastError());                                             \
    } while (0)

static inline u32 div_up(u64 a, u32 b) { return (u32)((a + b - 1) / b); }

// NVTX range per phase of an evaluation and per optimiser iteration (header-only NVTX 3: a no-op unless a profiler is
// attached), so that an Nsight Systems timeline reads like DESIGN.md §4
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct ArgError : public std::runtime_error { explicit ArgError(const std::string& s) : std::runtime_error(s) {} };
struct StateError : public std::runtime_error { explicit StateError(const std::string& s) : std::runtime_error(s) {} };
struct CapacityError : public std::runtime_error { explicit CapacityError(const std::string& s) : std::runtime_error(s) {} };
struct CanceledError : public std::runtime_error { explicit CanceledError(const std::string& s) : std::runtime_error(s) {} };

template <class F> static int guarded(F&& f) {
    try { f(); return B200CVT_OK; }
    catch (const ArgError& e) { g_last_error = e.what(); return B200CVT_ERR_ARG; }
    catch (const StateError& e) { g_last_error = e.what(); return B200CVT_ERR_STATE; }
    catch (const CapacityError& e) { g_last_error = e.what(); return B200CVT_ERR_CAPACITY; }
    catch (const CanceledError& e) { g_last_error = e.what(); return B200CVT_ERR_CANCELED; }
    catch (const CudaError& e) { g_last_error = e.what(); return B200CVT_ERR_CUDA; }
    catch (const std::exception& e) { g_last_error = e.what(); return B200CVT_ERR_CUDA; }
}

// stream sync that also folds the phase events of the last evaluation into the cumulative timers
static void sync_stream(b200cvt_ctx* h) {
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (h->ev_pending) {
        for (int i = 0; i < 4; ++i) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, h->ev[i], h->ev[i + 1]) == cudaSuccess) h->cum_ms[i] += t;
        }
        float tk = 0.f;
        if (h->evk[0] && cudaEventElapsedTime(&tk, h->evk[0], h->evk[1]) == cudaSuccess) h->cum_ms[4] += tk;
        if (h->evl_used) {
            float tl = 0.f;
            if (cudaEventElapsedTime(&tl, h->evl[0], h->evl[1]) == cudaSuccess) h->cum_ms[5] += tl;
            h->evl_used = false;
        }
        if (h->evc_used) {
            // volumetric: the cell stage runs between the kNN and the facet walk (its output restricts the walk); it is
            // accounted to the clip phase
            float tc = 0.f;
            if (cudaEventElapsedTime(&tc, h->evc[0], h->evc[1]) == cudaSuccess) { h->cum_ms[5] += tc; h->cum_ms[2] -= tc; h->cum_ms[3] += tc; }
            h->evc_used = false;
        }
        h->cum_evals++;
        h->ev_pending = false;
    }
}

// ---------------------------------------------------------------------------------------
// grid construction
// ---------------------------------------------------------------------------------------
#ifndef FACET_TASK_BLOCKS_PER_SM
#define FACET_TASK_BLOCKS_PER_SM 8u
#endif
#ifndef GRID_CELL_SURF
#define GRID_CELL_SURF 3.0
#endif
static void choose_grid(b200cvt_ctx* h, const double lo[3], const double hi[3]) {
    GridParams& g = h->g;
    double ext[3], maxext = 0.0;
    for (int a = 0; a < 3; ++a) { ext[a] = hi[a] - lo[a]; maxext = std::max(maxext, ext[a]); }
    if (!(maxext > 0.0)) maxext = 1.0;
    double S = (double)std::max<u32>(h->S, 1);
    double cell;
    if (h->has_mesh && h->mesh_measure > 0.0) {
        if (h->volumetric) cell = 2.0 * cbrt(h->mesh_measure / S);   // ~8 seeds per cell
        else cell = GRID_CELL_SURF * sqrt(h->mesh_measure / S);                 // ~9 seeds per occupied cell, 21-NN radius ~2.6 spacings
    } else {
        double vol = 1.0;
        for (int a = 0; a < 3; ++a) vol *= std::max(ext[a], maxext * 1e-3);
        cell = cbrt(vol * 8.0 / S);
    }
    cell = std::max(cell, maxext * 1e-6);
    for (;;) {
        int tb = 0;
        for (int a = 0; a < 3; ++a) {
            int res = (int)std::floor(ext[a] / cell) + 1;
            res = std::max(res, 1);
            int bits = 0;
            while ((1 << bits) < res) ++bits;
            g.res[a] = res; g.bits[a] = bits; tb += bits;
        }
        g.total_bits = tb;
        if (tb <= 26 && g.bits[0] <= 10 && g.bits[1] <= 10 && g.bits[2] <= 10) break;
        cell *= 1.2599210498948732;
    }
    // slight outward margin so that boundary points stay inside without clamping surprises
    for (int a = 0; a < 3; ++a) g.lo[a] = lo[a] - 1e-9 * maxext;
    g.h = cell; g.inv_h = 1.0 / cell;
    g.ncells = 1u << g.total_bits;
    // per-axis Morton bit tables (re-uploaded only when the bit layout changes)
    if (!h->mtab.p || h->mtab_bits[0] != g.bits[0] || h->mtab_bits[1] != g.bits[1] || h->mtab_bits[2] != g.bits[2]) {
        std::vector<u32> tab(3 * 1024, 0);
        GridParams t = g; t.mtab = nullptr;
        for (int c = 0; c < 1024; ++c) {
            if (c < (1 << g.bits[0])) tab[c] = morton_encode(t, c, 0, 0);
            if (c < (1 << g.bits[1])) tab[1024 + c] = morton_encode(t, 0, c, 0);
            if (c < (1 << g.bits[2])) tab[2048 + c] = morton_encode(t, 0, 0, c);
        }
        h->mtab.ensure(3 * 1024);
        CUDA_CHECK(cudaMemcpyAsync(h->mtab.p, tab.data(), sizeof(u32) * tab.size(), cudaMemcpyHostToDevice, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        for (int a = 0; a < 3; ++a) h->mtab_bits[a] = g.bits[a];
    }
    g.mtab = h->mtab.p;
    if (memcmp(&g, &h->g_prev, sizeof(GridParams)) != 0) { h->facet_cell_valid = false; h->g_prev = g; }
}

template <int D>
static void build_grid_t(b200cvt_ctx* h) {
    const u32 S = h->S;
    if (h->has_mesh) {
        choose_grid(h, h->bb_lo, h->bb_hi);
    } else {
        // bounding box of the seeds (host side: only taken without a mesh)
        std::vector<double> hx((size_t)S * D);
        CUDA_CHECK(cudaMemcpyAsync(hx.data(), h->x.p, sizeof(double) * (size_t)S * D, cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (u32 i = 0; i < S; ++i)
            for (int a = 0; a < 3; ++a) { double v = hx[(size_t)i * D + a]; lo[a] = std::min(lo[a], v); hi[a] = std::max(hi[a], v); }
        choose_grid(h, lo, hi);
    }
    h->keys.ensure(S); h->vals.ensure(S); h->keys2.ensure(S); h->vals2.ensure(S);
    h->xs.ensure((size_t)S * sizeof(SeedRec<D>));
    h->rank_of.ensure(S);
    h->cell_range.ensure(h->g.ncells);
    CUDA_CHECK(cudaMemsetAsync(h->cell_range.p, 0, sizeof(uint2) * (size_t)h->g.ncells, h->stream));
    LAUNCH(h, seed_keys_kernel<D>, div_up(S, 256), 256, 0, h->x.p, S, h->g, h->keys.p, h->vals.p);
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->keys.p, h->keys2.p, h->vals.p, h->vals2.p, (int)S, 0,
                                    std::max(h->g.total_bits, 1), h->stream);
    h->cub_tmp.ensure(tmp_bytes);
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(h->cub_tmp.p, tmp_bytes, h->keys.p, h->keys2.p, h->vals.p, h->vals2.p, (int)S, 0,
                                               std::max(h->g.total_bits, 1), h->stream));
    h->launches += 4;
    LAUNCH(h, gather_kernel<D>, div_up(S, 256), 256, 0, h->x.p, h->keys2.p, h->vals2.p, S,
           (SeedRec<D>*)h->xs.p, h->rank_of.p, h->cell_range.p);
    h->flags.ensure(S);
    CUDA_CHECK(cudaMemsetAsync(h->flags.p, 0, S, h->stream));
    h->grid_valid = true;
    h->knn_valid = false;
    h->has_results = false;
}

static void build_grid(b200cvt_ctx* h) {
    if (h->dim == 3) build_grid_t<3>(h); else build_grid_t<6>(h);
}

// ---------------------------------------------------------------------------------------
// kNN
// ---------------------------------------------------------------------------------------
template <int D>
static void launch_knn(b200cvt_ctx* h, const KnnArgs& a, u32 nq) {
    if (nq == 0) return;
    u32 kneed = std::min<u64>((u64)a.kstride + 2, a.S);
    u32 blocks = std::min<u32>(div_up(nq, KNN_WARPS), 148u * 16u);
    if (kneed <= 32) LAUNCH(h, (knn_kernel<D, 1>), blocks, KNN_WARPS * 32, 0, a);
    else if (kneed <= 64) LAUNCH(h, (knn_kernel<D, 2>), blocks, KNN_WARPS * 32, 0, a);
    else if (kneed <= 128) LAUNCH(h, (knn_kernel<D, 4>), blocks, KNN_WARPS * 32, 0, a);
    else LAUNCH(h, (knn_kernel<D, 8>), blocks, KNN_WARPS * 32, 0, a);
}

static void run_knn_main(b200cvt_ctx* h, u32 k, bool want_sqd, bool all_seeds, bool want_planes = false) {
    const u32 S = h->S;
    h->k = k; h->kstride = std::max<u32>(k, 1);
    h->nbr.ensure((size_t)S * h->kstride);
    h->nbr_n.ensure(S);
    if (want_sqd) h->sqd.ensure((size_t)S * h->kstride);
    KnnArgs a;
    memset(&a, 0, sizeof(a));
    a.xs = h->xs.p; a.cell_range = h->cell_range.p; a.rank_of = h->rank_of.p;
    a.query_list = nullptr; a.ksize = nullptr; a.out_by_slot = 0;
    a.k = k; a.kstride = h->kstride; a.S = S;
    a.qbegin = all_seeds ? 0 : h->qbegin(); a.qend = all_seeds ? S : h->qend();
    a.nbr = h->nbr.p; a.nbr_n = h->nbr_n.p; a.sqd = want_sqd ? h->sqd.p : nullptr; a.flags = h->flags.p; a.g = h->g;
    if (k == 20 && all_seeds) {
        // the evaluation path: seeds move little between evaluations, the old lists bound the new search radius
        h->nbr_prev.ensure((size_t)S * 20);
        a.prev_in = h->prev_valid ? h->nbr_prev.p : nullptr; a.prev_out = h->nbr_prev.p; a.prev_stride = 20;
    }
    if (want_planes) {
        h->planes.ensure((size_t)S * h->kstride * (h->dim == 3 ? 6 : h->dim + 2));
        h->planes32.ensure((size_t)S * h->kstride * (h->dim == 3 ? 4 : 8));
        a.planes = h->planes.p; a.planes32 = h->planes32.p;
    }
    if (h->dim == 3) launch_knn<3>(h, a, a.qend - a.qbegin); else launch_knn<6>(h, a, a.qend - a.qbegin);
    if (k == 20 && all_seeds) h->prev_valid = true;
    h->knn_valid = true;
    h->planes_valid = want_planes;
}

// ---------------------------------------------------------------------------------------
// evaluation = pairs + clip (+ neighbourhood enlargement when check_SR)
// ---------------------------------------------------------------------------------------
template <int D, int NC>
static void run_pairs_t(b200cvt_ctx* h) {
    const u32 S = h->S;
    if (h->pair_cap == 0) {
        double ratio = (double)h->T / (double)std::max<u32>(S, 1);
        u32 want = NC == 4 ? (u32)(ratio * 8.0 + 32.0) : (u32)(ratio * 4.0 + 24.0);
        u32 cap = 32; while (cap < want) cap <<= 1;
        h->pair_cap = cap;
    }
    const u32 nown = h->qend() - h->qbegin();
    h->pair_cnt.ensure((size_t)S + 1);
    h->pair_off.ensure((size_t)nown + 2);
    h->max_cnt.ensure(4);
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, h->pair_cnt.p, h->pair_off.p, (int)nown + 1, h->stream);
    h->cub_tmp.ensure(scan_bytes);
    for (int attempt = 0; attempt < 8; ++attempt) {
        h->pair_facet.ensure((size_t)S * h->pair_cap);
        h->pair_mask.ensure((size_t)S * h->pair_cap);
        CUDA_CHECK(cudaMemsetAsync(h->pair_cnt.p, 0, sizeof(u32) * ((size_t)S + 1), h->stream));
        CUDA_CHECK(cudaMemsetAsync(h->max_cnt.p, 0, 4 * sizeof(u32), h->stream));
        FacetPairArgs a;
        memset(&a, 0, sizeof(a));
        a.tri = h->tri.p; a.T = h->T; a.xs = h->xs.p; a.nbr = h->nbr.p; a.nbr_n = h->nbr_n.p; a.kstride = h->kstride;
        a.planes32 = h->planes32.p; a.has_planes = h->nranks > 1 ? h->has_planes.p : nullptr;
        a.cell_range = h->cell_range.p; a.rank_of = h->rank_of.p;
        a.facet_guess = h->facet_guess.p; a.S = S; a.qbegin = h->qbegin(); a.qend = h->qend();
        a.pair_cnt = h->pair_cnt.p; a.pair_facet = h->pair_facet.p; a.pair_mask = h->pair_mask.p; a.cap = h->pair_cap;
        a.max_cnt = h->max_cnt.p; a.stats = h->want_stats ? h->stats.p : nullptr; a.g = h->g;
        h->tasks.ensure(std::max<size_t>((size_t)h->T * 2, 1024));
        a.tasks = h->tasks.p; a.task_cap = (u32)std::min<size_t>(h->tasks.cap, 0xffffffffu); a.task_n = h->max_cnt.p + 1;
        if (NC == 3) {
            h->big_list.ensure(std::max<size_t>((size_t)h->T / 8, 4096));
            a.big_list = h->big_list.p; a.big_cap = (u32)std::min<size_t>(h->big_list.cap, 0xffffffffu); a.big_n = h->max_cnt.p + 2;
        }
        if (h->T > 0 && h->nranks > 1) {
            // the facets that can meet the cell of an owned seed
            h->facet_cell.ensure(h->T); h->facet_list.ensure(h->T); h->facet_list_n.ensure(1);
            if (!h->facet_cell_valid) {
                LAUNCH(h, (facet_cell_kernel<D, NC>), div_up(h->T, 256), 256, 0, h->tri.p, h->T, h->g, h->facet_cell.p);
                h->facet_cell_valid = true;
            }
            CUDA_CHECK(cudaMemsetAsync(h->facet_list_n.p, 0, sizeof(u32), h->stream));
            FacetFilterArgs fl;
            fl.ball = h->facet_ball.p; fl.rad = h->facet_rad.p; fl.facet_cell = h->facet_cell.p; fl.T = h->T; fl.cellflag = (const uint8_t*)h->cellflag.p;
            fl.cell_range = h->cell_range.p; fl.facet_guess = h->facet_guess.p; fl.rank_of = h->rank_of.p; fl.xs = h->xs.p;
            fl.g = h->g; fl.list = h->facet_list.p; fl.list_n = h->facet_list_n.p;
            LAUNCH(h, facet_filter_kernel<D>, div_up(h->T, 256 * FFILT_PER_THREAD), 256, 0, fl);
            a.facet_list = h->facet_list.p; a.facet_list_n = h->facet_list_n.p;
        }
        if (NC == 4 && h->T > 0 && h->vneed_active) {
            // volumetric: only the tets near a cell that goes through the (tet, seed) path (need bits of vcell_kernel)
            h->facet_list2.ensure((size_t)h->T + 1);
            u32* n2 = h->facet_list2.p + h->T;
            CUDA_CHECK(cudaMemsetAsync(n2, 0, sizeof(u32), h->stream));
            LAUNCH(h, vtet_filter_kernel, div_up(h->T, 256), 256, 0, h->tet_gbox.p, a.facet_list, a.facet_list_n, h->T, h->vg, h->facet_list2.p, n2);
            a.facet_list = h->facet_list2.p; a.facet_list_n = n2;
        }
        if (h->T > 0) {
            LAUNCH(h, (facet_home_kernel<D, NC>), div_up(h->T, 128), 128, 0, a);
            if (NC == 3) {
                const size_t smem_big = (size_t)BIG_WARPS * BIG_STACK * sizeof(BigPiece<D>);
                CUDA_CHECK(cudaFuncSetAttribute(facet_big_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big));
                LAUNCH(h, facet_big_kernel<D>, (u32)h->num_sms * 4u, BIG_WARPS * 32, smem_big, a);
            }
            LAUNCH(h, (facet_task_kernel<D, NC>), (u32)h->num_sms * FACET_TASK_BLOCKS_PER_SM, 128, 0, a);
        }
        // flat offsets of the owned seeds (pair_cnt[qend] is 0: only owned seeds receive pairs)
        size_t tb = h->cub_tmp.cap;
        CUDA_CHECK(cub::DeviceScan::ExclusiveSum(h->cub_tmp.p, tb, h->pair_cnt.p + h->qbegin(), h->pair_off.p, (int)nown + 1, h->stream));
        h->launches += 1;
        u32 mx[2] = {0, 0}, total = 0;
        CUDA_CHECK(cudaMemcpyAsync(mx, h->max_cnt.p, 2 * sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaMemcpyAsync(&total, h->pair_off.p + nown, sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (mx[1] > h->tasks.cap) { h->tasks.ensure((size_t)mx[1] + mx[1] / 4); continue; }   // task list overflow: rerun
        if (mx[0] <= h->pair_cap) { h->npairs = total; return; }
        u32 mx0 = mx[0];
        u32 cap = h->pair_cap; while (cap < mx0 + mx0 / 4) cap <<= 1;
        h->pair_cap = cap;
    }
    throw CapacityError("candidate pair rows keep overflowing");
}

template <int D>
static void launch_clip(b200cvt_ctx* h, ClipArgs& a) {
    if (a.nseeds == 0 && !a.nseeds_dev) return;
    size_t smem = (size_t)CLIP_WARPS * a.kstride * (D + 2) * sizeof(double);
    u32 blocks = a.nseeds_dev ? 148u * 4u : div_up(a.nseeds, CLIP_WARPS);
    if (h->weighted) {
        if (smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(clip_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(h, (clip_kernel<D, true>), blocks, CLIP_WARPS * 32, smem, a);
    } else {
        if (smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(clip_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(h, (clip_kernel<D, false>), blocks, CLIP_WARPS * 32, smem, a);
    }
}

// ---------------------------------------------------------------------------------------
// volumetric evaluation (DIM = 3): candidate (tet, seed) rows from the facet walk with 4 corners, one warp per seed
// ---------------------------------------------------------------------------------------
static void launch_clip_tet(b200cvt_ctx* h, TetClipArgs& a) {
    if (a.nseeds == 0 && !a.nseeds_dev) return;
    size_t smem = (size_t)TETC_WARPS * a.kstride * 4 * sizeof(double);
    u32 blocks = a.nseeds_dev ? 148u * 4u : div_up(a.nseeds, TETC_WARPS);
    if (a.mode == 3) {
        if (smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(clip_tet_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(h, clip_tet_kernel<true>, blocks, TETC_WARPS * 32, smem, a);
    } else {
        if (smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(clip_tet_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(h, clip_tet_kernel<false>, blocks, TETC_WARPS * 32, smem, a);
    }
}

#ifndef COMPACT_BLOCKS_PER_SM
#define COMPACT_BLOCKS_PER_SM 8u
#endif
#ifndef VC_NEWTON_K0
#define VC_NEWTON_K0 32
#endif
#ifndef VGRID_FACTOR
#define VGRID_FACTOR 2.0
#endif
// inside / boundary grid of the tet mesh (vcell.cuh), about one grid cell per seed spacing; rebuilt when the seed count
// changes by more than a quarter
static void ensure_vgrid(b200cvt_ctx* h) {
    double maxext = 0.0;
    for (int a = 0; a < 3; ++a) maxext = std::max(maxext, h->bb_hi[a] - h->bb_lo[a]);
    if (!(maxext > 0.0)) maxext = 1.0;
    u32 R = (u32)std::ceil(VGRID_FACTOR * std::cbrt((double)std::max<u32>(h->S, 1)));
    R = std::max<u32>(8, std::min<u32>(R, 250));
    if (h->vgrid_valid && 4 * R >= 3 * h->vgrid_R && 4 * R <= 5 * h->vgrid_R) return;
    VGrid& g = h->vg;
    const double margin = 1e-3 * maxext;
    g.h = (maxext + 2.0 * margin) / (double)R;
    g.inv_h = 1.0 / g.h;
    size_t ncell = 1;
    for (int a = 0; a < 3; ++a) {
        g.lo[a] = h->bb_lo[a] - margin;
        g.res[a] = std::max(1, (int)std::ceil((h->bb_hi[a] - h->bb_lo[a] + 2.0 * margin) * g.inv_h));
        ncell *= (size_t)g.res[a];
    }
    h->vgrid_cells.ensure(ncell / 4 + 1);
    h->vgrid_need.ensure(ncell / 32 + 2);
    h->tet_gbox.ensure(std::max<size_t>(h->T, 1));
    h->vgrid_ncell = ncell;
    g.cells = h->vgrid_cells.p;
    g.need = h->vgrid_need.p; g.need_all = h->vgrid_need.p + (ncell / 32 + 1);
    CUDA_CHECK(cudaMemsetAsync(g.cells, 0, sizeof(u32) * (ncell / 4 + 1), h->stream));
    if (h->T > 0) {
        const int nslab = std::max(1, std::min(g.res[2], 8));
        dim3 grid(div_up(h->T, 128), (unsigned)nslab);
        vgrid_mark_kernel<<<grid, 128, 0, h->stream>>>(h->tri.p, h->tet_inner.p, h->T, g, nslab, h->tet_gbox.p);
        h->launches++;
        CUDA_CHECK(cudaGetLastError());
    }
    h->vgrid_valid = true; h->vgrid_R = R;
}

static void launch_vcell(b200cvt_ctx* h, VCellArgs& a) {
    if (a.nseeds == 0 && !a.nseeds_dev) return;
    const u32 blocks = a.nseeds_dev ? (u32)h->num_sms * 8u : std::min<u32>(div_up(a.nseeds, VC_WARPS), (u32)h->num_sms * 8u);
    LAUNCH(h, vcell_kernel, blocks, VC_WARPS * 32, 0, a);
}

// kNN with kbig neighbours for the seeds of a list (rows by list slot)
static void knn_for_list(b200cvt_ctx* h, const u32* list, u32 n, u32 kbig) {
    h->nbr_big.ensure((size_t)n * kbig);
    h->nbr_big_n.ensure(n);
    KnnArgs a;
    memset(&a, 0, sizeof(a));
    a.xs = h->xs.p; a.cell_range = h->cell_range.p; a.rank_of = h->rank_of.p;
    a.query_list = list; a.ksize = nullptr; a.out_by_slot = 1;
    a.k = kbig; a.kstride = kbig; a.S = h->S; a.qbegin = 0; a.qend = n;
    a.nbr = h->nbr_big.p; a.nbr_n = h->nbr_big_n.p; a.sqd = nullptr; a.flags = h->flags.p; a.g = h->g;
    launch_knn<3>(h, a, n);
}

static void evaluate_volume(b200cvt_ctx* h, int mode, int check_SR) {
    const u32 S = h->S;
    const u32 nown = h->qend() - h->qbegin();
    h->out_s.ensure(S); h->out_v.ensure((size_t)S * 3);
    h->redo_a.ensure(S); h->redo_b.ensure(S); h->redo_n.ensure(4);
    CUDA_CHECK(cudaMemsetAsync(h->redo_n.p, 0, 4 * sizeof(u32), h->stream));
    const u32 kmax = std::min<u32>(B200CVT_KMAX, S > 0 ? S - 1 : 0);
    const bool vcell = h->use_vcell && nown > 0 && mode != 2;
    VCellArgs v;
    memset(&v, 0, sizeof(v));
    if (vcell) {
        // cell stage: every owned seed builds its Voronoi cell once; cells inside the domain are integrated on the spot
        NvtxRange r("b200cvt:cells");
        if (!h->evc[0]) for (int i = 0; i < 2; ++i) CUDA_CHECK(cudaEventCreate(&h->evc[i]));
        ensure_vgrid(h);
        h->vc_bnd.ensure(S); h->vc_redo_a.ensure(S); h->vc_redo_b.ensure(S); h->vc_n.ensure(4);
        CUDA_CHECK(cudaMemsetAsync(h->vc_n.p, 0, 4 * sizeof(u32), h->stream));
        CUDA_CHECK(cudaEventRecord(h->evc[0], h->stream));
        CUDA_CHECK(cudaMemsetAsync(h->vgrid_need.p, 0, sizeof(u32) * (h->vgrid_ncell / 32 + 2), h->stream));
        v.xs = h->xs.p; v.nbr = h->nbr.p; v.nbr_n = h->nbr_n.p; v.kstride = h->kstride; v.nbr_by_slot = 0;
        v.seed_list = nullptr; v.nseeds = nown; v.qbegin = h->qbegin();
        v.mode = mode; v.check_SR = check_SR; v.S = S;
        double maxext = 0.0;
        for (int a = 0; a < 3; ++a) maxext = std::max(maxext, h->bb_hi[a] - h->bb_lo[a]);
        if (!(maxext > 0.0)) maxext = 1.0;
        for (int a = 0; a < 3; ++a) { v.box_lo[a] = h->bb_lo[a] - 0.5 * maxext; v.box_hi[a] = h->bb_hi[a] + 0.5 * maxext; }
        v.vg = h->vg;
        v.out_s = h->out_s.p; v.out_v = h->out_v.p; v.flags = h->flags.p;
        v.redo_list = check_SR ? h->vc_redo_a.p : nullptr; v.redo_n = h->vc_n.p;
        v.bnd_list = h->vc_bnd.p; v.bnd_n = h->vc_n.p + 2;
        v.stats = h->want_stats ? h->stats.p : nullptr;
        v.tets = h->rdtv.p; v.tet_n = h->rdtv_n.p; v.tet_cap = h->rdtv.cap;
        u32 kbig0 = 40;
        if (check_SR && VC_NEWTON_K0 > 20 && kmax > 20) {
            // exact cells: whole cells need more than the 20 stored neighbours for most seeds (the radius test compares against
            // the farthest vertex of the whole cell), so the first pass already runs on longer lists
            const u32 k0 = std::min<u32>(VC_NEWTON_K0, kmax);
            h->iota.ensure(S);
            if (h->iota_filled < h->iota.cap) {
                LAUNCH(h, iota_u32_kernel, 1024, 256, 0, h->iota.p, h->iota.cap);
                h->iota_filled = h->iota.cap;
            }
            knn_for_list(h, h->iota.p + h->qbegin(), nown, k0);
            VCellArgs v0 = v;
            v0.nbr = h->nbr_big.p; v0.nbr_n = h->nbr_big_n.p; v0.kstride = k0; v0.nbr_by_slot = 1;
            v0.seed_list = h->iota.p + h->qbegin(); v0.nseeds = nown;
            if (k0 >= kmax) v0.redo_list = nullptr;
            launch_vcell(h, v0);
            kbig0 = 2 * k0;
        } else launch_vcell(h, v);
        if (check_SR) {
            // enlarge_neighborhood loop (generic_RVD.h:2330-2346) on whole cells, batched over the seeds that need it
            u32 kbig = kbig0;
            u32* cur_list = h->vc_redo_a.p; u32* nxt_list = h->vc_redo_b.p;
            int cur_slot = 0;
            for (;;) {
                u32 nredo = 0;
                CUDA_CHECK(cudaMemcpyAsync(&nredo, h->vc_n.p + cur_slot, sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
                CUDA_CHECK(cudaStreamSynchronize(h->stream));
                if (nredo == 0) break;
                h->host_stats[4] += nredo;
                kbig = std::min<u32>(kbig, kmax);
                knn_for_list(h, cur_list, nredo, kbig);
                const int nslot = cur_slot ^ 1;
                CUDA_CHECK(cudaMemsetAsync(h->vc_n.p + nslot, 0, sizeof(u32), h->stream));
                VCellArgs r2 = v;
                r2.nbr = h->nbr_big.p; r2.nbr_n = h->nbr_big_n.p; r2.kstride = kbig; r2.nbr_by_slot = 1;
                r2.seed_list = cur_list; r2.nseeds = nredo;
                r2.redo_list = kbig >= kmax ? nullptr : nxt_list; r2.redo_n = h->vc_n.p + nslot;
                launch_vcell(h, r2);
                std::swap(cur_list, nxt_list);
                cur_slot = nslot;
                if (kbig >= kmax) break;
                kbig *= 2;
            }
        }
        CUDA_CHECK(cudaEventRecord(h->evc[1], h->stream));
        h->evc_used = true;
    }
    // (tet, seed) stage for the rest: candidate rows from the facet walk with 4 corners, one warp per seed
    h->vneed_active = vcell;
    { NvtxRange r("b200cvt:facet walk (pairs)"); run_pairs_t<3, 4>(h); }
    h->vneed_active = false;
    CUDA_CHECK(cudaEventRecord(h->ev[3], h->stream));
    if (mode == 2) { CUDA_CHECK(cudaEventRecord(h->ev[4], h->stream)); return; }
    NvtxRange nvtx_clip("b200cvt:clip+integrate (tets)");
    TetClipArgs c;
    memset(&c, 0, sizeof(c));
    c.xs = h->xs.p; c.nbr = h->nbr.p; c.nbr_n = h->nbr_n.p; c.kstride = h->kstride; c.nbr_by_slot = 0;
    c.tet = h->tri.p; c.tet_inner = h->tet_inner.p;
    c.pair_cnt = h->pair_cnt.p; c.pair_facet = h->pair_facet.p; c.cap = h->pair_cap;
    c.seed_list = nullptr; c.nseeds = nown; c.qbegin = h->qbegin();
    if (vcell) { c.seed_list = h->vc_bnd.p; c.nseeds = 0; c.nseeds_dev = h->vc_n.p + 2; }
    c.mode = mode; c.check_SR = check_SR; c.S = S;
    c.out_s = h->out_s.p; c.out_v = h->out_v.p; c.flags = h->flags.p;
    c.redo_list = check_SR ? h->redo_a.p : nullptr; c.redo_n = h->redo_n.p;
    c.stats = h->want_stats ? h->stats.p : nullptr;
    c.tets = h->rdtv.p; c.tet_n = h->rdtv_n.p; c.tet_cap = h->rdtv.cap;
    CUDA_CHECK(cudaEventRecord(h->evk[0], h->stream));
    if (vcell && h->use_vcell_tet && mode != 3) {
        // the cells of the tet path, cooperatively: the cell is built once and clipped by the four planes of every candidate tet
        h->vt_ovf.ensure(S); h->vt_redo_a.ensure(S); h->vt_redo_b.ensure(S); h->vt_n.ensure(4);
        CUDA_CHECK(cudaMemsetAsync(h->vt_n.p, 0, 4 * sizeof(u32), h->stream));
        VCellTetArgs t;
        memset(&t, 0, sizeof(t));
        t.c = v;
        t.c.nbr = h->nbr.p; t.c.nbr_n = h->nbr_n.p; t.c.kstride = h->kstride; t.c.nbr_by_slot = 0;
        t.c.seed_list = h->vc_bnd.p; t.c.nseeds = 0; t.c.nseeds_dev = h->vc_n.p + 2;
        t.c.redo_list = check_SR ? h->vt_redo_a.p : nullptr; t.c.redo_n = h->vt_n.p;
        t.tet = h->tri.p; t.tet_inner = h->tet_inner.p;
        t.pair_cnt = h->pair_cnt.p; t.pair_facet = h->pair_facet.p; t.cap = h->pair_cap;
        t.ovf_list = h->vt_ovf.p; t.ovf_n = h->vt_n.p + 2;
        LAUNCH(h, vcell_tet_kernel, (u32)h->num_sms * (u32)VCT_MINBLK, VC_WARPS * 32, 0, t);
        if (check_SR) {
            u32 kbig = 40;
            u32* cur_list = h->vt_redo_a.p; u32* nxt_list = h->vt_redo_b.p;
            int cur_slot = 0;
            for (;;) {
                u32 nredo = 0;
                CUDA_CHECK(cudaMemcpyAsync(&nredo, h->vt_n.p + cur_slot, sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
                CUDA_CHECK(cudaStreamSynchronize(h->stream));
                if (nredo == 0) break;
                h->host_stats[4] += nredo;
                kbig = std::min<u32>(kbig, kmax);
                knn_for_list(h, cur_list, nredo, kbig);
                const int nslot = cur_slot ^ 1;
                CUDA_CHECK(cudaMemsetAsync(h->vt_n.p + nslot, 0, sizeof(u32), h->stream));
                VCellTetArgs r2 = t;
                r2.c.nbr = h->nbr_big.p; r2.c.nbr_n = h->nbr_big_n.p; r2.c.kstride = kbig; r2.c.nbr_by_slot = 1;
                r2.c.seed_list = cur_list; r2.c.nseeds = nredo; r2.c.nseeds_dev = nullptr;
                r2.c.redo_list = kbig >= kmax ? nullptr : nxt_list; r2.c.redo_n = h->vt_n.p + nslot;
                LAUNCH(h, vcell_tet_kernel, std::min<u32>(div_up(nredo, VC_WARPS), (u32)h->num_sms * 2u), VC_WARPS * 32, 0, r2);
                std::swap(cur_list, nxt_list);
                cur_slot = nslot;
                if (kbig >= kmax) break;
                kbig *= 2;
            }
        }
        // what is left: cells with more vertices than slots
        c.seed_list = h->vt_ovf.p; c.nseeds = 0; c.nseeds_dev = h->vt_n.p + 2;
    }
    launch_clip_tet(h, c);
    CUDA_CHECK(cudaEventRecord(h->evk[1], h->stream));
    if (check_SR) {
        // enlarge_neighborhood loop (generic_RVD.h:2330-2346), batched over the seeds that need it
        u32 kbig = 40;
        u32* cur_list = h->redo_a.p; u32* nxt_list = h->redo_b.p;
        int cur_slot = 0;
        for (;;) {
            u32 nredo = 0;
            CUDA_CHECK(cudaMemcpyAsync(&nredo, h->redo_n.p + cur_slot, sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
            if (nredo == 0) break;
            h->host_stats[4] += nredo;
            kbig = std::min<u32>(kbig, kmax);
            knn_for_list(h, cur_list, nredo, kbig);
            int nslot = cur_slot ^ 1;
            CUDA_CHECK(cudaMemsetAsync(h->redo_n.p + nslot, 0, sizeof(u32), h->stream));
            TetClipArgs r = c;
            r.nbr = h->nbr_big.p; r.nbr_n = h->nbr_big_n.p; r.kstride = kbig; r.nbr_by_slot = 1;
            r.seed_list = cur_list; r.nseeds = nredo; r.nseeds_dev = nullptr;
            r.redo_list = nxt_list; r.redo_n = h->redo_n.p + nslot;
            launch_clip_tet(h, r);
            std::swap(cur_list, nxt_list);
            cur_slot = nslot;
            if (kbig >= kmax) break;
            kbig *= 2;
        }
    }
    CUDA_CHECK(cudaEventRecord(h->ev[4], h->stream));
    h->has_results = (mode != 3);
    h->has_energy = (mode == 1);
    h->ev_valid = true;
    h->ev_pending = true;
}

// enlarge_neighborhood loop (generic_RVD.h:2179-2197), batched over the seeds that need it
template <int D>
static void surface_redo_loop(b200cvt_ctx* h, const ClipArgs& c) {
    const u32 S = h->S;
    u32 kbig = 40;
    u32* cur_list = h->redo_a.p; u32* nxt_list = h->redo_b.p;
    int cur_slot = 0;
    for (;;) {
        u32 nredo = 0;
        CUDA_CHECK(cudaMemcpyAsync(&nredo, h->redo_n.p + cur_slot, sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (nredo == 0) break;
        h->host_stats[4] += nredo;
        kbig = std::min<u32>(kbig, B200CVT_KMAX);
        kbig = std::min<u32>(kbig, S - 1);
        h->nbr_big.ensure((size_t)nredo * kbig);
        h->nbr_big_n.ensure(nredo);
        KnnArgs a;
        memset(&a, 0, sizeof(a));
        a.xs = h->xs.p; a.cell_range = h->cell_range.p; a.rank_of = h->rank_of.p;
        a.query_list = cur_list; a.ksize = nullptr; a.out_by_slot = 1;
        a.k = kbig; a.kstride = kbig; a.S = S; a.qbegin = 0; a.qend = nredo;
        a.nbr = h->nbr_big.p; a.nbr_n = h->nbr_big_n.p; a.sqd = nullptr; a.flags = h->flags.p; a.g = h->g;
        launch_knn<D>(h, a, nredo);
        int nslot = cur_slot ^ 1;
        CUDA_CHECK(cudaMemsetAsync(h->redo_n.p + nslot, 0, sizeof(u32), h->stream));
        ClipArgs r = c;
        r.nbr = h->nbr_big.p; r.nbr_n = h->nbr_big_n.p; r.kstride = kbig; r.nbr_by_slot = 1;
        r.seed_list = cur_list; r.nseeds = nredo;
        r.redo_list = nxt_list; r.redo_n = h->redo_n.p + nslot;
        launch_clip<D>(h, r);
        std::swap(cur_list, nxt_list);
        cur_slot = nslot;
        if (kbig >= std::min<u32>(B200CVT_KMAX, S - 1)) {
            // the kernel flags KMAX itself when the list cannot grow any more
            break;
        }
        kbig *= 2;
    }
}

template <int D>
static void evaluate_t(b200cvt_ctx* h, int mode, int check_SR) {
    if (!h->has_mesh) throw StateError("no mesh: call b200cvt_set_mesh first");
    if (!h->has_seeds) throw StateError("no seeds: call b200cvt_set_seeds first");
    const u32 S = h->S;
    if (!h->ev[0]) {
        for (int i = 0; i < 6; ++i) { CUDA_CHECK(cudaEventCreate(&h->ev[i])); CUDA_CHECK(cudaEventRecord(h->ev[i], h->stream)); }
        for (int i = 0; i < 2; ++i) CUDA_CHECK(cudaEventCreate(&h->evk[i]));
    }
    for (int i = 0; i < 2; ++i) CUDA_CHECK(cudaEventRecord(h->evk[i], h->stream));
    NvtxRange nvtx_eval(mode == 0 ? "b200cvt:evaluate(centroids)" : (mode == 1 ? "b200cvt:evaluate(func_grad)" : "b200cvt:evaluate(pairs)"));
    CUDA_CHECK(cudaEventRecord(h->ev[0], h->stream));
    { NvtxRange r("b200cvt:sort+grid"); if (!h->grid_valid) build_grid(h); }
    CUDA_CHECK(cudaEventRecord(h->ev[1], h->stream));
    nvtxRangePushA("b200cvt:kNN+bisectors");
    h->planes.ensure((size_t)S * h->kstride * PLANE_STRIDE(D));
    h->planes32.ensure((size_t)S * h->kstride * PLANE32_STRIDE(D));
    if (h->nranks == 1) {
        // neighbour lists + bisector tables of every seed (the facet walk may visit any seed)
        // (the kNN kernel writes the bisector rows with the lists; lists left by b200cvt_knn get theirs from plane_table_kernel)
        if (!h->knn_valid || h->k != 20) run_knn_main(h, 20, false, true, true);
        if (!h->planes_valid) {
            LAUNCH(h, plane_table_kernel<D>, std::min<u32>(div_up((u64)S * h->kstride, 256), (u32)h->num_sms * 16u), 256, 0,
                   h->xs.p, h->nbr.p, h->nbr_n.p, h->kstride, (const u32*)nullptr, 0u, S, (const u32*)nullptr, h->planes.p, h->planes32.p);
            h->planes_valid = true;
        }
    } else {
        // sharded: only the seeds within two grid cells of the owned Morton range get lists and bisector rows
        const u32 nown0 = h->qend() - h->qbegin();
        h->cellflag.ensure((size_t)h->g.ncells / 4 + 1); h->has_planes.ensure(S); h->need_list.ensure(S); h->need_n.ensure(1);
        h->iota.ensure(S);
        if (h->iota_filled < h->iota.cap) {
            LAUNCH(h, iota_u32_kernel, 1024, 256, 0, h->iota.p, h->iota.cap);
            h->iota_filled = h->iota.cap;
        }
        CUDA_CHECK(cudaMemsetAsync(h->cellflag.p, 0, sizeof(u32) * ((size_t)h->g.ncells / 4 + 1), h->stream));
        if (nown0 > 0) LAUNCH(h, mark_cells_kernel<D>, div_up(nown0, 256), 256, 0, h->keys2.p, h->xs.p, h->qbegin(), h->qend(), h->g, h->cellflag.p);
        LAUNCH(h, need_flags_kernel, div_up(S, 256), 256, 0, h->keys2.p, S, h->cellflag.p, h->has_planes.p);
        size_t sel_bytes = 0;
        cub::DeviceSelect::Flagged(nullptr, sel_bytes, h->iota.p, h->has_planes.p, h->need_list.p, h->need_n.p, (int)S, h->stream);
        h->sort_tmp.ensure(sel_bytes);
        CUDA_CHECK(cub::DeviceSelect::Flagged(h->sort_tmp.p, sel_bytes, h->iota.p, h->has_planes.p, h->need_list.p, h->need_n.p, (int)S, h->stream));
        h->launches += 2;
        h->k = 20; h->kstride = 20;
        h->nbr.ensure((size_t)S * h->kstride); h->nbr_n.ensure(S); h->nbr_prev.ensure((size_t)S * 20);
        KnnArgs a;
        memset(&a, 0, sizeof(a));
        a.xs = h->xs.p; a.cell_range = h->cell_range.p; a.rank_of = h->rank_of.p;
        a.query_list = h->need_list.p; a.nq_dev = h->need_n.p; a.ksize = nullptr; a.out_by_slot = 0;
        a.k = 20; a.kstride = 20; a.S = S; a.qbegin = 0; a.qend = S;
        a.nbr = h->nbr.p; a.nbr_n = h->nbr_n.p; a.sqd = nullptr; a.flags = h->flags.p; a.g = h->g;
        // previous lists are kept per original seed; seeds that were outside the halo last time hold stale but
        // still valid bounds (any 20 other seeds bound the 20th distance), so they stay usable once written
        if (!h->prev_valid) CUDA_CHECK(cudaMemsetAsync(h->nbr_prev.p, 0xff, sizeof(u32) * (size_t)S * 20, h->stream));
        a.prev_in = h->nbr_prev.p; a.prev_out = h->nbr_prev.p; a.prev_stride = 20;
        a.planes = h->planes.p; a.planes32 = h->planes32.p;
        launch_knn<D>(h, a, S);
        h->prev_valid = true; h->knn_valid = true; h->planes_valid = true;
    }
    nvtxRangePop();
    CUDA_CHECK(cudaEventRecord(h->ev[2], h->stream));
    h->stats.ensure(16);
    if (h->want_stats) CUDA_CHECK(cudaMemsetAsync(h->stats.p, 0, 16 * sizeof(unsigned long long), h->stream));
    if (h->volumetric) {
        if (D == 3) evaluate_volume(h, mode, check_SR);
        return;
    }
    { NvtxRange r("b200cvt:facet walk (pairs)"); run_pairs_t<D, 3>(h); }
    CUDA_CHECK(cudaEventRecord(h->ev[3], h->stream));
    NvtxRange nvtx_clip("b200cvt:clip+integrate");
    if (mode == 2) {            // candidate rows only (RDT extraction)
        CUDA_CHECK(cudaEventRecord(h->ev[4], h->stream));
        return;
    }

    h->out_s.ensure(S); h->out_v.ensure((size_t)S * D);
    h->redo_a.ensure(S); h->redo_b.ensure(S); h->redo_n.ensure(4);
    CUDA_CHECK(cudaMemsetAsync(h->redo_n.p, 0, 4 * sizeof(u32), h->stream));
    const u32 nown = h->qend() - h->qbegin();
    const u32 np = h->npairs;
    h->flat_seed.ensure(np); h->flat_facet.ensure(np); h->flat_mask.ensure(np);
    h->pstat.ensure(np); h->slow_list.ensure(nown);
    h->contrib.ensure((size_t)np * (D + 1));
    const size_t cstride = h->contrib.cap / (D + 1);
    if (nown > 0) {
        CompactArgs ca;
        ca.pair_cnt = h->pair_cnt.p; ca.pair_facet = h->pair_facet.p; ca.pair_mask = h->pair_mask.p; ca.cap = h->pair_cap;
        ca.pair_off = h->pair_off.p; ca.qbegin = h->qbegin(); ca.nown = nown;
        ca.flat_seed = h->flat_seed.p; ca.flat_facet = h->flat_facet.p; ca.flat_mask = h->flat_mask.p;
        LAUNCH(h, compact_pairs_kernel, std::min<u32>(div_up(nown, 8), (u32)h->num_sms * COMPACT_BLOCKS_PER_SM), 256, 0, ca);
    }
    if (np > 0) {
        ClipFlatArgs fa;
        memset(&fa, 0, sizeof(fa));
        fa.xs = h->xs.p; fa.nbr = h->nbr.p; fa.nbr_n = h->nbr_n.p; fa.kstride = h->kstride; fa.planes = h->planes.p;
        fa.tri = h->tri.p; fa.triw = h->weighted ? h->triw.p : nullptr; fa.facet_area = h->facet_area.p;
        fa.flat_seed = h->flat_seed.p; fa.flat_facet = h->flat_facet.p; fa.npairs_dev = h->pair_off.p + nown;
        fa.mode = mode; fa.contrib = h->contrib.p; fa.cstride = cstride; fa.pstat = h->pstat.p;
        fa.flat_mask = h->flat_mask.p;
        fa.planes32 = h->planes32.p; fa.vmax2 = h->mesh_vmax2;

        fa.stats = h->want_stats ? h->stats.p : nullptr;
        const int VW = D + (h->weighted ? 1 : 0);
        const size_t smem = (size_t)CLIPW_NW * CLIPF_CAP * 32 * (VW * sizeof(double) + CLIPF_QW(D) * sizeof(float4)) +
                            CLIPW_W * sizeof(unsigned short) + CLIPW_NCNT * sizeof(u32);
        // windows of the seed-major pair list, a few per resident block
        const u32 per_sm = std::max<u32>(1, std::min<u32>(8, (u32)((227 * 1024) / (smem + 1024))));
        const u32 blocks = std::min<u32>(div_up(np, CLIPW_W), (u32)h->num_sms * per_sm);
        CUDA_CHECK(cudaEventRecord(h->evk[0], h->stream));
        if (h->weighted) {
            CUDA_CHECK(cudaFuncSetAttribute(clip_win_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            LAUNCH(h, (clip_win_kernel<D, true>), blocks, CLIPW_THREADS, smem, fa);
        } else {
            CUDA_CHECK(cudaFuncSetAttribute(clip_win_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            LAUNCH(h, (clip_win_kernel<D, false>), blocks, CLIPW_THREADS, smem, fa);
        }
        CUDA_CHECK(cudaEventRecord(h->evk[1], h->stream));
    }
    ClipArgs c;
    memset(&c, 0, sizeof(c));
    c.xs = h->xs.p; c.nbr = h->nbr.p; c.nbr_n = h->nbr_n.p; c.kstride = h->kstride; c.nbr_by_slot = 0;
    c.tri = h->tri.p; c.triw = h->weighted ? h->triw.p : nullptr;
    c.pair_cnt = h->pair_cnt.p; c.pair_facet = h->pair_facet.p; c.cap = h->pair_cap;
    c.seed_list = nullptr; c.nseeds = nown; c.qbegin = h->qbegin();
    c.mode = mode; c.check_SR = check_SR; c.S = S;
    c.out_s = h->out_s.p; c.out_v = h->out_v.p; c.flags = h->flags.p;
    c.redo_list = check_SR ? h->redo_a.p : nullptr; c.redo_n = h->redo_n.p;
    c.stats = h->want_stats ? h->stats.p : nullptr;
    if (nown > 0) {
        ReduceArgs ra;
        ra.pair_off = h->pair_off.p; ra.qbegin = h->qbegin(); ra.nown = nown;
        ra.contrib = h->contrib.p; ra.cstride = cstride; ra.pstat = h->pstat.p;
        ra.nbr_n = h->nbr_n.p; ra.kstride = h->kstride; ra.check_SR = check_SR; ra.S = S;
        ra.out_s = h->out_s.p; ra.out_v = h->out_v.p; ra.flags = h->flags.p;
        ra.slow_list = h->slow_list.p; ra.slow_n = h->redo_n.p + 2;
        ra.redo_list = h->redo_a.p; ra.redo_n = h->redo_n.p;
        LAUNCH(h, reduce_pairs_kernel<D>, div_up(nown, 32), 256, 0, ra);
        // seeds with a pair the in-place fast path gave up on: warp-per-seed kernel, same neighbour table
        ClipArgs sl = c;
        sl.seed_list = h->slow_list.p; sl.nseeds = 0; sl.nseeds_dev = h->redo_n.p + 2;
        launch_clip<D>(h, sl);
    }

    if (check_SR) {
        if (h->defer_redo) {
            // Newton loop on one GPU: the count of seeds that need longer lists is read together with the line-search record
            // (lbfgs_post_eval_kernel leaves the search untouched while it is non-zero), one host round trip less per evaluation
            h->pending_clip = c; h->redo_deferred = true;
        } else surface_redo_loop<D>(h, c);
    }
    CUDA_CHECK(cudaEventRecord(h->ev[4], h->stream));
    h->has_results = true;
    h->has_energy = (mode == 1);
    h->ev_valid = true;
    h->ev_pending = true;
}

static void evaluate(b200cvt_ctx* h, int mode, int check_SR) {
    if (h->dim == 3) evaluate_t<3>(h, mode, check_SR); else evaluate_t<6>(h, mode, check_SR);
}

static void scatter_results(b200cvt_ctx* h, bool want_s, bool want_v, bool zero_locked) {
    const u32 S = h->S;
    h->s_orig.ensure(S); h->v_orig.ensure((size_t)S * h->dim); h->flags_orig.ensure(S); h->cnt_orig.ensure(S);
    u32 n = h->qend() - h->qbegin();
    if (n == 0) return;
    if (h->dim == 3)
        LAUNCH(h, scatter_results_kernel<3>, div_up(n, 256), 256, 0, (const SeedRec<3>*)h->xs.p, h->qbegin(), h->qend(),
               h->out_s.p, h->out_v.p, h->flags.p, h->pair_cnt.p, h->locked.p, zero_locked ? 1 : 0,
               want_s ? h->s_orig.p : nullptr, want_v ? h->v_orig.p : nullptr, h->flags_orig.p, h->cnt_orig.p);
    else
        LAUNCH(h, scatter_results_kernel<6>, div_up(n, 256), 256, 0, (const SeedRec<6>*)h->xs.p, h->qbegin(), h->qend(),
               h->out_s.p, h->out_v.p, h->flags.p, h->pair_cnt.p, h->locked.p, zero_locked ? 1 : 0,
               want_s ? h->s_orig.p : nullptr, want_v ? h->v_orig.p : nullptr, h->flags_orig.p, h->cnt_orig.p);
}

static void upload_locked(b200cvt_ctx* h, const uint8_t* locked, u32 S) {
    if (locked) {
        h->locked.ensure(S);
        CUDA_CHECK(cudaMemcpyAsync(h->locked.p, locked, S, cudaMemcpyHostToDevice, h->stream));
    } else if (h->locked.p) {
        // a handle reused with more seeds: the kernels read locked[o] for every o < S
        h->locked.ensure(S);
        CUDA_CHECK(cudaMemsetAsync(h->locked.p, 0, h->locked.cap, h->stream));
    }
}

static void set_seeds_common(b200cvt_ctx* h, u32 S) {
    h->rdt_valid = false; h->rdt_valid_mn = false;
    if (S != h->S) { h->pair_cap = 0; h->prev_valid = false; }
    h->S = S;
    h->has_seeds = true;
    h->grid_valid = false; h->knn_valid = false; h->has_results = false;
}

static void run_exchange(b200cvt_ctx* h) {
    if (h->has_comm) {
        // one all-gather of every rank's chunk over NVLink, ordered on the handle's stream (no host synchronisation)
        NCCL_CHECK(nccl_api().AllGather(h->x_slice, h->x_all, (size_t)h->slice_len() * (h->dim + 1), ncclDouble, h->nccl, h->stream));
        h->exchanges++;
        return;
    }
    if (!h->xcb || !h->x_slice || !h->x_all) throw StateError("seeds are partitioned but no exchange was set (b200cvt_set_exchange)");
    if (h->x_chunk < (u64)h->slice_len() * (h->dim + 1)) throw ArgError("exchange buffers too small");
    // the callback enqueues the all-gather on the handle's stream when the caller supplied it (b200cvt_set_stream);
    // with a private stream the library has to drain it first and the callback must return with the data in place
    if (h->own_stream) sync_stream(h);
    if (h->xcb(h->xuser) != 0) throw std::runtime_error("exchange callback failed");
    h->exchanges++;
}

template <int D>
static void pack_slice(b200cvt_ctx* h, const double* out_s, const double* out_v) {
    u32 n = h->slice_len();
    LAUNCH(h, pack_slice_kernel<D>, div_up(n, 256), 256, 0, out_s, out_v, h->qbegin(), h->qend(), n, h->x_slice);
}

// A cancel request of the progress callback must stop every rank at the same iteration (the next collective would hang
// otherwise). Group members agree through the shared flag; ranks of different processes must be given callbacks that
// return the same value.
static bool agree_cancel(b200cvt_ctx* h, bool mine) {
    if (!h->group) return mine;
    if (mine) { std::lock_guard<std::mutex> lk(h->group->mu); h->group->cancel = 1; }
    h->group->barrier();
    bool all;
    { std::lock_guard<std::mutex> lk(h->group->mu); all = h->group->cancel != 0; }
    h->group->barrier();
    return all;
}

// communicator mode: the exchange buffers belong to the library
static void comm_prepare(b200cvt_ctx* h) {
    if (!h->has_comm || h->nranks <= 1) return;
    const size_t chunk = (size_t)h->slice_len() * (h->dim + 1);
    h->xch_slice.ensure(chunk); h->xch_all.ensure(chunk * h->nranks);
    h->x_slice = h->xch_slice.p; h->x_all = h->xch_all.p; h->x_chunk = chunk;
}

// Lloyd_iterations (geogram/voronoi/CVT.cpp:133-167) on the device-resident seeds
static void lloyd_loop(b200cvt_ctx* h, u32 nb_iter, b200cvt_progress_cb cb, void* user) {
    const u32 S = h->S;
    if (nb_iter > 0) h->rdt_valid = false; h->rdt_valid_mn = false;      // the seeds move: a cached triangulation is stale
    comm_prepare(h);
    for (u32 it = 0; it < nb_iter; ++it) {
        NvtxRange nvtx_it("b200cvt:Lloyd iteration");
        evaluate(h, 0, 0);
        u32 n = h->slice_len();
        const uint8_t* lk = h->locked.p;
        if (h->nranks == 1) {
            if (h->dim == 3)
                LAUNCH(h, lloyd_update_kernel<3>, div_up(n, 256), 256, 0, (const SeedRec<3>*)h->xs.p, h->out_s.p, h->out_v.p, lk,
                       h->qbegin(), h->qend(), h->x.p, (double*)nullptr, n);
            else
                LAUNCH(h, lloyd_update_kernel<6>, div_up(n, 256), 256, 0, (const SeedRec<6>*)h->xs.p, h->out_s.p, h->out_v.p, lk,
                       h->qbegin(), h->qend(), h->x.p, (double*)nullptr, n);
        } else {
            // updated owned slice, sorted order -> all-gather -> every rank rebuilds the full seed array
            if (!h->x_slice) throw StateError("seeds are partitioned but no exchange was set (b200cvt_set_exchange)");
            if (h->dim == 3) {
                LAUNCH(h, lloyd_update_kernel<3>, div_up(n, 256), 256, 0, (const SeedRec<3>*)h->xs.p, h->out_s.p, h->out_v.p, lk,
                       h->qbegin(), h->qend(), (double*)nullptr, h->x_slice, n);
                run_exchange(h);
                LAUNCH(h, unpack_all_kernel<3>, div_up(S, 256), 256, 0, (const SeedRec<3>*)h->xs.p, h->x_all, S, n,
                       (const uint8_t*)nullptr, 0, h->x.p, (double*)nullptr);
            } else {
                LAUNCH(h, lloyd_update_kernel<6>, div_up(n, 256), 256, 0, (const SeedRec<6>*)h->xs.p, h->out_s.p, h->out_v.p, lk,
                       h->qbegin(), h->qend(), (double*)nullptr, h->x_slice, n);
                run_exchange(h);
                LAUNCH(h, unpack_all_kernel<6>, div_up(S, 256), 256, 0, (const SeedRec<6>*)h->xs.p, h->x_all, S, n,
                       (const uint8_t*)nullptr, 0, h->x.p, (double*)nullptr);
            }
        }
        CUDA_CHECK(cudaEventRecord(h->ev[5], h->stream));
        if (it + 1 == nb_iter) scatter_results(h, false, false, false);
        h->grid_valid = false; h->knn_valid = false;
        sync_stream(h);
        if (agree_cancel(h, cb && cb(user, it + 1, 0.0, 0.0))) throw CanceledError("canceled by the progress callback");
    }
    h->has_results = nb_iter > 0;
    h->has_energy = false;
}

// ---------------------------------------------------------------------------------------
// restricted Delaunay triangulation, simple mode (rdt.cuh)
// ---------------------------------------------------------------------------------------
// facet_corners.adjacent_facet of the device copy (sorted facet ids). Taken from the caller's adjacency when it was given
// to b200cvt_set_mesh, else rebuilt as MeshFacets::connect() does for manifold edges: two facets that share an edge.
static void ensure_facet_adj(b200cvt_ctx* h) {
    if (h->facet_adj_valid) return;
    const u32 T = h->T;
    std::vector<int32_t> adj;
    if (!h->host_adj.empty()) adj = h->host_adj;
    else {
        adj.assign((size_t)T * 3, -1);
        struct Edge { u32 a, b, f, c; };
        std::vector<Edge> E((size_t)T * 3);
        for (u32 f = 0; f < T; ++f)
            for (u32 c = 0; c < 3; ++c) {
                u32 v0 = h->host_elems[(size_t)f * 3 + c], v1 = h->host_elems[(size_t)f * 3 + (c + 1) % 3];
                E[(size_t)f * 3 + c] = Edge{std::min(v0, v1), std::max(v0, v1), f, c};
            }
        std::sort(E.begin(), E.end(), [](const Edge& x, const Edge& y) {
            if (x.a != y.a) return x.a < y.a;
            if (x.b != y.b) return x.b < y.b;
            return x.f < y.f;
        });
        for (size_t i = 0; i < E.size();) {
            size_t j = i + 1;
            while (j < E.size() && E[j].a == E[i].a && E[j].b == E[i].b) ++j;
            if (j - i == 2) {
                adj[(size_t)E[i].f * 3 + E[i].c] = (int32_t)E[i + 1].f;
                adj[(size_t)E[i + 1].f * 3 + E[i + 1].c] = (int32_t)E[i].f;
            }
            i = j;
        }
    }
    std::vector<u32> inv(T);
    for (u32 i = 0; i < T; ++i) inv[h->host_perm[i]] = i;
    std::vector<int> sorted((size_t)T * 3);
    for (u32 i = 0; i < T; ++i)
        for (u32 c = 0; c < 3; ++c) {
            const int32_t a = adj[(size_t)h->host_perm[i] * 3 + c];
            sorted[(size_t)i * 3 + c] = (a >= 0 && (u32)a < T) ? (int)inv[(u32)a] : -1;
        }
    h->facet_adj.ensure(std::max<size_t>(sorted.size(), 1));
    if (!sorted.empty())
        CUDA_CHECK(cudaMemcpyAsync(h->facet_adj.p, sorted.data(), sizeof(int) * sorted.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    h->facet_adj_valid = true;
}

template <int D>
static void launch_rdt(b200cvt_ctx* h, RdtArgs& a) {
    if (a.nseeds == 0) return;
    const size_t smem = (size_t)CLIP_WARPS * a.kstride * (D + 3) * sizeof(double);
    if (smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(rdt_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(h, rdt_kernel<D>, div_up(a.nseeds, CLIP_WARPS), CLIP_WARPS * 32, smem, a);
}

// RestrictedVoronoiDiagram::compute_RDT(simplices, embedding, RDTMode(0)) for the owned seeds (RVD.cpp:2353-2370):
// check_SR = true as CentroidalVoronoiTesselation::compute_surface sets it (CVT.cpp:194). Result in h->rdt_host,
// rows sorted lexicographically.
// ---------------------------------------------------------------------------------------
// volumetric RDT: RestrictedVoronoiDiagram::compute_RDT with set_volumetric(true) (RVD.cpp:2308-2335)
// ---------------------------------------------------------------------------------------
namespace {
inline void two_sum(double a, double b, double& s, double& e) { s = a + b; const double bb = s - a; e = (a - (s - bb)) + (b - bb); }
inline void two_prod(double a, double b, double& p, double& e) { p = a * b; e = std::fma(a, b, -p); }
// exact sum of doubles as a non-overlapping expansion (Shewchuk's Grow-Expansion), components by increasing magnitude
inline void grow(std::vector<double>& ex, double b) {
    double q = b;
    size_t w = 0;
    for (size_t i = 0; i < ex.size(); ++i) { double sum, r; two_sum(q, ex[i], sum, r); if (r != 0.0) ex[w++] = r; q = sum; }
    ex.resize(w);
    ex.push_back(q);
}
// sign of det(p1 - p0, p2 - p0, p3 - p0) = PCK::orient_3d (numerics/predicates/orient3d.h:4-24; exact fallback as the reference)
int orient3d_sign(const double* p0, const double* p1, const double* p2, const double* p3) {
    double a[3][3];
    for (int c = 0; c < 3; ++c) { a[0][c] = p1[c] - p0[c]; a[1][c] = p2[c] - p0[c]; a[2][c] = p3[c] - p0[c]; }
    const double m0 = a[1][1] * a[2][2] - a[1][2] * a[2][1], m1 = a[0][1] * a[2][2] - a[0][2] * a[2][1], m2 = a[0][1] * a[1][2] - a[0][2] * a[1][1];
    const double det = a[0][0] * m0 - a[1][0] * m1 + a[2][0] * m2;
    const double perm = std::fabs(a[0][0]) * (std::fabs(a[1][1] * a[2][2]) + std::fabs(a[1][2] * a[2][1])) +
                        std::fabs(a[1][0]) * (std::fabs(a[0][1] * a[2][2]) + std::fabs(a[0][2] * a[2][1])) +
                        std::fabs(a[2][0]) * (std::fabs(a[0][1] * a[1][2]) + std::fabs(a[0][2] * a[1][1]));
    if (std::fabs(det) > 1e-13 * perm) return det > 0.0 ? 1 : -1;
    // exact: every difference as a two-term expansion, every triple product expanded, one exact sum
    double hi[3][3], lo[3][3];
    const double* P[3] = {p1, p2, p3};
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) two_sum(P[r][c], -p0[c], hi[r][c], lo[r][c]);
    std::vector<double> ex;
    static const int perms[6][4] = {{0, 1, 2, 1}, {1, 2, 0, 1}, {2, 0, 1, 1}, {0, 2, 1, -1}, {1, 0, 2, -1}, {2, 1, 0, -1}};
    for (int q = 0; q < 6; ++q) {
        const int c0 = perms[q][0], c1 = perms[q][1], c2 = perms[q][2];
        const double sg = (double)perms[q][3];
        for (int m = 0; m < 8; ++m) {
            const double x = (m & 1) ? lo[0][c0] : hi[0][c0], y = (m & 2) ? lo[1][c1] : hi[1][c1], z = (m & 4) ? lo[2][c2] : hi[2][c2];
            if (x == 0.0 || y == 0.0 || z == 0.0) continue;
            double p, e;
            two_prod(x, y, p, e);
            double t0, t1, t2, t3;
            two_prod(p, z, t0, t1);
            two_prod(e, z, t2, t3);
            grow(ex, sg * t3); grow(ex, sg * t1); grow(ex, sg * t2); grow(ex, sg * t0);
        }
    }
    for (size_t i = ex.size(); i-- > 0;) if (ex[i] != 0.0) return ex[i] > 0.0 ? 1 : -1;
    return 0;
}
}

// The Delaunay tets whose Voronoi vertex lies inside the tetrahedralised domain, each once: rows of four original seed indices,
// ascending within a row except for the first two (swapped where orient_3d asks for it, as the reference does), rows sorted.
// Interior cells emit from vcell_kernel, the cells near the boundary from clip_tet_kernel (one row per piece vertex on three
// bisectors); check_SR = true as CentroidalVoronoiTesselation::compute_volume sets it (CVT.cpp:245).
static void compute_rdt_volume(b200cvt_ctx* h) {
    if (!h->has_mesh) throw StateError("no mesh: call b200cvt_set_mesh first");
    if (!h->has_seeds) throw StateError("no seeds: call b200cvt_set_seeds first");
    if (h->nranks > 1) throw ArgError("compute_RDT on a partitioned handle: gather the seeds on one handle first");
    const u32 S = h->S;
    h->rdtv_n.ensure(1);
    size_t cap = std::max<size_t>(h->rdtv.cap, (size_t)8 * S + 1024);
    unsigned long long n = 0;
    for (int attempt = 0; attempt < 4; ++attempt) {
        h->rdtv.ensure(cap);
        CUDA_CHECK(cudaMemsetAsync(h->rdtv_n.p, 0, sizeof(unsigned long long), h->stream));
        evaluate_t<3>(h, 3, 1);
        CUDA_CHECK(cudaMemcpyAsync(&n, h->rdtv_n.p, sizeof(n), cudaMemcpyDeviceToHost, h->stream));
        sync_stream(h);
        if (n <= h->rdtv.cap) break;
        cap = (size_t)n + n / 8 + 1024;
        if (attempt == 3) throw CapacityError("restricted Delaunay tets keep overflowing their buffer");
    }
    std::vector<uint4> rows((size_t)n);
    if (n > 0) CUDA_CHECK(cudaMemcpy(rows.data(), h->rdtv.p, sizeof(uint4) * (size_t)n, cudaMemcpyDeviceToHost));
    std::vector<double> xh((size_t)S * 3);
    CUDA_CHECK(cudaMemcpy(xh.data(), h->x.p, sizeof(double) * (size_t)S * 3, cudaMemcpyDeviceToHost));
    struct T4 { u32 v[4]; };
    std::vector<T4> t((size_t)n);
    for (size_t i = 0; i < (size_t)n; ++i) {
        u32 q[4] = {rows[i].x, rows[i].y, rows[i].z, rows[i].w};
        std::sort(q, q + 4);
        memcpy(t[i].v, q, sizeof(q));
    }
    auto less4 = [](const T4& a, const T4& b) { return std::lexicographical_compare(a.v, a.v + 4, b.v, b.v + 4); };
    std::sort(t.begin(), t.end(), less4);
    // a piece that passed its radius test in one pass and was computed again with a longer list emits its rows twice
    t.erase(std::unique(t.begin(), t.end(), [](const T4& a, const T4& b) { return memcmp(a.v, b.v, sizeof(a.v)) == 0; }), t.end());
    for (T4& r : t)
        if (orient3d_sign(&xh[(size_t)r.v[0] * 3], &xh[(size_t)r.v[1] * 3], &xh[(size_t)r.v[2] * 3], &xh[(size_t)r.v[3] * 3]) < 0) std::swap(r.v[0], r.v[1]);
    h->rdt_host.resize(t.size() * 4);
    if (!t.empty()) memcpy(h->rdt_host.data(), t.data(), sizeof(T4) * t.size());
    h->rdt_valid = true;
}

template <int D>
static void compute_rdt_t(b200cvt_ctx* h) {
    if (h->volumetric) throw ArgError("compute_RDT of a volumetric diagram stays on the reference implementation");
    if (!h->has_mesh) throw StateError("no mesh: call b200cvt_set_mesh first");
    if (!h->has_seeds) throw StateError("no seeds: call b200cvt_set_seeds first");
    ensure_facet_adj(h);
    evaluate_t<D>(h, 2, 1);
    const u32 S = h->S;
    const u32 nown = h->qend() - h->qbegin();
    h->redo_a.ensure(S); h->redo_b.ensure(S); h->redo_n.ensure(4); h->rdt_n.ensure(1);
    size_t out_cap = std::max<size_t>(h->rdt_dev.cap / 3, (size_t)3 * nown + 1024);
    for (int attempt = 0; attempt < 6; ++attempt) {
        h->rdt_dev.ensure(out_cap * 3);
        out_cap = std::min<size_t>(h->rdt_dev.cap / 3, 0xffffffffu);
        CUDA_CHECK(cudaMemsetAsync(h->redo_n.p, 0, 4 * sizeof(u32), h->stream));
        CUDA_CHECK(cudaMemsetAsync(h->rdt_n.p, 0, sizeof(u32), h->stream));
        RdtArgs r;
        memset(&r, 0, sizeof(r));
        r.xs = h->xs.p; r.nbr = h->nbr.p; r.nbr_n = h->nbr_n.p; r.kstride = h->kstride; r.nbr_by_slot = 0;
        r.tri = h->tri.p; r.facet_adj = h->facet_adj.p; r.T = h->T;
        r.pair_cnt = h->pair_cnt.p; r.pair_facet = h->pair_facet.p; r.cap = h->pair_cap;
        r.seed_list = nullptr; r.nseeds = nown; r.qbegin = h->qbegin(); r.S = S; r.flags = h->flags.p;
        r.redo_list = h->redo_a.p; r.redo_n = h->redo_n.p;
        r.out_tri = h->rdt_dev.p; r.out_cap = (u32)out_cap; r.out_n = h->rdt_n.p;
        launch_rdt<D>(h, r);
        // enlarge_neighborhood (generic_RVD.h:2183-2197), batched over the seeds that need it
        u32 kbig = 40;
        u32* cur_list = h->redo_a.p; u32* nxt_list = h->redo_b.p;
        int cur_slot = 0;
        for (;;) {
            u32 nredo = 0;
            CUDA_CHECK(cudaMemcpyAsync(&nredo, h->redo_n.p + cur_slot, sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
            if (nredo == 0) break;
            kbig = std::min<u32>(std::min<u32>(kbig, B200CVT_KMAX), S - 1);
            h->nbr_big.ensure((size_t)nredo * kbig);
            h->nbr_big_n.ensure(nredo);
            KnnArgs a;
            memset(&a, 0, sizeof(a));
            a.xs = h->xs.p; a.cell_range = h->cell_range.p; a.rank_of = h->rank_of.p;
            a.query_list = cur_list; a.ksize = nullptr; a.out_by_slot = 1;
            a.k = kbig; a.kstride = kbig; a.S = S; a.qbegin = 0; a.qend = nredo;
            a.nbr = h->nbr_big.p; a.nbr_n = h->nbr_big_n.p; a.sqd = nullptr; a.flags = h->flags.p; a.g = h->g;
            launch_knn<D>(h, a, nredo);
            const int nslot = cur_slot ^ 1;
            CUDA_CHECK(cudaMemsetAsync(h->redo_n.p + nslot, 0, sizeof(u32), h->stream));
            RdtArgs rr = r;
            rr.nbr = h->nbr_big.p; rr.nbr_n = h->nbr_big_n.p; rr.kstride = kbig; rr.nbr_by_slot = 1;
            rr.seed_list = cur_list; rr.nseeds = nredo;
            rr.redo_list = nxt_list; rr.redo_n = h->redo_n.p + nslot;
            launch_rdt<D>(h, rr);
            std::swap(cur_list, nxt_list);
            cur_slot = nslot;
            if (kbig >= std::min<u32>(B200CVT_KMAX, S - 1)) break;
            kbig *= 2;
        }
        u32 n = 0;
        CUDA_CHECK(cudaMemcpyAsync(&n, h->rdt_n.p, sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (n > out_cap) { out_cap = (size_t)n + n / 8 + 1024; continue; }
        h->rdt_host.resize((size_t)n * 3);
        if (n > 0) CUDA_CHECK(cudaMemcpy(h->rdt_host.data(), h->rdt_dev.p, sizeof(u32) * 3 * (size_t)n, cudaMemcpyDeviceToHost));
        struct T3 { u32 a, b, c; };
        T3* t = reinterpret_cast<T3*>(h->rdt_host.data());
        std::sort(t, t + n, [](const T3& x, const T3& y) {
            if (x.a != y.a) return x.a < y.a;
            if (x.b != y.b) return x.b < y.b;
            return x.c < y.c;
        });
        h->rdt_valid = true;
        // only the status flags (polygon overflow, neighbourhood cap) are results of this pass
        scatter_results(h, false, false, false);
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        h->has_results = true; h->has_energy = false;
        return;
    }
    throw CapacityError("RDT triangle list keeps overflowing");
}

// RestrictedVoronoiDiagram::compute_RDT with RDT_MULTINERVE (RVD.cpp:2338-2352) on the device (rdt_mn.cuh).
// Result in h->mn_tri_host (rows sorted, duplicates removed), h->mn_emb_host, h->mn_vseed_host.
template <int D>
static void compute_rdt_mn_t(b200cvt_ctx* h, int use_centroids, int prefer_seeds, const uint8_t* locked) {
    if (h->volumetric) throw ArgError("compute_RDT of a volumetric diagram stays on the reference implementation");
    if (h->nranks != 1) throw StateError("the multinerve RDT needs the whole seed range (nranks == 1)");
    if (!h->has_mesh) throw StateError("no mesh: call b200cvt_set_mesh first");
    if (!h->has_seeds) throw StateError("no seeds: call b200cvt_set_seeds first");
    ensure_facet_adj(h);
    evaluate_t<D>(h, 2, 1);
    const u32 S = h->S;
    const u32 cap = h->pair_cap;
    upload_locked(h, locked, S);
    h->redo_a.ensure(S); h->redo_b.ensure(S); h->redo_n.ensure(4);
    h->mn_pair_comp.ensure((size_t)S * cap);
    h->mn_ncomp.ensure((size_t)S + 1); h->mn_comp_base.ensure((size_t)S + 2);
    h->mn_comp_m.ensure((size_t)S * MN_MAXC); h->mn_comp_mg.ensure((size_t)S * MN_MAXC * D); h->mn_comp_border.ensure((size_t)S * MN_MAXC);
    h->mn_vert_n.ensure(1); h->mn_tri_n.ensure(1);
    size_t vcap = std::max<size_t>(h->mn_vert.cap, (size_t)8 * S + 1024);
    for (int attempt = 0; attempt < 6; ++attempt) {
        h->mn_vert.ensure(vcap);
        vcap = std::min<size_t>(h->mn_vert.cap, 0xffffffffu);
        CUDA_CHECK(cudaMemsetAsync(h->redo_n.p, 0, 4 * sizeof(u32), h->stream));
        CUDA_CHECK(cudaMemsetAsync(h->mn_vert_n.p, 0, sizeof(u32), h->stream));
        CUDA_CHECK(cudaMemsetAsync(h->mn_ncomp.p, 0, sizeof(u32) * ((size_t)S + 1), h->stream));
        CUDA_CHECK(cudaMemsetAsync(h->mn_pair_comp.p, 0xff, (size_t)S * cap, h->stream));
        RdtMnArgs m;
        memset(&m, 0, sizeof(m));
        RdtArgs& r = m.r;
        r.xs = h->xs.p; r.nbr = h->nbr.p; r.nbr_n = h->nbr_n.p; r.kstride = h->kstride; r.nbr_by_slot = 0;
        r.tri = h->tri.p; r.facet_adj = h->facet_adj.p; r.T = h->T;
        r.pair_cnt = h->pair_cnt.p; r.pair_facet = h->pair_facet.p; r.cap = cap;
        r.seed_list = nullptr; r.nseeds = S; r.qbegin = 0; r.S = S; r.flags = h->flags.p;
        r.redo_list = h->redo_a.p; r.redo_n = h->redo_n.p;
        m.pair_comp = h->mn_pair_comp.p; m.ncomp = h->mn_ncomp.p; m.comp_m = h->mn_comp_m.p; m.comp_mg = h->mn_comp_mg.p;
        m.comp_border = h->mn_comp_border.p; m.vert = h->mn_vert.p; m.vert_cap = (u32)vcap; m.vert_n = h->mn_vert_n.p;
        auto launch = [&](RdtMnArgs& q) {
            if (q.r.nseeds == 0) return;
            const size_t smem = (size_t)CLIP_WARPS * mn_warp_doubles<D>(q.r.kstride, q.r.cap) * sizeof(double);
            if (smem > 200 * 1024) throw CapacityError("candidate rows too long for the multinerve RDT kernel");
            if (smem > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(rdt_mn_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            LAUNCH(h, rdt_mn_kernel<D>, div_up(q.r.nseeds, CLIP_WARPS), CLIP_WARPS * 32, smem, q);
        };
        launch(m);
        // enlarge_neighborhood (generic_RVD.h:2183-2197), batched over the seeds that need it
        u32 kbig = 40;
        u32* cur_list = h->redo_a.p; u32* nxt_list = h->redo_b.p;
        int cur_slot = 0;
        for (;;) {
            u32 nredo = 0;
            CUDA_CHECK(cudaMemcpyAsync(&nredo, h->redo_n.p + cur_slot, sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
            if (nredo == 0) break;
            kbig = std::min<u32>(std::min<u32>(kbig, B200CVT_KMAX), S - 1);
            h->nbr_big.ensure((size_t)nredo * kbig);
            h->nbr_big_n.ensure(nredo);
            KnnArgs a;
            memset(&a, 0, sizeof(a));
            a.xs = h->xs.p; a.cell_range = h->cell_range.p; a.rank_of = h->rank_of.p;
            a.query_list = cur_list; a.ksize = nullptr; a.out_by_slot = 1;
            a.k = kbig; a.kstride = kbig; a.S = S; a.qbegin = 0; a.qend = nredo;
            a.nbr = h->nbr_big.p; a.nbr_n = h->nbr_big_n.p; a.sqd = nullptr; a.flags = h->flags.p; a.g = h->g;
            launch_knn<D>(h, a, nredo);
            const int nslot = cur_slot ^ 1;
            CUDA_CHECK(cudaMemsetAsync(h->redo_n.p + nslot, 0, sizeof(u32), h->stream));
            RdtMnArgs mm = m;
            mm.r.nbr = h->nbr_big.p; mm.r.nbr_n = h->nbr_big_n.p; mm.r.kstride = kbig; mm.r.nbr_by_slot = 1;
            mm.r.seed_list = cur_list; mm.r.nseeds = nredo;
            mm.r.redo_list = nxt_list; mm.r.redo_n = h->redo_n.p + nslot;
            launch(mm);
            std::swap(cur_list, nxt_list);
            cur_slot = nslot;
            if (kbig >= std::min<u32>(B200CVT_KMAX, S - 1)) break;
            kbig *= 2;
        }
        u32 nvtx = 0;
        CUDA_CHECK(cudaMemcpyAsync(&nvtx, h->mn_vert_n.p, sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (nvtx > vcap) { vcap = (size_t)nvtx + nvtx / 8 + 1024; continue; }
        // component numbering: by original seed index, then by smallest facet inside the cell
        size_t scan_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, h->mn_ncomp.p, h->mn_comp_base.p, (int)S + 1, h->stream);
        h->cub_tmp.ensure(scan_bytes);
        CUDA_CHECK(cub::DeviceScan::ExclusiveSum(h->cub_tmp.p, scan_bytes, h->mn_ncomp.p, h->mn_comp_base.p, (int)S + 1, h->stream));
        h->launches += 1;
        u32 ncomp_total = 0;
        CUDA_CHECK(cudaMemcpyAsync(&ncomp_total, h->mn_comp_base.p + S, sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        h->mn_tri.ensure(std::max<size_t>((size_t)nvtx * 3, 3));
        h->mn_emb.ensure(std::max<size_t>((size_t)ncomp_total * D, 1)); h->mn_vert_seed.ensure(std::max<size_t>(ncomp_total, 1));
        CUDA_CHECK(cudaMemsetAsync(h->mn_tri_n.p, 0, sizeof(u32), h->stream));
        if (nvtx > 0) {
            MnTriArgs t;
            memset(&t, 0, sizeof(t));
            t.vert = h->mn_vert.p; t.vert_n = h->mn_vert_n.p; t.vert_cap = (u32)vcap; t.xs = h->xs.p; t.rank_of = h->rank_of.p;
            t.pair_cnt = h->pair_cnt.p; t.pair_facet = h->pair_facet.p; t.pair_comp = h->mn_pair_comp.p; t.cap = cap;
            t.comp_base = h->mn_comp_base.p; t.out_tri = h->mn_tri.p; t.out_cap = nvtx; t.out_n = h->mn_tri_n.p;
            LAUNCH(h, mn_triangles_kernel<D>, std::min<u32>(div_up(nvtx, 128), (u32)h->num_sms * 16u), 128, 0, t);
        }
        LAUNCH(h, mn_embedding_kernel<D>, div_up(S, 128), 128, 0, h->xs.p, h->rank_of.p, S, h->mn_ncomp.p, h->mn_comp_base.p,
               h->mn_comp_m.p, h->mn_comp_mg.p, h->mn_comp_border.p, locked ? h->locked.p : (const uint8_t*)nullptr,
               use_centroids, prefer_seeds, h->mn_emb.p, h->mn_vert_seed.p);
        u32 ntri = 0;
        CUDA_CHECK(cudaMemcpyAsync(&ntri, h->mn_tri_n.p, sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        h->mn_tri_host.resize((size_t)ntri * 3);
        h->mn_emb_host.resize((size_t)ncomp_total * D); h->mn_vseed_host.resize(ncomp_total);
        if (ntri) CUDA_CHECK(cudaMemcpy(h->mn_tri_host.data(), h->mn_tri.p, sizeof(u32) * 3 * (size_t)ntri, cudaMemcpyDeviceToHost));
        if (ncomp_total) {
            CUDA_CHECK(cudaMemcpy(h->mn_emb_host.data(), h->mn_emb.p, sizeof(double) * (size_t)ncomp_total * D, cudaMemcpyDeviceToHost));
            CUDA_CHECK(cudaMemcpy(h->mn_vseed_host.data(), h->mn_vert_seed.p, sizeof(u32) * ncomp_total, cudaMemcpyDeviceToHost));
        }
        // every restricted Voronoi vertex is reported by each of its (up to three) cells: sort, drop the copies
        struct T3 { u32 a, b, c; };
        T3* t = reinterpret_cast<T3*>(h->mn_tri_host.data());
        auto less = [](const T3& x, const T3& y) { return x.a != y.a ? x.a < y.a : (x.b != y.b ? x.b < y.b : x.c < y.c); };
        std::sort(t, t + ntri, less);
        T3* e = std::unique(t, t + ntri, [](const T3& x, const T3& y) { return x.a == y.a && x.b == y.b && x.c == y.c; });
        h->mn_tri_host.resize((size_t)(e - t) * 3);
        scatter_results(h, false, false, false);
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        h->has_results = true; h->has_energy = false;
        return;
    }
    throw CapacityError("restricted Voronoi vertex list keeps overflowing");
}

// ---------------------------------------------------------------------------------------
// roofline denominators measured on the box: non-tensor FP32 / FP64 FMA rate
// and a STREAM-style copy
// ---------------------------------------------------------------------------------------
template <class T>
__global__ void fma_peak_kernel(T* out, int iters, T a, T b) {
    T v0 = (T)threadIdx.x, v1 = v0 + (T)1, v2 = v0 + (T)2, v3 = v0 + (T)3, v4 = v0 + (T)4, v5 = v0 + (T)5, v6 = v0 + (T)6, v7 = v0 + (T)7;
    for (int i = 0; i < iters; ++i) {
        v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
        v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
    }
    T r = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
    if (r == (T)123456789) out[0] = r;
}
__global__ void copy_peak_kernel(const double4* __restrict__ in, double4* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
static void measure_peaks(double* fp32_tflops, double* fp64_tflops, double* copy_gbs) {
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
    void* scratch = nullptr;
    CUDA_CHECK(cudaMalloc(&scratch, 64));
    const int blocks = 148 * 8, threads = 256, iters = 20000;
    auto run = [&](auto launch) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            CUDA_CHECK(cudaEventRecord(e0));
            launch();
            CUDA_CHECK(cudaEventRecord(e1));
            CUDA_CHECK(cudaEventSynchronize(e1));
            float ms = 0.f; CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        return best;
    };
    double flops = 2.0 * 8.0 * iters * (double)blocks * threads;
    if (fp32_tflops) {
        float ms = run([&] { fma_peak_kernel<float><<<blocks, threads>>>((float*)scratch, iters, 1.0000001f, 1e-7f); });
        *fp32_tflops = flops / (ms * 1e-3) / 1e12;
    }
    if (fp64_tflops) {
        float ms = run([&] { fma_peak_kernel<double><<<blocks, threads>>>((double*)scratch, iters, 1.0000001, 1e-7); });
        *fp64_tflops = flops / (ms * 1e-3) / 1e12;
    }
    if (copy_gbs) {
        size_t n = (size_t)1 << 25;   // 2 x 1 GiB
        double4 *a = nullptr, *b = nullptr;
        CUDA_CHECK(cudaMalloc((void**)&a, n * sizeof(double4)));
        CUDA_CHECK(cudaMalloc((void**)&b, n * sizeof(double4)));
        CUDA_CHECK(cudaMemset(a, 0, n * sizeof(double4)));
        float ms = run([&] { copy_peak_kernel<<<148 * 16, 512>>>(a, b, n); });
        *copy_gbs = 2.0 * n * sizeof(double4) / (ms * 1e-3) / 1e9;
        cudaFree(a); cudaFree(b);
    }
    CUDA_CHECK(cudaGetLastError());
    cudaFree(scratch);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

// ---------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------
extern "C" {

const char* b200cvt_last_error(void) { return g_last_error.c_str(); }

int b200cvt_create(int device, int dim, int volumetric, b200cvt_handle* out) {
    return guarded([&] {
        if (!out) throw ArgError("out is NULL");
        if (dim != 3 && dim != 6) throw ArgError("dim must be 3 or 6 (other dimensions stay on the reference implementation)");
        if (volumetric && dim != 3) throw ArgError("volumetric mode needs dim 3 (other dimensions stay on the reference implementation)");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) throw CudaError(std::string("no CUDA device: ") + cudaGetErrorString(e));
        if (device < 0) CUDA_CHECK(cudaGetDevice(&device));
        if (device >= ndev) throw ArgError("device ordinal out of range");
        CUDA_CHECK(cudaSetDevice(device));
        std::unique_ptr<b200cvt_ctx> h(new b200cvt_ctx);
        h->device = device; h->dim = dim; h->volumetric = volumetric;
        { const char* e = getenv("B200CVT_VCELL"); h->use_vcell = !(e && atoi(e) == 0); }
        { const char* e = getenv("B200CVT_LBFGS_GRAM"); h->use_lbfgs_gram = !(e && atoi(e) == 0); }
        { const char* e = getenv("B200CVT_VCELL_TET"); h->use_vcell_tet = (e && atoi(e) != 0); }
        CUDA_CHECK(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
        CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
        *out = h.release();
    });
}

static void comm_release(b200cvt_ctx* h);

void b200cvt_destroy(b200cvt_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    comm_release(h);
    for (int i = 0; i < 6; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 2; ++i) if (h->evk[i]) cudaEventDestroy(h->evk[i]);
    for (int i = 0; i < 2; ++i) if (h->evc[i]) cudaEventDestroy(h->evc[i]);
    for (int i = 0; i < 2; ++i) if (h->evl[i]) cudaEventDestroy(h->evl[i]);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;                         // DevBuf members release their device memory
}

int b200cvt_set_mesh(b200cvt_handle h, const double* vertices, uint32_t nv, uint32_t stride, const uint32_t* elems,
                     const int32_t* adjacency, uint32_t ne, const double* weights) {
    return guarded([&] {
        if (!h || !vertices || !elems) throw ArgError("null argument");
        if (stride < (u32)h->dim) throw ArgError("vertex stride smaller than the dimension (geo_assert(dimension_ <= mesh->vertices.dimension()), CVT.cpp:64)");
        CUDA_CHECK(cudaSetDevice(h->device));
        const int D = h->dim;
        const int per = h->volumetric ? 4 : 3;
        // the volumetric actions ignore the "weight" attribute (RVD.cpp:420,783)
        const double* weights_in = weights;
        if (h->volumetric) weights = nullptr;
        // On the device (mesh_prep.cuh): pass 1 validates the indices and reduces the bounding box and the total area /
        // volume; pass 2 sorts the elements in Morton order of their centroids (10 bits per axis, stable: ties keep the
        // caller's order), so that the elements one warp walks share home seeds and bisector rows — the reference reorders
        // the caller's mesh for the same reason (mesh_partition Hilbert sort, RVD.cpp:2390-2395), here only the device copy
        // is permuted; pass 3 gathers the corner coordinates (and weights) in that order.
        h->has_mesh = false;       // a call that fails leaves the handle without a mesh, not with half of the new one
        DevBuf<double> d_vert, d_w;
        DevBuf<u32> d_elems, d_keys, d_keys2, d_vals, d_order;
        DevBuf<MeshPartial> d_part;
        DevBuf<unsigned char> d_tmp;
        struct Scratch {
            DevBuf<double>& a; DevBuf<double>& b; DevBuf<u32>& c; DevBuf<u32>& d; DevBuf<u32>& e; DevBuf<u32>& f; DevBuf<u32>& g;
            DevBuf<MeshPartial>& p; DevBuf<unsigned char>& t;
            ~Scratch() { a.release(); b.release(); c.release(); d.release(); e.release(); f.release(); g.release(); p.release(); t.release(); }
        } scratch_guard{d_vert, d_w, d_elems, d_keys, d_keys2, d_vals, d_order, d_part, d_tmp};
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        double measure = 0.0;
        std::vector<u32> perm(ne);
        h->tri.ensure((size_t)ne * per * D);
        if (weights) h->triw.ensure((size_t)ne * 3);
        if (ne > 0) {
            d_vert.ensure((size_t)nv * stride); d_elems.ensure((size_t)ne * per);
            CUDA_CHECK(cudaMemcpyAsync(d_vert.p, vertices, sizeof(double) * (size_t)nv * stride, cudaMemcpyHostToDevice, h->stream));
            CUDA_CHECK(cudaMemcpyAsync(d_elems.p, elems, sizeof(u32) * (size_t)ne * per, cudaMemcpyHostToDevice, h->stream));
            if (weights) {
                d_w.ensure(nv);
                CUDA_CHECK(cudaMemcpyAsync(d_w.p, weights, sizeof(double) * nv, cudaMemcpyHostToDevice, h->stream));
            }
            const u32 nblk = std::min<u32>(div_up(ne, MESHPREP_THREADS), MESHPREP_BLOCKS);
            d_part.ensure(nblk);
            if (per == 3) LAUNCH(h, mesh_bounds_kernel<3>, nblk, MESHPREP_THREADS, 0, d_vert.p, nv, stride, d_elems.p, ne, d_part.p);
            else LAUNCH(h, mesh_bounds_kernel<4>, nblk, MESHPREP_THREADS, 0, d_vert.p, nv, stride, d_elems.p, ne, d_part.p);
            std::vector<MeshPartial> part(nblk);
            CUDA_CHECK(cudaMemcpyAsync(part.data(), d_part.p, sizeof(MeshPartial) * nblk, cudaMemcpyDeviceToHost, h->stream));
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
            bool bad = false;
            for (u32 i = 0; i < nblk; ++i) {
                bad = bad || part[i].bad;
                for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], part[i].lo[a]); hi[a] = std::max(hi[a], part[i].hi[a]); }
                measure += part[i].measure;
            }
            if (bad) throw ArgError("element references a vertex out of range");
            double maxext = 0.0;
            for (int a = 0; a < 3; ++a) maxext = std::max(maxext, hi[a] - lo[a]);
            const double sc = maxext > 0.0 ? 1023.999 / maxext : 0.0;
            d_keys.ensure(ne); d_keys2.ensure(ne); d_vals.ensure(ne); d_order.ensure(ne);
            if (per == 3) LAUNCH(h, mesh_codes_kernel<3>, div_up(ne, 256), 256, 0, d_vert.p, stride, d_elems.p, ne, lo[0], lo[1], lo[2], sc, d_keys.p, d_vals.p);
            else LAUNCH(h, mesh_codes_kernel<4>, div_up(ne, 256), 256, 0, d_vert.p, stride, d_elems.p, ne, lo[0], lo[1], lo[2], sc, d_keys.p, d_vals.p);
            size_t tmp_bytes = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys.p, d_keys2.p, d_vals.p, d_order.p, (int)ne, 0, 30, h->stream);
            d_tmp.ensure(tmp_bytes);
            CUDA_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_keys.p, d_keys2.p, d_vals.p, d_order.p, (int)ne, 0, 30, h->stream));
            h->launches += 4;
            double* sw = weights ? h->triw.p : nullptr;
            if (per == 4) LAUNCH(h, (mesh_soup_kernel<3, 4>), div_up(ne, 256), 256, 0, d_vert.p, stride, d_elems.p, d_order.p, ne, (const double*)nullptr, h->tri.p, (double*)nullptr);
            else if (D == 3) LAUNCH(h, (mesh_soup_kernel<3, 3>), div_up(ne, 256), 256, 0, d_vert.p, stride, d_elems.p, d_order.p, ne, d_w.p, h->tri.p, sw);
            else LAUNCH(h, (mesh_soup_kernel<6, 3>), div_up(ne, 256), 256, 0, d_vert.p, stride, d_elems.p, d_order.p, ne, d_w.p, h->tri.p, sw);
            CUDA_CHECK(cudaMemcpyAsync(perm.data(), d_order.p, sizeof(u32) * ne, cudaMemcpyDeviceToHost, h->stream));
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
        }
        // volumetric: which faces of a tet are shared with another tet (cells.tet_adjacent != NO_CELL); face lf is opposite
        // to corner lf. Taken from the caller's adjacency when given, else from matching sorted face triples.
        std::vector<uint8_t> inner;
        if (h->volumetric) {
            inner.assign(ne, 0);
            std::vector<uint8_t> by_elem(ne, 0);
            if (adjacency) {
                for (u32 t = 0; t < ne; ++t)
                    for (int lf = 0; lf < 4; ++lf)
                        if (adjacency[(size_t)t * 4 + lf] >= 0) by_elem[t] |= (uint8_t)(1u << lf);
            } else {
                struct Face { u32 a, b, c, t, lf; };
                std::vector<Face> F((size_t)ne * 4);
                for (u32 t = 0; t < ne; ++t)
                    for (u32 lf = 0; lf < 4; ++lf) {
                        u32 q[3]; int n = 0;
                        for (u32 lv = 0; lv < 4; ++lv) if (lv != lf) q[n++] = elems[(size_t)t * 4 + lv];
                        std::sort(q, q + 3);
                        F[(size_t)t * 4 + lf] = Face{q[0], q[1], q[2], t, lf};
                    }
                std::sort(F.begin(), F.end(), [](const Face& x, const Face& y) {
                    if (x.a != y.a) return x.a < y.a;
                    if (x.b != y.b) return x.b < y.b;
                    if (x.c != y.c) return x.c < y.c;
                    return x.t < y.t;
                });
                for (size_t i = 0; i + 1 < F.size(); ++i)
                    if (F[i].a == F[i + 1].a && F[i].b == F[i + 1].b && F[i].c == F[i + 1].c) {
                        by_elem[F[i].t] |= (uint8_t)(1u << F[i].lf);
                        by_elem[F[i + 1].t] |= (uint8_t)(1u << F[i + 1].lf);
                    }
            }
            for (u32 i = 0; i < ne; ++i) inner[i] = by_elem[perm[i]];
        }
        h->host_elems.clear(); h->host_perm.clear(); h->host_adj.clear(); h->facet_adj_valid = false; h->rdt_valid = false; h->rdt_valid_mn = false;
        h->host_perm = perm;
        h->perm_dev_valid = false;
        h->vol_weights_dropped = h->volumetric && weights_in != nullptr;
        if (!h->volumetric) {
            h->host_elems.assign(elems, elems + (size_t)ne * 3);
            if (adjacency) h->host_adj.assign(adjacency, adjacency + (size_t)ne * 3);
        }
        h->nv = nv; h->T = ne; h->weighted = weights != nullptr; h->mesh_measure = measure;
        for (int a = 0; a < 3; ++a) { h->bb_lo[a] = lo[a]; h->bb_hi[a] = hi[a]; }
        if (h->volumetric) {
            h->tet_inner.ensure(std::max<size_t>(ne, 1));
            if (ne > 0) CUDA_CHECK(cudaMemcpyAsync(h->tet_inner.p, inner.data(), ne, cudaMemcpyHostToDevice, h->stream));
        }
        h->facet_guess.ensure(ne);
        LAUNCH(h, fill_u32_kernel, 1024, 256, 0, h->facet_guess.p, (size_t)ne, B200_NONE);
        h->facet_ball.ensure(ne); h->facet_rad.ensure(ne); h->facet_cell_valid = false;
        if (ne > 0) {
            if (h->volumetric) LAUNCH(h, (facet_ball_kernel<3, 4>), div_up(ne, 256), 256, 0, h->tri.p, ne, h->facet_ball.p, h->facet_rad.p);
            else if (D == 3) LAUNCH(h, (facet_ball_kernel<3, 3>), div_up(ne, 256), 256, 0, h->tri.p, ne, h->facet_ball.p, h->facet_rad.p);
            else LAUNCH(h, (facet_ball_kernel<6, 3>), div_up(ne, 256), 256, 0, h->tri.p, ne, h->facet_ball.p, h->facet_rad.p);
        }
        h->mesh_vmax2 = 0.0;
        if (ne > 0) {
            DevBuf<unsigned long long> d_mx;
            d_mx.ensure(1);
            CUDA_CHECK(cudaMemsetAsync(d_mx.p, 0, sizeof(unsigned long long), h->stream));
            LAUNCH(h, soup_vmax2_kernel, 1024, 256, 0, h->tri.p, (size_t)ne * per, D, d_mx.p);
            unsigned long long bits = 0;
            CUDA_CHECK(cudaMemcpyAsync(&bits, d_mx.p, sizeof(bits), cudaMemcpyDeviceToHost, h->stream));
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
            memcpy(&h->mesh_vmax2, &bits, sizeof(double));
        }
        h->facet_area.ensure(ne);
        if (ne > 0 && !h->volumetric) {
            if (D == 3) LAUNCH(h, facet_area_kernel<3>, div_up(ne, 256), 256, 0, h->tri.p, ne, h->facet_area.p);
            else LAUNCH(h, facet_area_kernel<6>, div_up(ne, 256), 256, 0, h->tri.p, ne, h->facet_area.p);
        }
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        h->has_mesh = true; h->grid_valid = false; h->knn_valid = false; h->has_results = false; h->pair_cap = 0; h->vgrid_valid = false;
    });
}

int b200cvt_rdt(b200cvt_handle h, uint32_t* tri_out, uint64_t cap_triangles, uint64_t* n_out) {
    return guarded([&] {
        if (!h || !n_out) throw ArgError("null argument");
        CUDA_CHECK(cudaSetDevice(h->device));
        if (!h->rdt_valid) { if (h->volumetric) compute_rdt_volume(h); else if (h->dim == 3) compute_rdt_t<3>(h); else compute_rdt_t<6>(h); }
        const size_t per = h->volumetric ? 4 : 3;
        const uint64_t n = h->rdt_host.size() / per;
        *n_out = n;
        if (tri_out && cap_triangles > 0)
            memcpy(tri_out, h->rdt_host.data(), sizeof(u32) * per * (size_t)std::min<uint64_t>(n, cap_triangles));
    });
}

int b200cvt_rdt_multinerve(b200cvt_handle h, int use_rvc_centroids, int prefer_seeds, const uint8_t* locked,
                           uint32_t* tri_out, uint64_t cap_triangles, uint64_t* n_tri,
                           double* vertices_out, uint32_t* vertex_seed_out, uint64_t cap_vertices, uint64_t* n_vertices) {
    return guarded([&] {
        if (!h || !n_tri || !n_vertices) throw ArgError("null argument");
        CUDA_CHECK(cudaSetDevice(h->device));
        const int mode = (use_rvc_centroids ? 1 : 0) | (prefer_seeds ? 2 : 0) | (locked ? 4 : 0);
        // a call without output buffers (or the first call) computes; a call with buffers copies what the last one computed
        if (!tri_out || h->mn_mode != mode || !h->rdt_valid_mn) {
            if (h->dim == 3) compute_rdt_mn_t<3>(h, use_rvc_centroids, prefer_seeds, locked);
            else compute_rdt_mn_t<6>(h, use_rvc_centroids, prefer_seeds, locked);
            h->mn_mode = mode; h->rdt_valid_mn = true;
        }
        const uint64_t nt = h->mn_tri_host.size() / 3, nv = h->mn_vseed_host.size();
        *n_tri = nt; *n_vertices = nv;
        if (tri_out && cap_triangles > 0) memcpy(tri_out, h->mn_tri_host.data(), sizeof(u32) * 3 * (size_t)std::min<uint64_t>(nt, cap_triangles));
        if (vertices_out && cap_vertices > 0)
            memcpy(vertices_out, h->mn_emb_host.data(), sizeof(double) * h->dim * (size_t)std::min<uint64_t>(nv, cap_vertices));
        if (vertex_seed_out && cap_vertices > 0)
            memcpy(vertex_seed_out, h->mn_vseed_host.data(), sizeof(u32) * (size_t)std::min<uint64_t>(nv, cap_vertices));
    });
}

int b200cvt_set_seeds(b200cvt_handle h, const double* x, uint32_t S) {
    return guarded([&] {
        if (!h || !x) throw ArgError("null argument");
        if (S == 0) throw ArgError("no seeds");
        CUDA_CHECK(cudaSetDevice(h->device));
        h->x.ensure((size_t)S * h->dim);
        CUDA_CHECK(cudaMemcpyAsync(h->x.p, x, sizeof(double) * (size_t)S * h->dim, cudaMemcpyHostToDevice, h->stream));
        if (S != h->S && h->facet_guess.p) LAUNCH(h, fill_u32_kernel, 1024, 256, 0, h->facet_guess.p, (size_t)h->T, B200_NONE);
        set_seeds_common(h, S);
        build_grid(h);
        run_knn_main(h, 20, false, true);
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
    });
}

int b200cvt_initial_sampling(b200cvt_handle h, uint32_t nb_points, double* x_out, int* ok_out) {
    return guarded([&] {
        if (!h) throw ArgError("null handle");
        if (!h->has_mesh || h->T == 0) throw StateError("no mesh: call b200cvt_set_mesh first");
        if (nb_points == 0) throw ArgError("no points");
        if (h->vol_weights_dropped) throw ArgError("weighted volumetric sampling stays on the reference implementation");
        CUDA_CHECK(cudaSetDevice(h->device));
        const u32 T = h->T, S = nb_points;
        const int D = h->dim, per = h->volumetric ? 4 : 3;
        if (!h->perm_dev_valid) {
            std::vector<u32> inv(T);
            for (u32 i = 0; i < T; ++i) inv[h->host_perm[i]] = i;
            h->perm_dev.ensure(T); h->inv_perm_dev.ensure(T);
            CUDA_CHECK(cudaMemcpyAsync(h->perm_dev.p, h->host_perm.data(), sizeof(u32) * T, cudaMemcpyHostToDevice, h->stream));
            CUDA_CHECK(cudaMemcpyAsync(h->inv_perm_dev.p, inv.data(), sizeof(u32) * T, cudaMemcpyHostToDevice, h->stream));
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
            h->perm_dev_valid = true;
        }
        // 1. element masses in the caller's element order (device)
        DevBuf<double> d_mass, d_lam; DevBuf<u32> d_elem;
        d_mass.ensure(T);
        const double* w = (h->weighted && !h->volumetric) ? h->triw.p : nullptr;
        if (per == 4) LAUNCH(h, (sampling_mass_kernel<3, 4>), div_up(T, 256), 256, 0, h->tri.p, (const double*)nullptr, T, h->perm_dev.p, d_mass.p);
        else if (D == 3) LAUNCH(h, (sampling_mass_kernel<3, 3>), div_up(T, 256), 256, 0, h->tri.p, w, T, h->perm_dev.p, d_mass.p);
        else LAUNCH(h, (sampling_mass_kernel<6, 3>), div_up(T, 256), 256, 0, h->tri.p, w, T, h->perm_dev.p, d_mass.p);
        std::vector<double> mass(T);
        CUDA_CHECK(cudaMemcpyAsync(mass.data(), d_mass.p, sizeof(double) * T, cudaMemcpyDeviceToHost, h->stream));
        // 2. the recurrences (host): mt19937_64 stream, sorted uniforms, running sum of mass / total in element order
        std::mt19937_64 engine;                                       // Numeric::random_reset()
        auto rnd = [&] { return std::uniform_real_distribution<double>(0, 1)(engine); };   // Numeric::random_float64()
        std::vector<double> su(S);
        for (u32 i = 0; i < S; ++i) su[i] = rnd();
        std::sort(su.begin(), su.end());
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        double Atot = 0.0;
        for (u32 t = 0; t < T; ++t) Atot += mass[t];
        std::vector<u32> elem(S);
        std::vector<double> lam((size_t)S * 4, 0.0);
        u32 first_t = B200_NONE, last_t = 0, cur_t = 0;
        double cur_s = mass[0] / Atot;
        for (u32 i = 0; i < S; ++i) {
            while (su[i] > cur_s && cur_t < T - 1) { cur_t++; cur_s += mass[cur_t] / Atot; }
            if (first_t == B200_NONE) first_t = cur_t;
            last_t = std::max(last_t, cur_t);
            elem[i] = cur_t;
            double* l = lam.data() + (size_t)i * 4;
            if (per == 3) {
                double l1 = rnd(), l2 = rnd();
                if (l1 + l2 > 1.0) { l1 = 1.0 - l1; l2 = 1.0 - l2; }
                const double l3 = 1.0 - l1 - l2;
                // DIM = 3 takes the vec3 overload: the two draws weight the second and third corner (geometry.h:602-619)
                if (D == 3) { l[0] = l3; l[1] = l1; l[2] = l2; } else { l[0] = l1; l[1] = l2; l[2] = l3; }
            } else {
                double ss = rnd(), tt = rnd(), uu = rnd();
                if (ss + tt > 1.0) { ss = 1.0 - ss; tt = 1.0 - tt; }
                if (tt + uu > 1.0) { const double tmp = uu; uu = 1.0 - ss - tt; tt = 1.0 - tmp; }
                else if (ss + tt + uu > 1.0) { const double tmp = uu; uu = ss + tt + uu - 1.0; ss = 1.0 - tt - tmp; }
                l[0] = 1.0 - ss - tt - uu; l[1] = ss; l[2] = tt; l[3] = uu;
            }
        }
        if (ok_out) *ok_out = (T > 1 && last_t == first_t) ? 0 : 1;
        // 3. the sample points (device), written into the seed array of the handle
        d_elem.ensure(S); d_lam.ensure((size_t)S * 4);
        CUDA_CHECK(cudaMemcpyAsync(d_elem.p, elem.data(), sizeof(u32) * S, cudaMemcpyHostToDevice, h->stream));
        CUDA_CHECK(cudaMemcpyAsync(d_lam.p, lam.data(), sizeof(double) * (size_t)S * 4, cudaMemcpyHostToDevice, h->stream));
        h->x.ensure(std::max<size_t>((size_t)S * D, (size_t)h->slice_len() * h->nranks * D));
        if (per == 4) LAUNCH(h, (sampling_points_kernel<3, 4>), div_up(S, 256), 256, 0, h->tri.p, h->inv_perm_dev.p, d_elem.p, d_lam.p, S, h->x.p);
        else if (D == 3) LAUNCH(h, (sampling_points_kernel<3, 3>), div_up(S, 256), 256, 0, h->tri.p, h->inv_perm_dev.p, d_elem.p, d_lam.p, S, h->x.p);
        else LAUNCH(h, (sampling_points_kernel<6, 3>), div_up(S, 256), 256, 0, h->tri.p, h->inv_perm_dev.p, d_elem.p, d_lam.p, S, h->x.p);
        if (S != h->S && h->facet_guess.p) LAUNCH(h, fill_u32_kernel, 1024, 256, 0, h->facet_guess.p, (size_t)h->T, B200_NONE);
        set_seeds_common(h, S);
        if (x_out) CUDA_CHECK(cudaMemcpyAsync(x_out, h->x.p, sizeof(double) * (size_t)S * D, cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
    });
}

int b200cvt_set_seeds_device(b200cvt_handle h, const double* d_x, uint32_t S) {
    return guarded([&] {
        if (!h || !d_x) throw ArgError("null argument");
        if (S == 0) throw ArgError("no seeds");
        CUDA_CHECK(cudaSetDevice(h->device));
        h->x.ensure((size_t)S * h->dim);
        if (d_x != h->x.p)
            CUDA_CHECK(cudaMemcpyAsync(h->x.p, d_x, sizeof(double) * (size_t)S * h->dim, cudaMemcpyDeviceToDevice, h->stream));
        if (S != h->S && h->facet_guess.p) LAUNCH(h, fill_u32_kernel, 1024, 256, 0, h->facet_guess.p, (size_t)h->T, B200_NONE);
        set_seeds_common(h, S);
    });
}

int b200cvt_knn(b200cvt_handle h, uint32_t k, uint32_t* idx_out, uint32_t* count_out, double* sqdist_out, uint8_t* flags_out) {
    return guarded([&] {
        if (!h || !idx_out || !count_out) throw ArgError("null argument");
        if (!h->has_seeds) throw StateError("no seeds");
        if (k == 0 || k > B200CVT_KMAX) throw ArgError("k out of range");
        CUDA_CHECK(cudaSetDevice(h->device));
        const u32 S = h->S;
        if (!h->grid_valid) build_grid(h);
        CUDA_CHECK(cudaMemsetAsync(h->flags.p, 0, S, h->stream));
        run_knn_main(h, k, true, true);
        DevBuf<u32> d_idx, d_cnt; DevBuf<double> d_sq; DevBuf<uint8_t> d_fl;
        d_idx.ensure((size_t)S * k); d_cnt.ensure(S); d_sq.ensure((size_t)S * k); d_fl.ensure(S);
        if (h->dim == 3)
            LAUNCH(h, knn_export_kernel<3>, div_up(S, 128), 128, 0, (const SeedRec<3>*)h->xs.p, h->nbr.p, h->nbr_n.p, h->sqd.p,
                   h->flags.p, S, k, d_idx.p, d_cnt.p, d_sq.p, d_fl.p);
        else
            LAUNCH(h, knn_export_kernel<6>, div_up(S, 128), 128, 0, (const SeedRec<6>*)h->xs.p, h->nbr.p, h->nbr_n.p, h->sqd.p,
                   h->flags.p, S, k, d_idx.p, d_cnt.p, d_sq.p, d_fl.p);
        CUDA_CHECK(cudaMemcpyAsync(idx_out, d_idx.p, sizeof(u32) * (size_t)S * k, cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaMemcpyAsync(count_out, d_cnt.p, sizeof(u32) * S, cudaMemcpyDeviceToHost, h->stream));
        if (sqdist_out) CUDA_CHECK(cudaMemcpyAsync(sqdist_out, d_sq.p, sizeof(double) * (size_t)S * k, cudaMemcpyDeviceToHost, h->stream));
        if (flags_out) CUDA_CHECK(cudaMemcpyAsync(flags_out, d_fl.p, S, cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        d_idx.release(); d_cnt.release(); d_sq.release(); d_fl.release();
        if (k != 20) h->knn_valid = false;   // the evaluation path uses the default list size (delaunay_nn.cpp:49)
    });
}

int b200cvt_nearest(b200cvt_handle h, const double* q, uint32_t nq, uint32_t* out) {
    return guarded([&] {
        if (!h || !q || !out) throw ArgError("null argument");
        if (!h->has_seeds) throw StateError("no seeds");
        CUDA_CHECK(cudaSetDevice(h->device));
        if (!h->grid_valid) build_grid(h);
        DevBuf<double> dq; DevBuf<u32> dout;
        dq.ensure((size_t)nq * h->dim); dout.ensure(nq);
        CUDA_CHECK(cudaMemcpyAsync(dq.p, q, sizeof(double) * (size_t)nq * h->dim, cudaMemcpyHostToDevice, h->stream));
        if (h->dim == 3) LAUNCH(h, nearest_kernel<3>, div_up(nq, 128), 128, 0, (const SeedRec<3>*)h->xs.p, h->cell_range.p, h->g, dq.p, nq, dout.p);
        else LAUNCH(h, nearest_kernel<6>, div_up(nq, 128), 128, 0, (const SeedRec<6>*)h->xs.p, h->cell_range.p, h->g, dq.p, nq, dout.p);
        CUDA_CHECK(cudaMemcpyAsync(out, dout.p, sizeof(u32) * nq, cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        dq.release(); dout.release();
    });
}

static void accumulate_host(b200cvt_ctx* h, double* s_accum, double* v_accum) {
    const u32 S = h->S; const int D = h->dim;
    std::vector<double> hs(S), hv((size_t)S * D);
    CUDA_CHECK(cudaMemcpyAsync(hs.data(), h->s_orig.p, sizeof(double) * S, cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(hv.data(), h->v_orig.p, sizeof(double) * (size_t)S * D, cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (s_accum) for (u32 i = 0; i < S; ++i) s_accum[i] += hs[i];
    if (v_accum) for (size_t i = 0; i < (size_t)S * D; ++i) v_accum[i] += hv[i];
}

int b200cvt_centroids(b200cvt_handle h, int check_SR, double* mg_accum, double* m_accum) {
    return guarded([&] {
        if (!h || !mg_accum || !m_accum) throw ArgError("null argument");
        CUDA_CHECK(cudaSetDevice(h->device));
        if (h->nranks != 1) throw StateError("host-pointer API needs the whole seed range (nranks == 1)");
        evaluate(h, 0, check_SR);
        scatter_results(h, true, true, false);
        accumulate_host(h, m_accum, mg_accum);
    });
}

int b200cvt_funcgrad(b200cvt_handle h, int check_SR, double* f_accum, double* g_accum) {
    return guarded([&] {
        if (!h || !f_accum || !g_accum) throw ArgError("null argument");
        CUDA_CHECK(cudaSetDevice(h->device));
        if (h->nranks != 1) throw StateError("host-pointer API needs the whole seed range (nranks == 1)");
        evaluate(h, 1, check_SR);
        scatter_results(h, true, true, false);
        const u32 S = h->S;
        std::vector<double> fs(S);
        accumulate_host(h, nullptr, g_accum);
        CUDA_CHECK(cudaMemcpy(fs.data(), h->s_orig.p, sizeof(double) * S, cudaMemcpyDeviceToHost));
        double f = 0.0;
        for (u32 i = 0; i < S; ++i) f += fs[i];
        *f_accum += f;
    });
}

int b200cvt_get_flags(b200cvt_handle h, uint8_t* flags_out) {
    return guarded([&] {
        if (!h || !flags_out) throw ArgError("null argument");
        if (!h->has_results) throw StateError("no evaluation yet");
        CUDA_CHECK(cudaSetDevice(h->device));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        CUDA_CHECK(cudaMemcpy(flags_out, h->flags_orig.p, h->S, cudaMemcpyDeviceToHost));
    });
}

int b200cvt_get_seed_energy(b200cvt_handle h, double* f_seed_out) {
    return guarded([&] {
        if (!h || !f_seed_out) throw ArgError("null argument");
        if (!h->has_results || !h->has_energy) throw StateError("no func/grad evaluation yet");
        CUDA_CHECK(cudaSetDevice(h->device));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        CUDA_CHECK(cudaMemcpy(f_seed_out, h->s_orig.p, sizeof(double) * h->S, cudaMemcpyDeviceToHost));
    });
}

// stats: [0] planes tested [1] planes that cut [2] integration triangles [3] non-empty pairs
//        [4] seeds re-clipped with an enlarged neighbourhood [5] candidate pairs [6] pair row capacity [7] grid cells
//        [8] bisectors tested by the clip kernel  [9..10], [13..15] facet walk (facet_pairs.cuh)  [11] kNN queries of the last evaluation
int b200cvt_get_stats(b200cvt_handle h, uint64_t* out) {
    return guarded([&] {
        if (!h || !out) throw ArgError("null argument");
        CUDA_CHECK(cudaSetDevice(h->device));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        unsigned long long st[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (h->stats.p && h->want_stats) CUDA_CHECK(cudaMemcpy(st, h->stats.p, sizeof(st), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 4; ++i) out[i] = st[i];
        for (int i = 8; i < 16; ++i) out[i] = st[i];
        out[4] = h->host_stats[4];
        u64 total = 0;
        if (h->has_results && h->cnt_orig.p) {
            std::vector<u32> c(h->S);
            CUDA_CHECK(cudaMemcpy(c.data(), h->cnt_orig.p, sizeof(u32) * h->S, cudaMemcpyDeviceToHost));
            for (u32 v : c) total += v;
        }
        out[5] = total; out[6] = h->pair_cap; out[7] = h->g.ncells;
        // [11] seeds served by the kNN launch of the last evaluation (sharded runs: owned range + two-cell halo)
        u32 nq = h->S;
        if (h->nranks > 1 && h->need_n.p) CUDA_CHECK(cudaMemcpy(&nq, h->need_n.p, sizeof(u32), cudaMemcpyDeviceToHost));
        out[11] = nq;
        h->want_stats = true;   // counters are collected from the next evaluation on
    });
}

int b200cvt_set_partition(b200cvt_handle h, uint32_t rank, uint32_t nranks) {
    return guarded([&] {
        if (!h) throw ArgError("null handle");
        if (nranks == 0 || rank >= nranks) throw ArgError("bad partition");
        h->rank = rank; h->nranks = nranks;
        h->knn_valid = false; h->has_results = false; h->rdt_valid = false; h->rdt_valid_mn = false;
    });
}

int b200cvt_get_seeds_device(b200cvt_handle h, double* d_x_out) {
    return guarded([&] {
        if (!h || !d_x_out) throw ArgError("null argument");
        if (!h->has_seeds) throw StateError("no seeds");
        CUDA_CHECK(cudaSetDevice(h->device));
        CUDA_CHECK(cudaMemcpyAsync(d_x_out, h->x.p, sizeof(double) * (size_t)h->S * h->dim, cudaMemcpyDeviceToDevice, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
    });
}

int b200cvt_get_seeds(b200cvt_handle h, double* x_out) {
    return guarded([&] {
        if (!h || !x_out) throw ArgError("null argument");
        if (!h->has_seeds) throw StateError("no seeds");
        CUDA_CHECK(cudaSetDevice(h->device));
        CUDA_CHECK(cudaMemcpyAsync(x_out, h->x.p, sizeof(double) * (size_t)h->S * h->dim, cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
    });
}

uint64_t b200cvt_exchange_chunk_doubles(int dim, uint32_t S, uint32_t nranks) {
    if (nranks == 0) return 0;
    u64 sl = ((u64)S + nranks - 1) / nranks;
    return sl * (u64)(dim + 1);
}

int b200cvt_set_exchange(b200cvt_handle h, double* d_slice, double* d_all, uint64_t chunk_doubles,
                         b200cvt_exchange_cb cb, void* user) {
    return guarded([&] {
        if (!h) throw ArgError("null handle");
        h->x_slice = d_slice; h->x_all = d_all; h->x_chunk = chunk_doubles; h->xcb = cb; h->xuser = user;
    });
}

int b200cvt_set_locked(b200cvt_handle h, const uint8_t* locked, uint32_t S) {
    return guarded([&] {
        if (!h) throw ArgError("null handle");
        CUDA_CHECK(cudaSetDevice(h->device));
        upload_locked(h, locked, S);
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
    });
}

int b200cvt_lloyd_device(b200cvt_handle h, uint32_t nb_iter, b200cvt_progress_cb cb, void* user) {
    return guarded([&] {
        if (!h) throw ArgError("null handle");
        if (!h->has_seeds) throw StateError("no seeds");
        CUDA_CHECK(cudaSetDevice(h->device));
        lloyd_loop(h, nb_iter, cb, user);
    });
}

int b200cvt_lloyd(b200cvt_handle h, uint32_t nb_iter, const uint8_t* locked, double* x_inout, uint32_t S,
                  b200cvt_progress_cb cb, void* user) {
    return guarded([&] {
        if (!h || !x_inout) throw ArgError("null argument");
        if (S == 0) throw ArgError("no seeds");
        CUDA_CHECK(cudaSetDevice(h->device));
        h->x.ensure((size_t)S * h->dim);
        CUDA_CHECK(cudaMemcpyAsync(h->x.p, x_inout, sizeof(double) * (size_t)S * h->dim, cudaMemcpyHostToDevice, h->stream));
        if (S != h->S && h->facet_guess.p) LAUNCH(h, fill_u32_kernel, 1024, 256, 0, h->facet_guess.p, (size_t)h->T, B200_NONE);
        set_seeds_common(h, S);
        upload_locked(h, locked, S);
        bool canceled = false;
        try { lloyd_loop(h, nb_iter, cb, user); } catch (const CanceledError&) { canceled = true; }
        CUDA_CHECK(cudaMemcpyAsync(x_inout, h->x.p, sizeof(double) * (size_t)S * h->dim, cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (canceled) throw CanceledError("canceled by the progress callback");
    });
}

int b200cvt_set_stream(b200cvt_handle h, void* stream) {
    return guarded([&] {
        if (!h) throw ArgError("null handle");
        CUDA_CHECK(cudaSetDevice(h->device));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
        h->stream = (cudaStream_t)stream;
    });
}

int b200cvt_get_cumulative(b200cvt_handle h, double* ms_out, uint64_t* evals_out, int reset) {
    return guarded([&] {
        if (!h || !ms_out) throw ArgError("null argument");
        CUDA_CHECK(cudaSetDevice(h->device));
        sync_stream(h);
        for (int i = 0; i < 6; ++i) ms_out[i] = h->cum_ms[i];
        if (evals_out) *evals_out = h->cum_evals;
        if (reset) { for (int i = 0; i < 6; ++i) h->cum_ms[i] = 0.0; h->cum_evals = 0; }
    });
}

int b200cvt_measure_peaks(int device, double* fp32_tflops, double* fp64_tflops, double* copy_gbs) {
    return guarded([&] {
        if (device >= 0) CUDA_CHECK(cudaSetDevice(device));
        measure_peaks(fp32_tflops, fp64_tflops, copy_gbs);
    });
}

int b200cvt_get_timings(b200cvt_handle h, float* ms) {
    return guarded([&] {
        if (!h || !ms) throw ArgError("null argument");
        if (!h->ev_valid) throw StateError("no evaluation yet");
        CUDA_CHECK(cudaSetDevice(h->device));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < 4; ++i) CUDA_CHECK(cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]));
        ms[4] = 0.f; ms[5] = 0.f;
        if (cudaEventQuery(h->ev[5]) == cudaSuccess) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, h->ev[4], h->ev[5]) == cudaSuccess && t >= 0.f) ms[4] = t;
            if (cudaEventElapsedTime(&t, h->ev[0], h->ev[5]) == cudaSuccess && t >= 0.f) ms[5] = t;
        }
        cudaGetLastError();
    });
}

uint64_t b200cvt_launch_count(b200cvt_handle h) { return h ? h->launches : 0; }

}  // extern "C"

// ---------------------------------------------------------------------------------------
// in-library communicator
// ---------------------------------------------------------------------------------------
#define B200_BOX_BYTES (2u << 20)       // the mailbox allocation (a whole 2 MiB block: its IPC handle maps exactly it)

static void comm_alloc_boxes(b200cvt_ctx* h) {
    h->boxes.ensure(B200_BOX_BYTES / sizeof(PeerBox));
    h->pc_seq.ensure(1); h->pc_gtot.ensure(16); h->pc_err.ensure(1);
    CUDA_CHECK(cudaMemsetAsync(h->boxes.p, 0, B200_BOX_BYTES, h->stream));
    CUDA_CHECK(cudaMemsetAsync(h->pc_seq.p, 0, sizeof(unsigned long long), h->stream));
    CUDA_CHECK(cudaMemsetAsync(h->pc_gtot.p, 0, 16 * sizeof(double), h->stream));
    CUDA_CHECK(cudaMemsetAsync(h->pc_err.p, 0, sizeof(unsigned int), h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

static void comm_finish(b200cvt_ctx* h, u32 rank, u32 nranks, PeerBox** peers) {
    memset(&h->pc, 0, sizeof(h->pc));
    h->pc.rank = (int)rank; h->pc.nranks = (int)nranks;
    for (u32 p = 0; p < nranks; ++p) h->pc.boxes[p] = peers[p];
    h->pc.seq = h->pc_seq.p; h->pc.gtot = h->pc_gtot.p; h->pc.error = h->pc_err.p;
    h->rank = rank; h->nranks = nranks; h->has_comm = true;
    { const char* e = getenv("B200CVT_PEER_EXCHANGE"); h->use_peer_exchange = !(e && atoi(e) == 0); }
    h->knn_valid = false; h->has_results = false; h->rdt_valid = false; h->rdt_valid_mn = false; h->prev_valid = false;
}

static void comm_release(b200cvt_ctx* h) {
    if (!h->has_comm) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int p = 0; p < B200_MAX_RANKS; ++p) {
        if (h->ipc_opened[p]) { cudaIpcCloseMemHandle(h->ipc_opened[p]); h->ipc_opened[p] = nullptr; }
        for (int b = 0; b < 2; ++b)
            if (h->pb_ipc_ptr[b][p]) { cudaIpcCloseMemHandle(h->pb_ipc_ptr[b][p]); h->pb_ipc_ptr[b][p] = nullptr; }
    }
    h->pb_valid = false;
    if (h->nccl) { nccl_api().CommDestroy(h->nccl); h->nccl = nullptr; }
    h->has_comm = false;
    memset(&h->pc, 0, sizeof(h->pc)); h->pc.nranks = 1;
}

struct b200cvt_group {
    std::vector<b200cvt_ctx*> members;
    GroupShared shared;
};

// runs fn(rank) on one host thread per member (the device loops synchronise with the host and issue collectives, so the
// ranks must run concurrently); returns the first non-zero status and leaves its message in this thread's last error
template <class F> static int group_run(b200cvt_group* g, F&& fn) {
    const size_t n = g->members.size();
    std::vector<int> rc(n, 0);
    std::vector<std::string> msg(n);
    std::vector<std::thread> th;
    { std::lock_guard<std::mutex> lk(g->shared.mu); g->shared.cancel = 0; }
    for (size_t r = 0; r < n; ++r)
        th.emplace_back([&, r] { rc[r] = fn((u32)r); if (rc[r] != 0) msg[r] = g_last_error; });
    for (auto& t : th) t.join();
    for (size_t r = 0; r < n; ++r)
        if (rc[r] != 0) { g_last_error = "rank " + std::to_string(r) + ": " + msg[r]; return rc[r]; }
    return B200CVT_OK;
}

extern "C" {

int b200cvt_comm_unique_id(uint8_t* id_out) {
    return guarded([&] {
        if (!id_out) throw ArgError("null argument");
        static_assert(sizeof(ncclUniqueId) == B200CVT_COMM_ID_BYTES, "ncclUniqueId size");
        ncclUniqueId id;
        NCCL_CHECK(nccl_api().GetUniqueId(&id));
        memcpy(id_out, &id, sizeof(id));
    });
}

int b200cvt_comm_init(b200cvt_handle h, const uint8_t* id_bytes, uint32_t rank, uint32_t nranks) {
    return guarded([&] {
        if (!h || !id_bytes) throw ArgError("null argument");
        if (nranks == 0 || rank >= nranks || nranks > B200_MAX_RANKS) throw ArgError("bad communicator size");
        CUDA_CHECK(cudaSetDevice(h->device));
        comm_release(h);
        if (nranks == 1) { h->rank = 0; h->nranks = 1; return; }
        ncclUniqueId id;
        memcpy(&id, id_bytes, sizeof(id));
        NCCL_CHECK(nccl_api().CommInitRank(&h->nccl, (int)nranks, id, (int)rank));
        comm_alloc_boxes(h);
        // every rank's mailbox block, opened in this process through CUDA IPC (the handles travel with one all-gather)
        cudaIpcMemHandle_t mine;
        CUDA_CHECK(cudaIpcGetMemHandle(&mine, h->boxes.p));
        DevBuf<unsigned char> d_handles;
        d_handles.ensure(sizeof(mine) * nranks);
        CUDA_CHECK(cudaMemcpyAsync(d_handles.p + sizeof(mine) * rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
        NCCL_CHECK(nccl_api().AllGather(d_handles.p + sizeof(mine) * rank, d_handles.p, sizeof(mine), ncclUint8, h->nccl, h->stream));
        std::vector<cudaIpcMemHandle_t> all(nranks);
        CUDA_CHECK(cudaMemcpyAsync(all.data(), d_handles.p, sizeof(mine) * nranks, cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        PeerBox* peers[B200_MAX_RANKS];
        for (u32 p = 0; p < nranks; ++p) {
            if (p == rank) { peers[p] = h->boxes.p; continue; }
            void* ptr = nullptr;
            CUDA_CHECK(cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess));
            h->ipc_opened[p] = ptr;
            peers[p] = (PeerBox*)ptr;
        }
        comm_finish(h, rank, nranks, peers);
    });
}

int b200cvt_comm_destroy(b200cvt_handle h) {
    return guarded([&] {
        if (!h) throw ArgError("null handle");
        comm_release(h);
        h->rank = 0; h->nranks = 1; h->knn_valid = false; h->has_results = false;
    });
}

int b200cvt_group_create(int n_gpus, int dim, int volumetric, b200cvt_group_handle* out) {
    return guarded([&] {
        if (!out) throw ArgError("out is NULL");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) throw CudaError(std::string("no CUDA device: ") + cudaGetErrorString(e));
        if (n_gpus <= 0) n_gpus = ndev;
        if (n_gpus > ndev || n_gpus > B200_MAX_RANKS) throw ArgError("more GPUs requested than visible");
        std::unique_ptr<b200cvt_group> g(new b200cvt_group);
        g->shared.n = n_gpus;
        struct Cleanup { b200cvt_group* g; bool armed = true; ~Cleanup() { if (armed) for (auto* m : g->members) b200cvt_destroy(m); } } cleanup{g.get()};
        for (int d = 0; d < n_gpus; ++d) {
            b200cvt_handle m = nullptr;
            if (b200cvt_create(d, dim, volumetric, &m) != B200CVT_OK) throw CudaError(g_last_error);
            m->group = &g->shared;
            g->members.push_back(m);
            g->shared.members.push_back(m);
        }
        if (n_gpus > 1) {
            std::vector<ncclComm_t> comms(n_gpus);
            std::vector<int> devs(n_gpus);
            for (int d = 0; d < n_gpus; ++d) devs[d] = d;
            if (!nccl_api().CommInitAll) throw std::runtime_error("ncclCommInitAll is not available");
            NCCL_CHECK(nccl_api().CommInitAll(comms.data(), n_gpus, devs.data()));
            PeerBox* peers[B200_MAX_RANKS];
            for (int d = 0; d < n_gpus; ++d) {
                CUDA_CHECK(cudaSetDevice(d));
                g->members[d]->nccl = comms[d];
                comm_alloc_boxes(g->members[d]);
                peers[d] = g->members[d]->boxes.p;
                for (int o = 0; o < n_gpus; ++o) {
                    if (o == d) continue;
                    cudaError_t pe = cudaDeviceEnablePeerAccess(o, 0);
                    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CUDA_CHECK(pe);
                    cudaGetLastError();
                }
            }
            for (int d = 0; d < n_gpus; ++d) comm_finish(g->members[d], (u32)d, (u32)n_gpus, peers);
        }
        cleanup.armed = false;
        *out = g.release();
    });
}

void b200cvt_group_destroy(b200cvt_group_handle g) {
    if (!g) return;
    for (auto* m : g->members) { comm_release(m); b200cvt_destroy(m); }
    delete g;
}

uint32_t b200cvt_group_size(b200cvt_group_handle g) { return g ? (uint32_t)g->members.size() : 0; }

b200cvt_handle b200cvt_group_member(b200cvt_group_handle g, uint32_t rank) {
    return (g && rank < g->members.size()) ? g->members[rank] : nullptr;
}

int b200cvt_group_set_mesh(b200cvt_group_handle g, const double* vertices, uint32_t nv, uint32_t stride, const uint32_t* elems,
                           const int32_t* adjacency, uint32_t ne, const double* weights) {
    if (!g) { g_last_error = "null group"; return B200CVT_ERR_ARG; }
    return group_run(g, [&](u32 r) { return b200cvt_set_mesh(g->members[r], vertices, nv, stride, elems, adjacency, ne, weights); });
}

int b200cvt_group_lloyd(b200cvt_group_handle g, uint32_t nb_iter, const uint8_t* locked, double* x_inout, uint32_t S,
                        b200cvt_progress_cb cb, void* user) {
    if (!g || !x_inout) { g_last_error = "null argument"; return B200CVT_ERR_ARG; }
    // every rank starts from the caller's seeds and ends with the same ones; rank 0 writes them back and reports progress
    std::vector<std::vector<double>> copy(g->members.size());
    return group_run(g, [&](u32 r) {
        double* x = x_inout;
        if (r != 0) { copy[r].assign(x_inout, x_inout + (size_t)S * g->members[r]->dim); x = copy[r].data(); }
        return b200cvt_lloyd(g->members[r], nb_iter, locked, x, S, r == 0 ? cb : nullptr, user);
    });
}

int b200cvt_group_newton(b200cvt_group_handle g, uint32_t nb_iter, uint32_t m, const uint8_t* locked, double* x_inout, uint32_t S,
                         b200cvt_progress_cb cb, void* user, uint32_t* info_out) {
    if (!g || !x_inout) { g_last_error = "null argument"; return B200CVT_ERR_ARG; }
    std::vector<std::vector<double>> copy(g->members.size());
    return group_run(g, [&](u32 r) {
        double* x = x_inout;
        if (r != 0) { copy[r].assign(x_inout, x_inout + (size_t)S * g->members[r]->dim); x = copy[r].data(); }
        uint32_t info[4] = {0, 0, 0, 0};
        const int rc = b200cvt_newton(g->members[r], nb_iter, m, locked, x, S, r == 0 ? cb : nullptr, user, info);
        if (r == 0 && info_out) memcpy(info_out, info, sizeof(info));
        return rc;
    });
}

}  // extern "C"

#include "newton.inl"
