// knn.cuh — Morton-sorted uniform-grid k-nearest-neighbour search.
//
// Replaces Delaunay_NearestNeighbors + BalancedKdTree (geogram/delaunay/delaunay_nn.cpp:44-149,
// geogram/points/kd_tree.cpp:156-387). Result semantics kept: exact FP64 squared distances
// (sum of squared differences in coordinate order, no FMA), ascending order, query k+1
// including the seed itself, then the self / duplicate rule of get_neighbors_internal
// (delaunay_nn.cpp:105-145). Exact ties are ordered by original index and flagged
// (the reference orders them by kd-tree traversal, kd_tree.h:173-195).
//
// One warp per query seed. The (2r+1)^3 block of grid cells around the query is gathered
// 32 cells at a time (one cell range per lane), the candidate indices are staged in shared
// memory, then scanned 32 at a time. Candidates below the current k-th distance go to a
// shared-memory buffer; every 32 of them are bitonic-sorted in registers and merged into
// the warp's sorted top list (32*NPL entries, NPL per lane). The search radius grows ring by
// ring until the k-th distance is certified by the distance to the block boundary.
#pragma once
#include "common.cuh"

#ifndef KNN_WARPS
#define KNN_WARPS 16          // measured at C2: 4 warps 0.377 ms, 8 warps 0.363 ms, 16 warps 0.345 ms per kNN phase
#endif
#define KNN_CAND_CAP 512

struct KnnKey { u64 d; u32 id; };

__device__ __forceinline__ bool key_less(u64 da, u32 ia, u64 db, u32 ib) {
    return (da < db) | ((da == db) & (ia < ib));     // no short-circuit: one predicate, no divergence
}

// compare-exchange with lane^mask; keep_min selects which of the pair this lane keeps
__device__ __forceinline__ void key_cex(u64& d, u32& id, int mask, bool keep_min) {
    u64 od = __shfl_xor_sync(B200_FULL, d, mask);
    u32 oi = __shfl_xor_sync(B200_FULL, id, mask);
    // keys are distinct (ids are) except among padding entries, where either choice is the same: "other > mine" is
    // "not (other < mine)", so one comparison serves both directions (measured: kNN phase 0.350 -> 0.307 ms at C2)
    const bool take = key_less(od, oi, d, id) == keep_min;
    if (take) { d = od; id = oi; }
}

// full bitonic sort of one element per lane, ascending by lane
__device__ __forceinline__ void warp_bitonic_sort(u64& d, u32& id, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            bool up = ((lane & k) == 0);
            bool lower = ((lane & j) == 0);
            key_cex(d, id, j, lower == up);
        }
    }
}

// sorts a bitonic sequence ascending
__device__ __forceinline__ void warp_bitonic_merge(u64& d, u32& id, int lane) {
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) key_cex(d, id, j, (lane & j) == 0);
}

template <int NPL>
struct WarpTopK {
    u64 d[NPL];
    u32 id[NPL];
    bool fresh;
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int r = 0; r < NPL; ++r) { d[r] = ~0ull; id[r] = B200_NONE; }
        fresh = true;
    }
    // merges 32 new keys (one per lane, any order) into the sorted list
    __device__ __forceinline__ void merge(u64 bd, u32 bi, int lane) {
        warp_bitonic_sort(bd, bi, lane);
        if (fresh) { d[0] = bd; id[0] = bi; fresh = false; return; }      // empty list: the sorted batch is the list
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
            u64 rd = __shfl_sync(B200_FULL, bd, 31 - lane);
            u32 ri = __shfl_sync(B200_FULL, bi, 31 - lane);
            bool a_less = key_less(d[r], id[r], rd, ri);
            u64 lod = a_less ? d[r] : rd;  u32 loi = a_less ? id[r] : ri;
            u64 hid = a_less ? rd : d[r];  u32 hii = a_less ? ri : id[r];
            warp_bitonic_merge(lod, loi, lane);
            d[r] = lod; id[r] = loi;
            if (r + 1 < NPL) { warp_bitonic_merge(hid, hii, lane); bd = hid; bi = hii; }
        }
    }
    // element e of the sorted list (warp-uniform e)
    __device__ __forceinline__ void get(int e, u64& od, u32& oi) const {
        u64 sd = 0; u32 si = 0;
#pragma unroll
        for (int r = 0; r < NPL; ++r) if ((e >> 5) == r) { sd = d[r]; si = id[r]; }
        od = __shfl_sync(B200_FULL, sd, e & 31);
        oi = __shfl_sync(B200_FULL, si, e & 31);
    }
};

struct KnnArgs {
    const void* xs;          // SeedRec<D>[S], Morton-sorted
    const uint2* cell_range; // [ncells] (start, end) in sorted positions
    const u32* rank_of;      // orig -> sorted position
    const u32* query_list;   // optional: sorted positions to process (NULL: [qbegin, qend))
    const u32* nq_dev;       // optional: number of entries of query_list, read from device memory
    const u32* ksize;        // per ORIGINAL seed list size (NULL: k)
    int out_by_slot;         // write row qi (position in query_list) instead of row q
    u32 k;                   // default list size
    u32 kstride;             // row stride of nbr
    u32 S, qbegin, qend;
    u32* nbr;                // [S][kstride] neighbours as SORTED positions, ascending distance
    u32* nbr_n;              // [S]
    double* sqd;             // optional [S][kstride]
    uint8_t* flags;          // [S] sorted order, OR-ed with B200CVT_FLAG_TIE
    // temporal coherence: the neighbour lists of the previous evaluation, rows by ORIGINAL seed index, entries
    // ORIGINAL indices. The largest new distance to the old neighbours bounds the new k-th distance, so one
    // pass over the grid cells that meet that ball is enough. prev_in may be NULL; prev_out (optional) is written.
    const u32* prev_in; u32* prev_out; u32 prev_stride;
    // optional: bisector table row of every stored neighbour, written with the list (clip_flat.cuh: PLANE_STRIDE,
    // same values as plane_table_kernel): n = pq - pj, d = sum (pq + pj) n, |pq - pj|^2. Rows by sorted position.
    double* planes;
    float* planes32;         // with planes: FP32 filter copy (facet_pairs.cuh: PLANE32_STRIDE floats: n, |n|^2)
    GridParams g;
};

template <int D, int NPL>
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_kernel(KnnArgs a) {
    __shared__ u32 s_cand[KNN_WARPS][KNN_CAND_CAP];
    __shared__ u64 s_bufd[KNN_WARPS][64];
    __shared__ u32 s_bufi[KNN_WARPS][64];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const SeedRec<D>* xs = (const SeedRec<D>*)a.xs;
    u32* cand = s_cand[w];
    u64* bufd = s_bufd[w];
    u32* bufi = s_bufi[w];
    const u32 nq = a.nq_dev ? *a.nq_dev : a.qend - a.qbegin;
    for (u32 qi = blockIdx.x * KNN_WARPS + w; qi < nq; qi += gridDim.x * KNN_WARPS) {
        const u32 q = a.query_list ? a.query_list[qi] : a.qbegin + qi;
        double pq[D];
#pragma unroll
        for (int c = 0; c < D; ++c) pq[c] = xs[q].p[c];
        const u32 qorig = (u32)xs[q].orig;
        const size_t orow = a.out_by_slot ? (size_t)qi : (size_t)q;
        u32 kk = a.ksize ? a.ksize[qorig] : a.k;
        if (kk > a.kstride) kk = a.kstride;
        if (kk > a.S - 1) kk = a.S - 1;
        // kq = kk+1 (with the seed itself) and one more to detect a tie at the cut
        u32 kq = kk + 1; if (kq > a.S) kq = a.S;
        u32 kreq = kq + 1; if (kreq > a.S) kreq = a.S;
        if (kreq > 32u * NPL) kreq = 32u * NPL;
        // grid coordinates are evaluated once per warp, one (axis, offset) per lane, and broadcast
        int c0[3];
        {
            const int ax = lane % 3;
            const double pv = ax == 0 ? pq[0] : (ax == 1 ? pq[1] : pq[2]);
            const int gcv = grid_coord(a.g, pv, ax);
#pragma unroll
            for (int x = 0; x < 3; ++x) c0[x] = __shfl_sync(B200_FULL, gcv, x);
        }

        WarpTopK<NPL> top;
        // threshold from the previous lists
        bool have_thr = false;
        u64 thr0_d = ~0ull;
        if (a.prev_in && kk <= 32u && kk <= a.prev_stride) {
            u64 dk = 0;
            bool ok = true;
            if ((u32)lane < kk) {
                const u32 pid = a.prev_in[(size_t)qorig * a.prev_stride + lane];
                ok = pid != B200_NONE;
                if (ok) {
                    const SeedRec<D>* rec = xs + a.rank_of[pid];
                    double pc[D];
#pragma unroll
                    for (int c = 0; c < D; ++c) pc[c] = rec->p[c];
                    dk = (u64)__double_as_longlong(dist2<D>(pq, pc));
                }
            }
            if (__all_sync(B200_FULL, ok)) {
#pragma unroll
                for (int m = 16; m > 0; m >>= 1) { u64 o = __shfl_xor_sync(B200_FULL, dk, m); dk = max(dk, o); }
                thr0_d = dk;       // kk old neighbours + the seed itself lie within this distance: >= the kq-th distance
                have_thr = true;
            }
        }
        for (int r = 1;; ++r) {
            top.reset();
            u64 thr_d = have_thr ? thr0_d : ~0ull; u32 thr_i = B200_NONE;
            int nbuf = 0;
            // block of cells: the bounding box of the threshold ball, or the ring r around the query's cell
            int blo[3], bhi[3];
            if (have_thr) {
                const double rad = sqrt(__longlong_as_double((long long)thr0_d)) * (1.0 + 1e-9) + 1e-300;
                const int ax = lane % 3;
                const double pv = ax == 0 ? pq[0] : (ax == 1 ? pq[1] : pq[2]);
                const int gcv = grid_coord(a.g, (lane < 3) ? pv - rad : pv + rad, ax);     // lanes 0-2: low corner, 3-5: high corner
#pragma unroll
                for (int x = 0; x < 3; ++x) { blo[x] = __shfl_sync(B200_FULL, gcv, x); bhi[x] = __shfl_sync(B200_FULL, gcv, 3 + x); }
            } else {
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) { blo[ax] = c0[ax] - r; bhi[ax] = c0[ax] + r; }
            }
            const int sx = bhi[0] - blo[0] + 1, sy = bhi[1] - blo[1] + 1, sz = bhi[2] - blo[2] + 1;
            const int ncell_blk = sx * sy * sz;
            for (int cb = 0; cb < ncell_blk; cb += 32) {
                // one cell per lane
                int ci = cb + lane;
                u32 start = 0, len = 0;
                if (ci < ncell_blk) {
                    int dz = ci / (sx * sy), rem = ci - dz * sx * sy;
                    int dy = rem / sx, dx = rem - dy * sx;
                    int cx = blo[0] + dx, cy = blo[1] + dy, cz = blo[2] + dz;
                    if (cx >= 0 && cy >= 0 && cz >= 0 && cx < a.g.res[0] && cy < a.g.res[1] && cz < a.g.res[2]) {
                        uint2 rg = a.cell_range[morton_encode(a.g, cx, cy, cz)];
                        start = rg.x; len = rg.y - rg.x;
                    }
                }
                // the candidates of these 32 cells, in chunks that fit the staging buffer
                u32 done = 0;            // per lane: candidates of my cell already staged
                while (__any_sync(B200_FULL, done < len)) {
                    u32 rest = len - done;
                    u32 incl = rest;
#pragma unroll
                    for (int m = 1; m < 32; m <<= 1) {
                        u32 t = __shfl_up_sync(B200_FULL, incl, m);
                        if (lane >= m) incl += t;
                    }
                    u32 excl = incl - rest;
                    u32 take = 0;
                    if (excl < KNN_CAND_CAP) take = min(rest, (u32)KNN_CAND_CAP - excl);
                    for (u32 t = 0; t < take; ++t) cand[excl + t] = start + done + t;
                    done += take;
                    u32 total = __shfl_sync(B200_FULL, incl, 31);
                    if (total > KNN_CAND_CAP) total = KNN_CAND_CAP;
                    __syncwarp();
                    for (u32 t0 = 0; t0 < total; t0 += 32) {
                        u32 t = t0 + lane;
                        bool pass = false;
                        u64 dk = 0; u32 oid = 0;
                        if (t < total) {
                            const SeedRec<D>* rec = xs + cand[t];
                            double pc[D];
#pragma unroll
                            for (int c = 0; c < D; ++c) pc[c] = rec->p[c];
                            oid = (u32)rec->orig;
                            dk = (u64)__double_as_longlong(dist2<D>(pq, pc));
                            pass = key_less(dk, oid, thr_d, thr_i);
                        }
                        u32 mask = __ballot_sync(B200_FULL, pass);
                        if (mask) {
                            if (pass) {
                                int pos = nbuf + __popc(mask & ((1u << lane) - 1u));
                                bufd[pos] = dk; bufi[pos] = oid;
                            }
                            nbuf += __popc(mask);
                            __syncwarp();
                            if (nbuf >= 32) {
                                u64 bd = bufd[lane]; u32 bi = bufi[lane];
                                u64 td = 0; u32 ti = 0;
                                if (32 + lane < nbuf) { td = bufd[32 + lane]; ti = bufi[32 + lane]; }
                                __syncwarp();
                                if (32 + lane < nbuf) { bufd[lane] = td; bufi[lane] = ti; }
                                nbuf -= 32;
                                __syncwarp();
                                top.merge(bd, bi, lane);
                                u64 nd; u32 ni;
                                top.get((int)kreq - 1, nd, ni);
                                if (key_less(nd, ni, thr_d, thr_i)) { thr_d = nd; thr_i = ni; }
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            if (nbuf > 0) {
                u64 bd = ~0ull; u32 bi = B200_NONE;
                if (lane < nbuf) { bd = bufd[lane]; bi = bufi[lane]; }
                __syncwarp();
                top.merge(bd, bi, lane);
                u64 nd; u32 ni;
                top.get((int)kreq - 1, nd, ni);
                if (key_less(nd, ni, thr_d, thr_i)) { thr_d = nd; thr_i = ni; }
            }
            if (have_thr) break;     // every seed within the threshold ball was scanned
            // certified when the kreq-th distance is inside the block, or the block is the grid
            bool whole = true;
            double b = 1e300;
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                if (c0[ax] - r > 0) {
                    whole = false;
                    double lo = a.g.lo[ax] + (double)(c0[ax] - r) * a.g.h;
                    b = fmin(b, pq[ax] - lo);
                }
                if (c0[ax] + r < a.g.res[ax] - 1) {
                    whole = false;
                    double hi = a.g.lo[ax] + (double)(c0[ax] + r + 1) * a.g.h;
                    b = fmin(b, hi - pq[ax]);
                }
            }
            if (whole) break;
            b = b * (1.0 - 1e-9) - 1e-300;
            if (b > 0.0 && thr_d != ~0ull && __longlong_as_double((long long)thr_d) < b * b) break;
        }

        // delaunay_nn.cpp:119-144 on the sorted list (first kq entries)
        bool dup_smaller = false;
        bool tie = false;
        u32 nres = 0;
        u64 prev_last = 0;
#pragma unroll
        for (int rr = 0; rr < NPL; ++rr) {
            u32 e = rr * 32 + lane;
            u64 dd = top.d[rr]; u32 ii = top.id[rr];
            bool in = e < kq && ii != B200_NONE;
            bool zero = in && dd == 0ull && ii != qorig;
            dup_smaller |= __any_sync(B200_FULL, zero && ii < qorig);
            bool keep = in && ii != qorig && dd != 0ull;
            // tie: equal to the previous list entry (up to and including the entry after the cut)
            u64 pd = __shfl_up_sync(B200_FULL, dd, 1);
            if (lane == 0) pd = prev_last;
            bool eq = (e > 0) && (e < kreq) && ii != B200_NONE && dd == pd;
            tie |= __any_sync(B200_FULL, eq);
            prev_last = __shfl_sync(B200_FULL, dd, 31);
            u32 kmask = __ballot_sync(B200_FULL, keep);
            if (keep) {
                u32 pos = nres + __popc(kmask & ((1u << lane) - 1u));
                if (pos < kk) {
                    if (a.prev_out && pos < a.prev_stride) a.prev_out[(size_t)qorig * a.prev_stride + pos] = ii;
                    const u32 jpos = a.rank_of[ii];
                    a.nbr[orow * a.kstride + pos] = jpos;
                    if (a.sqd) a.sqd[orow * a.kstride + pos] = __longlong_as_double((long long)dd);
                    if (a.planes) {
                        // generic_RVD_polygon.h:257-274; the squared distance is the list key itself
                        constexpr int PS = (D == 3) ? 6 : D + 2;
                        double row[PS];
                        const SeedRec<D>* rj = xs + jpos;
                        double d = 0.0;
#pragma unroll
                        for (int c = 0; c < D; ++c) {
                            const double pj = rj->p[c];
                            const double nc = pq[c] - pj;
                            row[c] = nc;
                            d += (pq[c] + pj) * nc;
                        }
                        row[D] = d;
                        row[D + 1] = __longlong_as_double((long long)dd);
#pragma unroll
                        for (int c = D + 2; c < PS; ++c) row[c] = 0.0;
                        double2* o = (double2*)(a.planes + (orow * a.kstride + pos) * PS);
#pragma unroll
                        for (int c = 0; c < PS / 2; ++c) o[c] = make_double2(row[2 * c], row[2 * c + 1]);
                        if (a.planes32) {
                            constexpr int PS32 = (D == 3) ? 4 : 8;
                            float r32[PS32];
#pragma unroll
                            for (int c = 0; c < D; ++c) r32[c] = (float)row[c];
                            r32[D] = (float)row[D + 1];
#pragma unroll
                            for (int c = D + 1; c < PS32; ++c) r32[c] = 0.0f;
                            float4* o32 = (float4*)(a.planes32 + (orow * a.kstride + pos) * PS32);
#pragma unroll
                            for (int c = 0; c < PS32 / 4; ++c) o32[c] = make_float4(r32[4 * c], r32[4 * c + 1], r32[4 * c + 2], r32[4 * c + 3]);
                        }
                    }
                }
            }
            nres += __popc(kmask);
        }
        if (nres > kk) nres = kk;
        if (dup_smaller) nres = 0;
        if (a.prev_out) for (u32 e = nres + lane; e < a.prev_stride; e += 32) a.prev_out[(size_t)qorig * a.prev_stride + e] = B200_NONE;
        for (u32 e = nres + lane; e < a.kstride; e += 32) {
            a.nbr[orow * a.kstride + e] = B200_NONE;
            if (a.sqd) a.sqd[orow * a.kstride + e] = -1.0;
        }
        if (lane == 0) {
            a.nbr_n[orow] = nres;
            if (tie) a.flags[q] |= 2;
        }
    }
}

// nearest seed of arbitrary points (Delaunay_NearestNeighbors::nearest_vertex,
// delaunay_nn.cpp:147-149): one thread per query, ring expansion on the grid.
template <int D>
__device__ __forceinline__ u32 grid_nearest(const SeedRec<D>* xs, const uint2* cell_range, const GridParams& g,
                                            const double* p, double* best_d_out) {
    int c0[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) c0[ax] = grid_coord(g, p[ax], ax);
    double best = 1e300; u32 besti = B200_NONE; u32 best_orig = B200_NONE;
    int maxr = 0;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) maxr = max(maxr, max(c0[ax], g.res[ax] - 1 - c0[ax]));
    for (int r = 0; r <= maxr; ++r) {
        for (int dz = -r; dz <= r; ++dz) {
            int cz = c0[2] + dz; if (cz < 0 || cz >= g.res[2]) continue;
            for (int dy = -r; dy <= r; ++dy) {
                int cy = c0[1] + dy; if (cy < 0 || cy >= g.res[1]) continue;
                bool shell_yz = (dz == -r || dz == r || dy == -r || dy == r);
                int step = shell_yz ? 1 : (r > 0 ? 2 * r : 1);
                for (int dx = -r; dx <= r; dx += step) {
                    int cx = c0[0] + dx; if (cx < 0 || cx >= g.res[0]) continue;
                    uint2 rg = cell_range[morton_encode(g, cx, cy, cz)];
                    for (u32 s = rg.x; s < rg.y; ++s) {
                        double d = dist2<D>(p, xs[s].p);
                        u32 o = (u32)xs[s].orig;
                        if (d < best || (d == best && o < best_orig)) { best = d; besti = s; best_orig = o; }
                    }
                }
            }
        }
        if (besti != B200_NONE) {
            double b = 1e300; bool whole = true;
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                if (c0[ax] - r > 0) { whole = false; b = fmin(b, p[ax] - (g.lo[ax] + (double)(c0[ax] - r) * g.h)); }
                if (c0[ax] + r < g.res[ax] - 1) { whole = false; b = fmin(b, (g.lo[ax] + (double)(c0[ax] + r + 1) * g.h) - p[ax]); }
            }
            if (whole) break;
            b = b * (1.0 - 1e-9) - 1e-300;
            if (b > 0.0 && best < b * b) break;
        }
    }
    if (best_d_out) *best_d_out = best;
    return besti;
}

template <int D>
__global__ void nearest_kernel(const SeedRec<D>* xs, const uint2* cell_range, GridParams g,
                               const double* q, u32 nq, u32* out_orig) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    double p[D];
#pragma unroll
    for (int c = 0; c < D; ++c) p[c] = q[(size_t)i * D + c];
    u32 s = grid_nearest<D>(xs, cell_range, g, p, nullptr);
    out_orig[i] = (u32)xs[s].orig;
}
