// rdt.cuh — restricted Delaunay triangulation, simple mode.
//
// Replaces RestrictedVoronoiDiagram::compute_RDT with RDTMode(0) for surfaces
//   (geogram/voronoi/RVD.cpp:2302-2372: for_each_primal_triangle(GetPrimalTriangles)),
//   PrimalTriangleAction (geogram/voronoi/generic_RVD.h:575-619): every vertex of a clipped polygon of seed s
//   whose symbolic representation holds two bisectors (b0, b1) is a restricted Voronoi vertex and yields the
//   triangle (s, b0, b1), emitted once, from the cell of its smallest seed;
//   the symbolic bookkeeping of Polygon::initialize_from_mesh_facet (generic_RVD_polygon.cpp:46-100),
//   Polygon::clip_by_plane_fast (generic_RVD_polygon.h:241-366) and SymbolicVertex::intersect_symbolic
//   (generic_RVD_vertex.h:582-640).
//
// Mapping: the warp-per-seed / lane-per-candidate-facet scheme of clip.cuh (same plane table in shared memory, same
// arithmetic, check_SR = true semantics: a seed whose neighbour list is used up before the radius test passes is pushed
// to the redo list and comes back with a longer list). Every polygon vertex carries its symbolic set (three sorted
// integers: -(facet + 1) for mesh facets, seed + 1 for bisectors, original indices). Triangles are appended to one
// global list; a seed emits only once it is known to be final (second pass when its candidates do not fit one round).
#pragma once
#include "common.cuh"
#include "clip.cuh"

struct RdtArgs {
    const void* xs;
    const u32* nbr; const u32* nbr_n; u32 kstride;
    int nbr_by_slot;
    const double* tri;         // [T][3][D]
    const int* facet_adj;      // [T][3] facet across the edge (corner c, corner c+1), -1: border; sorted facet ids
    u32 T;
    const u32* pair_cnt; const u32* pair_facet; u32 cap;
    const u32* seed_list; u32 nseeds; const u32* nseeds_dev; u32 qbegin;
    u32 S;
    uint8_t* flags;            // [S] sorted order (B200CVT_FLAG_POLY_OVERFLOW / KMAX are OR-ed in)
    u32* redo_list; u32* redo_n;
    u32* out_tri;              // [out_cap][3] original seed indices (s, b0, b1), s < b1 < b0 as the reference emits them
    u32 out_cap;
    u32* out_n;                // number of triangles found (may exceed out_cap: the caller grows the list and reruns)
};

// small_set<int, 3> (generic_RVD_vertex.h:60-260): sorted, no duplicates; n = 4 records an overflow
struct SymSet { int v[3]; int n; };

__device__ __forceinline__ void sym_insert(SymSet& s, int x) {
    int pos = 0;
    while (pos < s.n && pos < 3 && s.v[pos] < x) ++pos;
    if (pos < s.n && pos < 3 && s.v[pos] == x) return;
    if (s.n >= 3) { s.n = 4; return; }
    for (int i = s.n; i > pos; --i) s.v[i] = s.v[i - 1];
    s.v[pos] = x;
    ++s.n;
}

// sets_intersect (generic_RVD_vertex.h:354-375)
__device__ __forceinline__ SymSet sym_common(const SymSet& a, const SymSet& b) {
    SymSet r; r.n = 0; r.v[0] = r.v[1] = r.v[2] = 0;
    int i = 0, j = 0;
    const int na = min(a.n, 3), nb = min(b.n, 3);
    while (i < na && j < nb) {
        if (a.v[i] < b.v[j]) ++i;
        else if (b.v[j] < a.v[i]) ++j;
        else { r.v[r.n++] = a.v[i]; ++i; ++j; }
    }
    return r;
}

template <int D>
__global__ void __launch_bounds__(CLIP_WARPS * 32)
rdt_kernel(RdtArgs a) {
    extern __shared__ double s_dyn[];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const SeedRec<D>* xs = (const SeedRec<D>*)a.xs;
    // per-warp plane table: [kstride][D] normals, [kstride] offsets, [kstride] squared distances, [kstride] seed ids
    double* pl_n = s_dyn + (size_t)w * a.kstride * (D + 3);
    double* pl_d = pl_n + (size_t)a.kstride * D;
    double* pl_dij = pl_d + a.kstride;
    int* pl_id = (int*)(pl_dij + a.kstride);

    const u32 nseeds = a.nseeds_dev ? *a.nseeds_dev : a.nseeds;
    for (u32 si = blockIdx.x * CLIP_WARPS + w; si < nseeds; si += gridDim.x * CLIP_WARPS) {
        const u32 s = a.seed_list ? a.seed_list[si] : a.qbegin + si;
        double pi[D];
#pragma unroll
        for (int c = 0; c < D; ++c) pi[c] = xs[s].p[c];
        const u32 s_orig = (u32)xs[s].orig;
        const size_t nrow = a.nbr_by_slot ? (size_t)si : (size_t)s;
        const u32 nn = min(a.nbr_n[nrow], a.kstride);
        __syncwarp();
        for (u32 j = lane; j < nn; j += 32) {
            const SeedRec<D>* rj = xs + a.nbr[nrow * a.kstride + j];
            double pj[D];
#pragma unroll
            for (int c = 0; c < D; ++c) pj[c] = rj->p[c];
            double d = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double nc = pi[c] - pj[c];
                pl_n[j * D + c] = nc;
                d += (pi[c] + pj[c]) * nc;
            }
            pl_d[j] = d;
            pl_dij[j] = dist2<D>(pi, pj);
            pl_id[j] = (int)rj->orig + 1;
        }
        __syncwarp();

        const u32 npairs = min(a.pair_cnt[s], a.cap);
        if (npairs == 0) continue;
        const u32* row = a.pair_facet + (size_t)s * a.cap;
        const int npass = npairs > 32 ? 2 : 1;
        u32 lflags = 0;
        bool emit_ok = (npass == 1);      // one round: decided after the clip of that round
        for (int pass = 0; pass < npass; ++pass) {
            bool lexh = false;
            for (u32 base = 0; base < npairs; base += 32) {
                const u32 pidx = base + lane;
                const bool active = pidx < npairs;
                double P[2][CLIP_MAXV][D];
                SymSet Sy[2][CLIP_MAXV];
                double L[CLIP_MAXV];
                int n = 0, cur = 0;
                double R2 = 0.0;
                if (active) {
                    const u32 f = row[pidx];
                    const double* t = a.tri + (size_t)f * 3 * D;
                    int adj[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
#pragma unroll
                        for (int c = 0; c < D; ++c) P[0][i][c] = t[i * D + c];
                        R2 = fmax(R2, dist2<D>(pi, P[0][i]));
                        const int af = a.facet_adj[(size_t)f * 3 + i];
                        adj[i] = af >= 0 ? af : (int)(a.T + i);      // "virtual" boundary facet (generic_RVD_polygon.cpp:84-88)
                    }
                    // corner i2: the facet and the facets across its two incident edges (generic_RVD_polygon.cpp:69-100)
#pragma unroll
                    for (int i2 = 0; i2 < 3; ++i2) {
                        const int i1 = (i2 + 2) % 3;
                        SymSet q; q.n = 0; q.v[0] = q.v[1] = q.v[2] = 0;
                        sym_insert(q, -((int)f + 1));
                        sym_insert(q, -(adj[i1] + 1));
                        sym_insert(q, -(adj[i2] + 1));
                        Sy[0][i2] = q;
                    }
                    n = 3;
                }
                bool done = !active;
                bool sr_ok = !active;
                for (u32 jj = 0; jj < nn; ++jj) {
                    if (__all_sync(B200_FULL, done)) break;
                    if (!done) {
                        if (pl_dij[jj] > 4.1 * R2) { done = true; sr_ok = true; }
                        else {
                            const double d = pl_d[jj];
                            double nj[D];
#pragma unroll
                            for (int c = 0; c < D; ++c) nj[c] = pl_n[jj * D + c];
                            bool cut = false;
                            for (int k = 0; k < n; ++k) {
                                double l = 0.0;
#pragma unroll
                                for (int c = 0; c < D; ++c) l += P[cur][k][c] * nj[c];
                                L[k] = l;
                                cut |= !(2.0 * l - d > 0.0);
                            }
                            if (cut) {
                                const int nxt = cur ^ 1;
                                int m = 0;
                                int prev = n - 1;
                                double prev_l = L[prev];
                                double tp = 2.0 * prev_l - d;
                                int prev_status = (tp > 0.0) - (tp < 0.0);
                                for (int k = 0; k < n; ++k) {
                                    const double l = L[k];
                                    const double tk = 2.0 * l - d;
                                    const int status = (tk > 0.0) - (tk < 0.0);
                                    if (status != prev_status && prev_status != 0) {
                                        // intersect_symbolic (generic_RVD_vertex.h:582-640)
                                        SymSet q = sym_common(Sy[cur][prev], Sy[cur][k]);
                                        sym_insert(q, pl_id[jj]);
                                        const bool sym_ok = (q.n == 3);
                                        const double denom = 2.0 * (prev_l - l);
                                        double l1, l2;
                                        if (fabs(denom) < 1e-20) { l1 = 0.5; l2 = 0.5; }
                                        else { l1 = (d - 2.0 * l) / denom; l2 = 1.0 - l1; }
                                        if (m < CLIP_MAXV) {
                                            if (sym_ok) {
#pragma unroll
                                                for (int c = 0; c < D; ++c)
                                                    P[nxt][m][c] = l1 * P[cur][prev][c] + l2 * P[cur][k][c];
                                                Sy[nxt][m] = q;
                                            } else {
                                                // the reference's workaround: the previous vertex is copied into the
                                                // result (generic_RVD_polygon.h:305-314)
#pragma unroll
                                                for (int c = 0; c < D; ++c) P[nxt][m][c] = P[cur][prev][c];
                                                Sy[nxt][m] = Sy[cur][prev];
                                            }
                                        }
                                        ++m;
                                    }
                                    if (status > 0) {
                                        if (m < CLIP_MAXV) {
#pragma unroll
                                            for (int c = 0; c < D; ++c) P[nxt][m][c] = P[cur][k][c];
                                            Sy[nxt][m] = Sy[cur][k];
                                        }
                                        ++m;
                                    }
                                    prev = k; prev_l = l; prev_status = status;
                                }
                                if (m > CLIP_MAXV) { lflags |= 4; m = CLIP_MAXV; }
                                n = m; cur = nxt;
                                R2 = 0.0;
                                for (int k = 0; k < n; ++k) R2 = fmax(R2, dist2<D>(pi, P[cur][k]));
                            }
                        }
                    }
                }
                if (active && !sr_ok && nn > 0 && n > 0) lexh = true;
                if (npass == 1) {
                    // the only round: the seed is final unless some lane used its list up and the list can still grow
                    const bool any_exh = __any_sync(B200_FULL, lexh);
                    emit_ok = !(any_exh && nn + 1 < a.S && nn < B200CVT_KMAX_DEV);
                }
                if (emit_ok && (npass == 1 || pass == 1)) {
                    for (int k = 0; k < n; ++k) {
                        const SymSet q = Sy[cur][k];
                        if (q.n == 3 && q.v[1] > 0 && q.v[0] <= 0) {
                            // two bisectors: bisector(0) is the LAST entry of the sorted set (generic_RVD_vertex.h:481-484)
                            const u32 b0 = (u32)(q.v[2] - 1), b1 = (u32)(q.v[1] - 1);
                            if (s_orig < b0 && s_orig < b1) {
                                const u32 pos = atomicAdd(a.out_n, 1u);
                                if (pos < a.out_cap) {
                                    a.out_tri[(size_t)pos * 3] = s_orig; a.out_tri[(size_t)pos * 3 + 1] = b0; a.out_tri[(size_t)pos * 3 + 2] = b1;
                                }
                            }
                        }
                    }
                }
            }
            if (npass == 2 && pass == 0) {
                const bool any_exh = __any_sync(B200_FULL, lexh);
                emit_ok = !(any_exh && nn + 1 < a.S && nn < B200CVT_KMAX_DEV);
                if (!emit_ok) {
                    if (lane == 0) lflags |= 0x100;
                    break;
                }
                if (lane == 0 && any_exh) lflags |= 0x200;
            } else if (npass == 1) {
                const bool any_exh = __any_sync(B200_FULL, lexh);
                if (lane == 0) lflags |= !emit_ok ? 0x100 : (any_exh ? 0x200 : 0);
            }
        }
        u32 fl = lflags;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) fl |= __shfl_xor_sync(B200_FULL, fl, m);
        if (lane == 0) {
            uint8_t f8 = (uint8_t)(a.flags[s] & ~(uint8_t)(4 | 8));
            f8 |= (uint8_t)(fl & 4u);
            if (fl & 0x100) {
                const u32 pos = atomicAdd(a.redo_n, 1u);
                a.redo_list[pos] = s;
            } else if ((fl & 0x200) && nn + 1 < a.S) {
                f8 |= 8;       // list used up at the implementation cap: the cell may be truncated
            }
            a.flags[s] = f8;
        }
    }
}
