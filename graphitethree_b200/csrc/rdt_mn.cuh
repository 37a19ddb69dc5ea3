// rdt_mn.cuh — restricted Delaunay triangulation in MULTINERVE mode (the mode every default
// remesh_smooth call takes: remesh:multi_nerve = true, remesh:RVC_centroids = true).
//
// Replaces RestrictedVoronoiDiagram::compute_RDT with RDT_MULTINERVE (| RDT_RVC_CENTROIDS | RDT_PREFER_SEEDS) for surfaces:
//   GetConnectedComponentsPrimalTriangles (geogram/voronoi/RVD.cpp:1901-2264) driven by
//   compute_surfacic_with_cnx_priority (geogram/voronoi/generic_RVD.h:1856-2001):
//   * one primal vertex per CONNECTED COMPONENT of a restricted Voronoi cell (a cell of a thin plate or of two nearby
//     sheets has several); two (facet, seed) pairs of the same seed are connected when the clipped polygon of one has an
//     edge on the facet border they share (Vertex::adjacent_facet, generic_RVD_polygon.h:348-354);
//   * its position: the seed when the component touches the surface border, the seed is locked, centroids are off, or
//     (RDT_PREFER_SEEDS) the seed has a single component; else the centroid of the component (RVD.cpp:2195-2237, 2123-2146);
//   * one triangle per restricted Voronoi vertex (a polygon vertex on two bisectors, on facet f): the components of
//     (f, s1), (f, s2), (f, s3), provided the three pairs exist and touch a bisector (FacetSeedMarking, RVD.cpp:2017-2039).
//
// The reference numbers components in the order its sequential flood fill discovers them and emits each triangle from
// whichever of the three cells it visits last, so numbering, row order and orientation are traversal-defined. Here:
// components are numbered by (original seed index, smallest facet of the component), every triangle is written as
// (component of the smallest seed, of the largest, of the middle one) — the orientation the simple mode emits — and rows
// are sorted and deduplicated. tests compare both sides in an order-independent form (oracle/port.py: canonical_multinerve).
//
// Mapping: rdt.cuh's warp-per-seed / lane-per-candidate-facet clip with symbolic vertices, plus the reference's edge
// bookkeeping (adjacent facet / adjacent seed per polygon vertex); the pairs of a seed are linked in shared memory and
// labelled by min-label propagation inside the warp.
#pragma once
#include "common.cuh"
#include "clip.cuh"
#include "rdt.cuh"

#define MN_MAXC 16            // connected components per cell (more: B200CVT_FLAG_POLY_OVERFLOW)
#define MN_DEAD 0xffu
#define MN_TOUCH 0x80u        // the polygon has an edge on a bisector (FacetSeedMarking only records those pairs)

struct RdtMnArgs {
    RdtArgs r;                 // clip inputs, redo lists (out_tri / out_cap / out_n unused)
    uint8_t* pair_comp;        // [S][cap] per candidate pair: component of its cell | MN_TOUCH, or MN_DEAD
    u32* ncomp;                // [S] by ORIGINAL seed index: number of components
    double* comp_m;            // [S][MN_MAXC] sorted order: area of the component
    double* comp_mg;           // [S][MN_MAXC][D]: area-weighted sum of triangle centroids
    uint8_t* comp_border;      // [S][MN_MAXC]
    uint4* vert;               // restricted Voronoi vertices (facet, emitting seed (sorted pos), bisector(0), bisector(1) (orig))
    u32 vert_cap; u32* vert_n;
};

// per-warp shared memory of the labelling: [cap] links (3 facets), status, label, area, area * centroid
template <int D>
__host__ __device__ inline size_t mn_warp_doubles(u32 kstride, u32 cap) {
    // plane table (D + 3 doubles per row) + per pair: (1 + D) doubles + 3 + 1 u32 (links, label | status)
    return (size_t)kstride * (D + 3) + (size_t)cap * (1 + D) + ((size_t)cap * 4 + 1) / 2;
}

template <int D>
__global__ void __launch_bounds__(CLIP_WARPS * 32)
rdt_mn_kernel(RdtMnArgs A) {
    const RdtArgs& a = A.r;
    extern __shared__ double s_dyn[];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const SeedRec<D>* xs = (const SeedRec<D>*)a.xs;
    double* base = s_dyn + (size_t)w * mn_warp_doubles<D>(a.kstride, a.cap);
    double* pl_n = base;
    double* pl_d = pl_n + (size_t)a.kstride * D;
    double* pl_dij = pl_d + a.kstride;
    int* pl_id = (int*)(pl_dij + a.kstride);
    double* pm = pl_dij + 2 * (size_t)a.kstride;          // [cap] area
    double* pmg = pm + a.cap;                             // [cap][D]
    u32* plink = (u32*)(pmg + (size_t)a.cap * D);         // [cap][3] facets across the border edges of the polygon
    u32* plab = plink + (size_t)a.cap * 3;                // [cap] label (index of a pair) | status << 16

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
        uint8_t* pcomp = A.pair_comp + (size_t)s * a.cap;
        if (npairs == 0) {
            if (lane == 0) A.ncomp[s_orig] = 0;
            continue;
        }
        const u32* row = a.pair_facet + (size_t)s * a.cap;
        const int npass = npairs > 32 ? 2 : 1;
        u32 lflags = 0;
        bool emit_ok = (npass == 1);
        for (int pass = 0; pass < npass; ++pass) {
            bool lexh = false;
            for (u32 base_i = 0; base_i < npairs; base_i += 32) {
                const u32 pidx = base_i + lane;
                const bool active = pidx < npairs;
                double P[2][CLIP_MAXV][D];
                SymSet Sy[2][CLIP_MAXV];
                int AF[2][CLIP_MAXV], AS[2][CLIP_MAXV];       // adjacent facet / adjacent seed of every vertex
                double L[CLIP_MAXV];
                int n = 0, cur = 0;
                double R2 = 0.0;
                u32 f = 0;
                if (active) {
                    f = row[pidx];
                    const double* t = a.tri + (size_t)f * 3 * D;
                    int adj[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
#pragma unroll
                        for (int c = 0; c < D; ++c) P[0][i][c] = t[i * D + c];
                        R2 = fmax(R2, dist2<D>(pi, P[0][i]));
                        const int af = a.facet_adj[(size_t)f * 3 + i];
                        AF[0][i] = af; AS[0][i] = -1;
                        adj[i] = af >= 0 ? af : (int)(a.T + i);
                    }
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
#pragma unroll
                                                for (int c = 0; c < D; ++c) P[nxt][m][c] = P[cur][prev][c];
                                                Sy[nxt][m] = Sy[cur][prev];
                                            }
                                            // edge bookkeeping (generic_RVD_polygon.h:348-354): entering the kept side, the new
                                            // vertex starts an edge of the facet border and ends the edge on the bisector
                                            if (status > 0) { AF[nxt][m] = AF[cur][prev]; AS[nxt][m] = pl_id[jj] - 1; }
                                            else { AF[nxt][m] = -1; AS[nxt][m] = AS[cur][k]; }
                                        }
                                        ++m;
                                    }
                                    if (status > 0) {
                                        if (m < CLIP_MAXV) {
#pragma unroll
                                            for (int c = 0; c < D; ++c) P[nxt][m][c] = P[cur][k][c];
                                            Sy[nxt][m] = Sy[cur][k];
                                            AF[nxt][m] = AF[cur][k]; AS[nxt][m] = AS[cur][k];
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
                    const bool any_exh = __any_sync(B200_FULL, lexh);
                    emit_ok = !(any_exh && nn + 1 < a.S && nn < B200CVT_KMAX_DEV);
                }
                if (emit_ok && (npass == 1 || pass == 1) && active) {
                    // what the component labelling and the vertex positions need from this pair
                    u32 lk[3] = {B200_NONE, B200_NONE, B200_NONE};
                    u32 status = 0;
                    if (n > 0) status |= 1u;
                    int nl = 0;
                    for (int k = 0; k < n; ++k) {
                        const int af = AF[cur][k];
                        if (af >= 0 && (u32)af != f) {
                            bool seen = false;
                            for (int e = 0; e < nl; ++e) seen |= lk[e] == (u32)af;
                            if (!seen && nl < 3) lk[nl++] = (u32)af;
                        }
                        if (AS[cur][k] != -1) status |= 4u;                                        // touches the cell border
                        if (af == -1 && AS[cur][(k + 1) % n] == -1) status |= 2u;                  // edge on the surface border
                    }
                    double mm = 0.0, mg[D];
#pragma unroll
                    for (int c = 0; c < D; ++c) mg[c] = 0.0;
                    for (int i = 1; i + 1 < n; ++i) {
                        // Geom::triangle_area (geometry_nd.h:143-156) and the accumulation of RVD.cpp:1988-2006
                        const double ea = sqrt(dist2<D>(P[cur][0], P[cur][i]));
                        const double eb = sqrt(dist2<D>(P[cur][i], P[cur][i + 1]));
                        const double ec = sqrt(dist2<D>(P[cur][i + 1], P[cur][0]));
                        const double sh = 0.5 * (ea + eb + ec);
                        const double A2 = sh * (sh - ea) * (sh - eb) * (sh - ec);
                        const double cur_m = sqrt(fmax(A2, 0.0));
#pragma unroll
                        for (int c = 0; c < D; ++c) mg[c] += cur_m / 3.0 * (P[cur][0][c] + P[cur][i][c] + P[cur][i + 1][c]);
                        mm += cur_m;
                    }
                    pm[pidx] = mm;
#pragma unroll
                    for (int c = 0; c < D; ++c) pmg[(size_t)pidx * D + c] = mg[c];
#pragma unroll
                    for (int e = 0; e < 3; ++e) plink[(size_t)pidx * 3 + e] = lk[e];
                    plab[pidx] = pidx | (status << 16);
                    // restricted Voronoi vertices of this polygon
                    for (int k = 0; k < n; ++k) {
                        const SymSet q = Sy[cur][k];
                        if (q.n == 3 && q.v[1] > 0 && q.v[0] <= 0) {
                            const u32 pos = atomicAdd(A.vert_n, 1u);
                            if (pos < A.vert_cap)
                                A.vert[pos] = make_uint4((u32)(-q.v[0] - 1), s, (u32)(q.v[2] - 1), (u32)(q.v[1] - 1));
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
        __syncwarp();
        if (!(fl & 0x100)) {
            // ---- connected components of the cell: min-label propagation over the links ----
            // links become pair indices (a link to a facet that is not a live candidate of this seed is dropped)
            for (u32 i = lane; i < npairs; i += 32) {
                for (int e = 0; e < 3; ++e) {
                    const u32 lf = plink[(size_t)i * 3 + e];
                    u32 j = B200_NONE;
                    if (lf != B200_NONE)
                        for (u32 t = 0; t < npairs; ++t)
                            if (row[t] == lf) { j = ((plab[t] >> 16) & 1u) ? t : B200_NONE; break; }
                    plink[(size_t)i * 3 + e] = j;
                }
            }
            __syncwarp();
            for (int sweep = 0; sweep < 4096; ++sweep) {
                bool changed = false;
                for (u32 i = lane; i < npairs; i += 32) {
                    const u32 v = plab[i];
                    if (!((v >> 16) & 1u)) continue;
                    u32 lab = v & 0xffffu;
                    for (int e = 0; e < 3; ++e) {
                        const u32 j = plink[(size_t)i * 3 + e];
                        if (j != B200_NONE) {
                            const u32 lj = plab[j] & 0xffffu;
                            if (lj < lab) { lab = lj; changed = true; }
                            else if (lab < lj) {
                                // the link is used in both directions (the neighbour's polygon may miss the shared edge)
                                atomicMin(&plab[j], (plab[j] & 0xffff0000u) | lab);
                                changed = true;
                            }
                        }
                    }
                    if (lab != (v & 0xffffu)) atomicMin(&plab[i], (v & 0xffff0000u) | lab);
                }
                __syncwarp();
                if (!__any_sync(B200_FULL, changed)) break;
            }
            // ---- components in the order of their smallest facet; per-component sums ----
            // a root is a live pair whose label is its own index
            u32 nc = 0;
            if (lane == 0) {
                // roots sorted by facet id (selection over at most MN_MAXC roots)
                u32 roots[MN_MAXC];
                for (u32 i = 0; i < npairs; ++i) {
                    const u32 v = plab[i];
                    if (((v >> 16) & 1u) && (v & 0xffffu) == i) {
                        if (nc < MN_MAXC) roots[nc] = i;
                        ++nc;
                    }
                }
                if (nc > MN_MAXC) { fl |= 4u; nc = MN_MAXC; }
                // smallest facet of each component
                u32 minf[MN_MAXC];
                for (u32 c = 0; c < nc; ++c) minf[c] = B200_NONE;
                for (u32 i = 0; i < npairs; ++i) {
                    const u32 v = plab[i];
                    if (!((v >> 16) & 1u)) continue;
                    for (u32 c = 0; c < nc; ++c) if (roots[c] == (v & 0xffffu)) minf[c] = min(minf[c], row[i]);
                }
                // order
                u32 ord[MN_MAXC];
                for (u32 c = 0; c < nc; ++c) ord[c] = c;
                for (u32 x = 1; x < nc; ++x) {
                    const u32 o = ord[x];
                    int y = (int)x - 1;
                    while (y >= 0 && minf[ord[y]] > minf[o]) { ord[y + 1] = ord[y]; --y; }
                    ord[y + 1] = o;
                }
                for (u32 c = 0; c < nc; ++c) {
                    double mm = 0.0, mg[D];
                    bool border = false;
#pragma unroll
                    for (int q = 0; q < D; ++q) mg[q] = 0.0;
                    const u32 r = roots[ord[c]];
                    for (u32 i = 0; i < npairs; ++i) {
                        const u32 v = plab[i];
                        if (!((v >> 16) & 1u) || (v & 0xffffu) != r) continue;
                        mm += pm[i];
#pragma unroll
                        for (int q = 0; q < D; ++q) mg[q] += pmg[(size_t)i * D + q];
                        border |= ((v >> 16) & 2u) != 0;
                        pcomp[i] = (uint8_t)(c | (((v >> 16) & 4u) ? MN_TOUCH : 0u));
                    }
                    A.comp_m[(size_t)s * MN_MAXC + c] = mm;
#pragma unroll
                    for (int q = 0; q < D; ++q) A.comp_mg[((size_t)s * MN_MAXC + c) * D + q] = mg[q];
                    A.comp_border[(size_t)s * MN_MAXC + c] = border ? 1 : 0;
                }
                for (u32 i = 0; i < npairs; ++i) {
                    const u32 v = plab[i];
                    bool placed = false;
                    if ((v >> 16) & 1u)
                        for (u32 c = 0; c < nc; ++c) placed |= roots[ord[c]] == (v & 0xffffu);
                    if (!placed) pcomp[i] = (uint8_t)MN_DEAD;
                }
                A.ncomp[s_orig] = nc;
            }
        }
        fl = __shfl_sync(B200_FULL, fl, 0) | fl;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) fl |= __shfl_xor_sync(B200_FULL, fl, m);
        if (lane == 0) {
            uint8_t f8 = (uint8_t)(a.flags[s] & ~(uint8_t)(4 | 8));
            f8 |= (uint8_t)(fl & 4u);
            if (fl & 0x100) {
                const u32 pos = atomicAdd(a.redo_n, 1u);
                a.redo_list[pos] = s;
            } else if ((fl & 0x200) && nn + 1 < a.S) {
                f8 |= 8;
            }
            a.flags[s] = f8;
        }
        __syncwarp();
    }
}

// one thread per restricted Voronoi vertex: the three (facet, seed) pairs -> component ids -> one triangle row
struct MnTriArgs {
    const uint4* vert; const u32* vert_n; u32 vert_cap;
    const void* xs; const u32* rank_of;
    const u32* pair_cnt; const u32* pair_facet; const uint8_t* pair_comp; u32 cap;
    const u32* comp_base;      // [S + 1] by original seed index (exclusive scan of ncomp)
    u32* out_tri; u32 out_cap; u32* out_n;
};

template <int D>
__global__ void mn_triangles_kernel(MnTriArgs a) {
    const SeedRec<D>* xs = (const SeedRec<D>*)a.xs;
    const u32 n = min(*a.vert_n, a.vert_cap);
    for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const uint4 v = a.vert[e];
        const u32 f = v.x;
        u32 so[3] = {(u32)xs[v.y].orig, v.z, v.w};        // original seed indices: the emitting cell, bisector(0), bisector(1)
        u32 cc[3];
        bool ok = true;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const u32 pos = q == 0 ? v.y : a.rank_of[so[q]];
            const u32 np = min(a.pair_cnt[pos], a.cap);
            const u32* row = a.pair_facet + (size_t)pos * a.cap;
            u32 comp = MN_DEAD;
            for (u32 t = 0; t < np; ++t) if (row[t] == f) { comp = a.pair_comp[(size_t)pos * a.cap + t]; break; }
            // FacetSeedMarking only knows the pairs whose polygon touches a bisector (generic_RVD.h:1989-1994)
            if (comp == MN_DEAD || !(comp & MN_TOUCH)) { ok = false; break; }
            cc[q] = a.comp_base[so[q]] + (comp & 0x7fu);
        }
        if (!ok) continue;
        // (smallest seed, largest, middle): the row the simple mode emits from the cell of the smallest seed
        int lo = 0, hi = 0;
#pragma unroll
        for (int q = 1; q < 3; ++q) { if (so[q] < so[lo]) lo = q; if (so[q] > so[hi]) hi = q; }
        const int mid = 3 - lo - hi;
        const u32 pos = atomicAdd(a.out_n, 1u);
        if (pos < a.out_cap) {
            a.out_tri[(size_t)pos * 3] = cc[lo]; a.out_tri[(size_t)pos * 3 + 1] = cc[hi]; a.out_tri[(size_t)pos * 3 + 2] = cc[mid];
        }
    }
}

// one thread per seed (original order): the position of each of its components (RVD.cpp:2195-2237, 2123-2146)
template <int D>
__global__ void mn_embedding_kernel(const void* xs_, const u32* rank_of, u32 S, const u32* ncomp, const u32* comp_base,
                                    const double* comp_m, const double* comp_mg, const uint8_t* comp_border, const uint8_t* locked,
                                    int use_centroids, int prefer_seeds, double* emb, u32* vert_seed) {
    const SeedRec<D>* xs = (const SeedRec<D>*)xs_;
    const u32 o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= S) return;
    const u32 s = rank_of[o];
    const u32 nc = ncomp[o], b = comp_base[o];
    for (u32 c = 0; c < nc; ++c) {
        const bool border = comp_border[(size_t)s * MN_MAXC + c] != 0;
        bool seed = !use_centroids || (locked && locked[o]) || border;
        if (prefer_seeds && nc == 1 && !border) seed = true;
        double* out = emb + (size_t)(b + c) * D;
        if (seed) {
#pragma unroll
            for (int q = 0; q < D; ++q) out[q] = xs[s].p[q];
        } else {
            const double mm = comp_m[(size_t)s * MN_MAXC + c];
            const double scal = (mm < 1e-30 ? 0.0 : 1.0 / mm);
#pragma unroll
            for (int q = 0; q < D; ++q) out[q] = comp_mg[((size_t)s * MN_MAXC + c) * D + q] * scal;
        }
        if (vert_seed) vert_seed[b + c] = o;
    }
}
