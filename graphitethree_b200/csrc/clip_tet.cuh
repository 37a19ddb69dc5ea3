// clip_tet.cuh — volumetric mode: clip + integrate for (tetrahedron, seed) pairs.
//
// Replaces, for DIM = 3 and fast predicates,
//   GEOGen::RestrictedVoronoiDiagram::compute_volumetric_with_seeds_priority (generic_RVD.h:1464-1597),
//   clip_by_cell_SR(index_t, Polyhedron&) (generic_RVD.h:2282-2347),
//   GEOGen::ConvexCell::{initialize_from_mesh_tetrahedron, clip_by_plane, find_furthest_point_linear_scan,
//   propagate_conflict_list, triangulate_hole} (generic_RVD_cell.cpp:208-282, generic_RVD_cell.h:281-334, 894-1143),
//   Vertex::intersect_geom / side_fast (generic_RVD_vertex.h:976-1030),
//   TetrahedronAction + ComputeCentroidsVolumetric (generic_RVD.h:901-978, RVD.cpp:428-497) and
//   VolumetricIntegrationSimplexAction + ComputeCVTFuncGradVolumetric (generic_RVD.h:726-786, RVD.cpp:791-876).
//
// Mapping: one warp owns one seed (its bisector data staged in shared memory), each lane takes one candidate
// tetrahedron of the seed's row (facet_pairs.cuh, 4 corners) and keeps the convex cell in dual form — dual
// triangles = cell vertices (3 plane ids, 3 adjacent triangles, one point), planes = cell faces — in local
// memory with a free list, exactly the reference's data structure, so that the conflict-zone flood fill and the
// hole triangulation visit the same triangles in the same order and every point is the same interpolation of
// the same two points. Per-seed sums are taken over tets in ascending id with a fixed warp tree.
#pragma once
#include "common.cuh"
#include "clip.cuh"

#define TETC_WARPS 4
#define TETC_MAXT 48      // dual triangles (cell vertices) per cell, free slots included
#define TETC_MAXP 40      // planes (cell faces): 4 tet faces + cutting bisectors
#define TETC_NONE 0xffu

template <bool EMIT>
struct TetCellT {
    double p[TETC_MAXT][3];
    uint8_t v[TETC_MAXT][3];     // plane ids
    uint8_t t[TETC_MAXT][3];     // adjacent triangles
    uint8_t next[TETC_MAXT];
    uint8_t status[TETC_MAXT];   // 0 used, 1 conflict, 2 free
    uint8_t pn[EMIT ? TETC_MAXP : 1];   // RDT emission only: plane -> position of the neighbour in the seed's list
    uint8_t nt;                  // slots in use (max_t)
    uint8_t np;                  // planes
    uint8_t first_free;
    bool overflow;
};

__device__ __forceinline__ int tetc_plus1(int i) { return i == 2 ? 0 : i + 1; }
__device__ __forceinline__ int tetc_minus1(int i) { return i == 0 ? 2 : i - 1; }

template <class TetCell>
__device__ __forceinline__ int tetc_create_triangle(TetCell& C) {
    if (C.first_free == TETC_NONE) {
        if (C.nt >= TETC_MAXT) { C.overflow = true; return 0; }
        const int r = C.nt++;
        C.status[r] = 0;
        return r;
    }
    const int r = C.first_free;
    C.first_free = C.next[r];
    C.status[r] = 0;
    return r;
}

template <class TetCell>
__device__ __forceinline__ int tetc_find_vertex(const TetCell& C, int t, int v) {
    return (int)((C.v[t][1] == v) | ((C.v[t][2] == v) * 2));
}

// ConvexCell::clip_by_plane<3>, fast predicates. Returns false if the bisector did not touch the cell.
template <bool EMIT>
__device__ __noinline__ bool tetc_clip(TetCellT<EMIT>& C, const double* pi, const double* pj, u32 jj, unsigned long long& st_pv) {
    // Phase I: furthest point on pj's side, then flood fill of the conflict zone (side_fast)
    int furthest = -1;
    double fd = 0.0;
    for (int t = 0; t < C.nt; ++t) {
        if (C.status[t] != 0) continue;
        double d = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double q = C.p[t][c];
            d += (q - pj[c]) * (q - pj[c]);
            d -= (q - pi[c]) * (q - pi[c]);
        }
        ++st_pv;
        if (d < fd) { furthest = t; fd = d; }
    }
    if (!(fd < 0.0)) return false;
    if (C.np >= TETC_MAXP) { C.overflow = true; return false; }
    const int new_v = C.np++;
    if (EMIT) C.pn[new_v] = (uint8_t)jj;
    int cbegin = TETC_NONE, cend = TETC_NONE;
    uint8_t stack[TETC_MAXT];
    int sn = 0;
    stack[sn++] = (uint8_t)furthest;
    C.next[furthest] = (uint8_t)cbegin; C.status[furthest] = 1; cbegin = furthest; cend = furthest;
    while (sn > 0) {
        const int t = stack[--sn];
        for (int e = 0; e < 3; ++e) {
            const int nb = C.t[t][e];
            if (C.status[nb] == 1) continue;
            double r = 0.0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double q = C.p[nb][c];
                r += (pj[c] - q) * (pj[c] - q);
                r -= (pi[c] - q) * (pi[c] - q);
            }
            if (r < 0.0) {
                stack[sn++] = (uint8_t)nb;
                C.next[nb] = (uint8_t)cbegin; C.status[nb] = 1; cbegin = nb;
            }
        }
    }
    // Phase II: a conflict triangle with a used neighbour
    int t1 = cbegin, e1 = 0;
    bool found = false;
    do {
        for (e1 = 0; e1 < 3; ++e1)
            if (C.status[C.t[t1][e1]] == 0) { found = true; break; }
        if (found) break;
        t1 = C.next[t1];
    } while (t1 != TETC_NONE);
    if (!found) { C.nt = 0; C.np = 0; C.first_free = TETC_NONE; return true; }   // everything removed
    // Phase III: triangulate the hole
    {
        double n[3], d = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) { n[c] = pi[c] - pj[c]; d -= n[c] * (pj[c] + pi[c]); }
        d = 0.5 * d;
        int t = t1, e = e1, t_adj = C.t[t][e];
        int new_first = -1, new_prev = -1;
        do {
            const int v1 = C.v[t][tetc_plus1(e)], v2 = C.v[t][tetc_minus1(e)];
            const int nt = tetc_create_triangle(C);
            if (C.overflow) return true;
            C.v[nt][0] = (uint8_t)new_v; C.v[nt][1] = (uint8_t)v1; C.v[nt][2] = (uint8_t)v2;
            {
                double l1 = 0.0, l2 = 0.0;
                const int ta = C.t[t][e];
#pragma unroll
                for (int c = 0; c < 3; ++c) { l1 += C.p[ta][c] * n[c]; l2 += C.p[t][c] * n[c]; }
                l1 = fabs(l1 + d); l2 = fabs(l2 + d);
                const double l12 = l1 + l2;
                if (l12 > 1e-30) { l1 /= l12; l2 /= l12; } else { l1 = 0.5; l2 = 0.5; }
#pragma unroll
                for (int c = 0; c < 3; ++c) C.p[nt][c] = l1 * C.p[t][c] + l2 * C.p[ta][c];
            }
            C.t[nt][0] = (uint8_t)t_adj;
            C.t[t_adj][(int)((C.t[t_adj][1] == t) | ((C.t[t_adj][2] == t) * 2))] = (uint8_t)nt;
            e = tetc_plus1(e);
            t_adj = C.t[t][e];
            int guard = 0;
            while (C.status[t_adj] == 1) {
                t = t_adj;
                e = tetc_minus1(tetc_find_vertex(C, t, v2));
                t_adj = C.t[t][e];
                if (++guard > 3 * TETC_MAXT) { C.overflow = true; return true; }
            }
            if (new_prev < 0) new_first = nt;
            else { C.t[new_prev][1] = (uint8_t)nt; C.t[nt][2] = (uint8_t)new_prev; }
            new_prev = nt;
        } while (t != t1 || e != e1);
        C.t[new_prev][1] = (uint8_t)new_first;
        C.t[new_first][2] = (uint8_t)new_prev;
    }
    // Phase IV: conflict zone -> free list
    {
        int cur = cbegin;
        while (cur != cend) { C.status[cur] = 2; cur = C.next[cur]; }
        C.status[cend] = 2;
        C.next[cend] = C.first_free;
        C.first_free = (uint8_t)cbegin;
    }
    return true;
}

// Geom::tetra_volume<3> (geometry.h:483-524)
__device__ __forceinline__ double tetc_volume(const double* p1, const double* p2, const double* p3, const double* p4) {
    double U[3], V[3], W[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { U[c] = p2[c] - p1[c]; V[c] = p3[c] - p1[c]; W[c] = p4[c] - p1[c]; }
    const double cx = V[1] * W[2] - V[2] * W[1];
    const double cy = V[2] * W[0] - V[0] * W[2];
    const double cz = V[0] * W[1] - V[1] * W[0];
    return fabs((U[0] * cx + U[1] * cy + U[2] * cz) / 6.0);
}

struct TetClipArgs {
    const void* xs;
    const u32* nbr; const u32* nbr_n; u32 kstride;
    int nbr_by_slot;
    const double* tet;         // [T][4][3] corners
    const uint8_t* tet_inner;  // [T] bit lf: the face opposite to corner lf is shared with another tet
    const u32* pair_cnt; u32* pair_facet; u32 cap;
    const u32* seed_list; u32 nseeds; const u32* nseeds_dev; u32 qbegin;
    int mode, check_SR;
    u32 S;
    double* out_s; double* out_v; uint8_t* flags;
    u32* redo_list; u32* redo_n;
    unsigned long long* stats;
    // mode 3: restricted Delaunay tets (PrimalTetrahedronAction, generic_RVD.h:1036-1058), rows of four ORIGINAL seed indices
    uint4* tets; unsigned long long* tet_n; unsigned long long tet_cap;
};

#ifndef TETC_MINBLK
#define TETC_MINBLK 1
#endif
template <bool EMIT>
__global__ void __launch_bounds__(TETC_WARPS * 32, TETC_MINBLK)
clip_tet_kernel(TetClipArgs a) {
    extern __shared__ double s_dyn[];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const SeedRec<3>* xs = (const SeedRec<3>*)a.xs;
    // per warp: [kstride][3] neighbour positions, [kstride] squared distances
    double* nb_p = s_dyn + (size_t)w * a.kstride * 4;
    double* nb_d = nb_p + (size_t)a.kstride * 3;

    const u32 nseeds = a.nseeds_dev ? *a.nseeds_dev : a.nseeds;
    for (u32 si = blockIdx.x * TETC_WARPS + w; si < nseeds; si += gridDim.x * TETC_WARPS) {
        const u32 s = a.seed_list ? a.seed_list[si] : a.qbegin + si;
        double pi[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) pi[c] = xs[s].p[c];
        const size_t nrow = a.nbr_by_slot ? (size_t)si : (size_t)s;
        const u32 nn = min(a.nbr_n[nrow], a.kstride);
        __syncwarp();
        for (u32 j = lane; j < nn; j += 32) {
            const SeedRec<3>* rj = xs + a.nbr[nrow * a.kstride + j];
            double pj[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) { pj[c] = rj->p[c]; nb_p[j * 3 + c] = pj[c]; }
            nb_d[j] = dist2<3>(pi, pj);
        }
        __syncwarp();

        const u32 npairs = min(a.pair_cnt[s], a.cap);
        u32* row = a.pair_facet + (size_t)s * a.cap;
        // canonical order: ascending tet id (rows are filled through atomics)
        if (npairs > 1) {
            u32 n2 = 32; while (n2 < npairs) n2 <<= 1;
            n2 = min(n2, a.cap);
            for (u32 t = npairs + lane; t < n2; t += 32) row[t] = B200_NONE;
            __syncwarp();
            for (u32 k = 2; k <= n2; k <<= 1)
                for (u32 j = k >> 1; j > 0; j >>= 1) {
                    for (u32 t = lane; t < n2; t += 32) {
                        const u32 p = t ^ j;
                        if (p > t) {
                            const u32 vt = row[t], vp = row[p];
                            const bool up = ((t & k) == 0);
                            if ((vt > vp) == up) { row[t] = vp; row[p] = vt; }
                        }
                    }
                    __syncwarp();
                }
        }

        double acc_s = 0.0, acc_v[3] = {0.0, 0.0, 0.0};
        u32 lflags = 0;
        bool lexh = false;
        unsigned long long st_planes = 0, st_pv = 0, st_tri = 0, st_ne = 0;

        for (u32 base = 0; base < npairs; base += 32) {
            const u32 pidx = base + lane;
            if (pidx >= npairs) continue;
            const u32 f = row[pidx];
            TetCellT<EMIT> C;
            // ConvexCell::initialize_from_mesh_tetrahedron (generic_RVD_cell.cpp:208-252)
            {
                const double* t = a.tet + (size_t)f * 12;
                const uint8_t tv[4][3] = {{2, 1, 3}, {3, 0, 2}, {0, 3, 1}, {2, 0, 1}};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) C.p[k][c] = t[k * 3 + c];
#pragma unroll
                    for (int i = 0; i < 3; ++i) { C.v[k][i] = tv[k][i]; C.t[k][i] = tv[k][i]; }
                    C.status[k] = 0; C.next[k] = TETC_NONE;
                }
                C.nt = 4; C.np = 4; C.first_free = TETC_NONE; C.overflow = false;
            }
            const uint8_t inner = a.tet_inner[f];
            bool sr_ok = false;
            // clip_by_cell_SR (generic_RVD.h:2295-2328): neighbours in increasing distance
            for (u32 jj = 0; jj < nn; ++jj) {
                double R2 = 0.0;
                for (int k = 0; k < C.nt; ++k)
                    if (C.status[k] == 0) R2 = fmax(R2, dist2<3>(pi, C.p[k]));
                if (nb_d[jj] > 4.1 * R2) { sr_ok = true; break; }
                ++st_planes;
                tetc_clip(C, pi, nb_p + jj * 3, jj, st_pv);
                if (C.overflow) break;
            }
            if (C.overflow) { lflags |= 4; continue; }
            if (!sr_ok && nn > 0) {
                // the list is used up: one more radius test on the FINAL piece. Every seed outside the list is at least as far as
                // its last entry, so a piece that passes is exact although the loop never saw a farther neighbour (the reference
                // would enlarge the list by one and stop at once, generic_RVD.h:2330-2346)
                double R2 = 0.0;
                for (int k = 0; k < C.nt; ++k)
                    if (C.status[k] == 0) R2 = fmax(R2, dist2<3>(pi, C.p[k]));
                if (nb_d[nn - 1] > 4.1 * R2) sr_ok = true;
            }
            int t0 = -1;
            for (int k = 0; k < C.nt; ++k) if (C.status[k] == 0) { t0 = k; break; }
            if (t0 < 0) continue;                       // empty cell
            if (!sr_ok && nn > 0) lexh = true;          // list used up before the radius test passed
            ++st_ne;
            if (EMIT) {
                // a vertex of the piece on three bisectors is a Voronoi vertex inside this tet: its Delaunay tet, once
                // (a cell whose list was used up is run again with a longer one: its rows come from the final pass only)
                if (sr_ok || nn == 0 || nn + 1 >= a.S || nn >= B200CVT_KMAX_DEV) {
                    const u32 me = (u32)xs[s].orig;
                    for (int k = 0; k < C.nt; ++k) {
                        if (C.status[k] != 0) continue;
                        if (C.v[k][0] < 4 || C.v[k][1] < 4 || C.v[k][2] < 4) continue;
                        const u32 n1 = (u32)xs[a.nbr[nrow * a.kstride + C.pn[C.v[k][0]]]].orig;
                        const u32 n2 = (u32)xs[a.nbr[nrow * a.kstride + C.pn[C.v[k][1]]]].orig;
                        const u32 n3 = (u32)xs[a.nbr[nrow * a.kstride + C.pn[C.v[k][2]]]].orig;
                        if (me < n1 && me < n2 && me < n3) {
                            const unsigned long long r = atomicAdd(a.tet_n, 1ull);
                            if (r < a.tet_cap) a.tets[r] = make_uint4(me, n1, n2, n3);
                        }
                    }
                }
                continue;
            }
            // v_to_t (init_v_to_t, generic_RVD_cell.h:640-652): the last used triangle incident to each plane
            uint8_t vt[TETC_MAXP];
            for (int v = 0; v < C.np; ++v) vt[v] = TETC_NONE;
            for (int k = 0; k < C.nt; ++k)
                if (C.status[k] == 0) { vt[C.v[k][0]] = (uint8_t)k; vt[C.v[k][1]] = (uint8_t)k; vt[C.v[k][2]] = (uint8_t)k; }
            for (int cv = 0; cv < C.np; ++cv) {
                const int ct = vt[cv];
                if (ct == TETC_NONE) continue;
                // tet-tet faces are skipped by the func/grad action (visit_inner_tets = false)
                if (a.mode == 1 && cv < 4 && ((inner >> cv) & 1)) continue;
                int c1t = ct, c1v = tetc_find_vertex(C, ct, cv);
                if (a.mode == 0) {
                    // facet_is_incident_to_vertex (generic_RVD.h:993-1004): faces through the origin vertex give flat tets
                    int qt = c1t, qv = c1v;
                    bool inc = false;
                    int guard = 0;
                    do {
                        if (qt == t0) { inc = true; break; }
                        const int t2 = C.t[qt][tetc_plus1(qv)];
                        qv = tetc_find_vertex(C, t2, cv); qt = t2;
                    } while ((qt != c1t || qv != c1v) && ++guard < TETC_MAXT);
                    if (inc) continue;
                }
                // fan of the face from its first corner (move_to_next_around_vertex, generic_RVD_cell.h:631-636)
                int c2t = C.t[c1t][tetc_plus1(c1v)], c2v = tetc_find_vertex(C, c2t, cv);
                int c3t = C.t[c2t][tetc_plus1(c2v)], c3v = tetc_find_vertex(C, c3t, cv);
                int guard = 0;
                do {
                    ++st_tri;
                    const double* v1 = C.p[c1t];
                    const double* v2 = C.p[c2t];
                    const double* v3 = C.p[c3t];
                    if (a.mode == 0) {
                        const double* q0 = C.p[t0];
                        const double cur_m = tetc_volume(q0, v1, v2, v3);
                        const double sc = cur_m / 4.0;
                        acc_s += cur_m;
#pragma unroll
                        for (int c = 0; c < 3; ++c) acc_v[c] += sc * (q0[c] + v1[c] + v2[c] + v3[c]);
                    } else {
                        const double mi = tetc_volume(pi, v1, v2, v3);
                        double fi = 0.0;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const double Uc = v1[c] - pi[c], Vc = v2[c] - pi[c], Wc = v3[c] - pi[c];
                            fi += Uc * Uc + Vc * Vc + Wc * Wc;
                            fi += (Uc * Vc + Vc * Wc + Wc * Uc);
                        }
                        fi *= (mi / 10.0);
                        acc_s += fi;
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            acc_v[c] += 2.0 * mi * (0.75 * pi[c] - 0.25 * v1[c] - 0.25 * v2[c] - 0.25 * v3[c]);
                    }
                    c2t = c3t; c2v = c3v;
                    const int t2 = C.t[c3t][tetc_plus1(c3v)];
                    c3v = tetc_find_vertex(C, t2, cv); c3t = t2;
                } while ((c3t != c1t || c3v != c1v) && ++guard < TETC_MAXT);
            }
        }

        acc_s = warp_sum(acc_s);
#pragma unroll
        for (int c = 0; c < 3; ++c) acc_v[c] = warp_sum(acc_v[c]);
        const bool any_exh = __any_sync(B200_FULL, lexh);
        u32 fl = lflags;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) fl |= __shfl_xor_sync(B200_FULL, fl, m);
        if (a.stats) {
            st_planes = (unsigned long long)warp_sum((double)st_planes);
            st_pv = (unsigned long long)warp_sum((double)st_pv);
            st_tri = (unsigned long long)warp_sum((double)st_tri);
            st_ne = (unsigned long long)warp_sum((double)st_ne);
        }
        if (lane == 0) {
            a.out_s[s] = acc_s;
#pragma unroll
            for (int c = 0; c < 3; ++c) a.out_v[(size_t)s * 3 + c] = acc_v[c];
            uint8_t f8 = (uint8_t)(a.flags[s] & ~(uint8_t)(1 | 4 | 8));
            f8 |= (uint8_t)fl;
            if (any_exh) {
                if (!a.check_SR) f8 |= 1;
                else if (nn + 1 >= a.S) { }
                else if (nn >= B200CVT_KMAX_DEV) f8 |= 8;
                else if (a.redo_list) { u32 pos = atomicAdd(a.redo_n, 1u); a.redo_list[pos] = s; }
            }
            a.flags[s] = f8;
            if (a.stats) {
                atomicAdd(&a.stats[0], st_planes); atomicAdd(&a.stats[1], st_pv);
                atomicAdd(&a.stats[2], st_tri); atomicAdd(&a.stats[3], st_ne);
            }
        }
    }
}
