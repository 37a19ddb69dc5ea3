// clip.cuh — warp-per-seed clip/integrate kernel for the surface RVD: the general path (any neighbour-list length,
// polygons of up to CLIP_MAXV vertices) behind the fast thread-per-pair path of clip_flat.cuh. It re-evaluates the seeds
// whose neighbourhood had to be enlarged (check_SR) or whose polygon overflowed the fast path.
//
// Replaces the facet-driven double flood-fill of
//   GEOGen::RestrictedVoronoiDiagram::compute_surfacic_with_seeds_priority
//   (geogram/voronoi/generic_RVD.h:1318-1424), clip_by_cell_SR (:2134-2199),
//   Polygon::clip_by_plane_fast (geogram/voronoi/generic_RVD_polygon.h:241-366) and the
//   integration actions ComputeCentroids / ComputeCVTFuncGrad (+Weighted)
//   (geogram/voronoi/RVD.cpp:254-369, 543-724)
// by a seed-driven evaluation: one warp owns one seed, each lane clips one candidate facet
// against the seed's bisector planes (staged in shared memory), integrates its polygon and
// the warp reduces mass / centroid / energy / gradient in FP64.
//
// Candidate (facet, seed) rows come from facet_pairs.cuh.
//
// Arithmetic: every expression follows the reference's operation order and this translation
// unit is compiled with -fmad=false (the reference is built with -ffp-contract=off), so that
// polygons are bit-identical to the reference's; only the summation order of per-seed
// accumulators differs (warp tree instead of traversal order).
#pragma once
#include "common.cuh"
#include "knn.cuh"

#define CLIP_WARPS 4
#define CLIP_MAXV 24
#define B200CVT_KMAX_DEV 252u

// ---------------------------------------------------------------------------------------
// clip + integrate: one warp per seed, one lane per candidate facet
// ---------------------------------------------------------------------------------------
struct ClipArgs {
    const void* xs;
    const u32* nbr; const u32* nbr_n; u32 kstride;   // neighbour table used by this launch
    int nbr_by_slot;           // rows indexed by position in seed_list instead of sorted position
    const double* tri;         // [T][3][D]
    const double* triw;        // [T][3] corner weights or NULL
    const u32* pair_cnt; u32* pair_facet; u32 cap;
    const u32* seed_list;      // optional sorted positions (redo pass); NULL: [qbegin,qend)
    u32 nseeds;                // number of seeds processed (qend-qbegin or list length)
    const u32* nseeds_dev;     // optional: list length read from device memory instead
    u32 qbegin;
    int mode;                  // 0: m, mg   1: f_seed, g
    int check_SR;
    u32 S;
    double* out_s;             // [S] sorted order: m or f_seed
    double* out_v;             // [S][D] sorted order: mg or g
    uint8_t* flags;            // [S] sorted order
    u32* redo_list; u32* redo_n; // seeds whose neighbour list was exhausted (check_SR only)
    unsigned long long* stats; // [0] planes tested [1] cuts [2] triangles [3] non-empty pairs
};

template <int D, bool WEIGHTED>
__global__ void __launch_bounds__(CLIP_WARPS * 32)
clip_kernel(ClipArgs a) {
    extern __shared__ double s_dyn[];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const SeedRec<D>* xs = (const SeedRec<D>*)a.xs;
    // per-warp plane table: [kstride][D] normals, [kstride] offsets, [kstride] squared seed distances
    double* pl_n = s_dyn + (size_t)w * a.kstride * (D + 2);
    double* pl_d = pl_n + (size_t)a.kstride * D;
    double* pl_dij = pl_d + a.kstride;

    const u32 nseeds = a.nseeds_dev ? *a.nseeds_dev : a.nseeds;
    for (u32 si = blockIdx.x * CLIP_WARPS + w; si < nseeds; si += gridDim.x * CLIP_WARPS) {
        const u32 s = a.seed_list ? a.seed_list[si] : a.qbegin + si;
        double pi[D];
#pragma unroll
        for (int c = 0; c < D; ++c) pi[c] = xs[s].p[c];
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
        }
        __syncwarp();

        u32 npairs = min(a.pair_cnt[s], a.cap);
        u32* row = a.pair_facet + (size_t)s * a.cap;
        // canonical order: sort the row by facet id (the fill order comes from atomics)
        if (npairs > 1) {
            if (npairs <= 32) {
                u32 v = lane < npairs ? row[lane] : B200_NONE;
#pragma unroll
                for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        u32 o = __shfl_xor_sync(B200_FULL, v, j);
                        bool up = ((lane & k) == 0), lower = ((lane & j) == 0);
                        bool keep_min = (lower == up);
                        v = keep_min ? min(v, o) : max(v, o);
                    }
                if (lane < npairs) row[lane] = v;
            } else {
                // rows are powers of two long (cap is), so the tail [npairs, n2) is free: pad it, then
                // bitonic-sort n2 entries in place (the row stays L1/L2 resident)
                u32 n2 = 64; while (n2 < npairs) n2 <<= 1;
                for (u32 t = npairs + lane; t < n2; t += 32) row[t] = B200_NONE;
                __syncwarp();
                for (u32 k = 2; k <= n2; k <<= 1)
                    for (u32 j = k >> 1; j > 0; j >>= 1) {
                        for (u32 t = lane; t < n2; t += 32) {
                            u32 p = t ^ j;
                            if (p > t) {
                                u32 vt = row[t], vp = row[p];
                                bool up = ((t & k) == 0);
                                if ((vt > vp) == up) { row[t] = vp; row[p] = vt; }
                            }
                        }
                        __syncwarp();
                    }
            }
            __syncwarp();
        }

        double acc_s = 0.0, acc_v[D];
#pragma unroll
        for (int c = 0; c < D; ++c) acc_v[c] = 0.0;
        u32 lflags = 0;
        bool lexh = false;
        unsigned long long st_planes = 0, st_cuts = 0, st_tri = 0, st_ne = 0;

        for (u32 base = 0; base < npairs; base += 32) {
            const u32 pidx = base + lane;
            const bool active = pidx < npairs;
            double P[2][CLIP_MAXV][D];
            double Wt[2][WEIGHTED ? CLIP_MAXV : 1];
            double L[CLIP_MAXV];
            int n = 0, cur = 0;
            double R2 = 0.0;
            if (active) {
                const u32 f = row[pidx];
                const double* t = a.tri + (size_t)f * 3 * D;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
#pragma unroll
                    for (int c = 0; c < D; ++c) P[0][i][c] = t[i * D + c];
                    if (WEIGHTED) Wt[0][i] = a.triw[(size_t)f * 3 + i];
                    R2 = fmax(R2, dist2<D>(pi, P[0][i]));
                }
                n = 3;
            }
            bool done = !active;
            bool sr_ok = !active;
            // clip_by_cell_SR (generic_RVD.h:2155-2177): neighbours in increasing distance
            for (u32 jj = 0; jj < nn; ++jj) {
                if (__all_sync(B200_FULL, done)) break;
                if (!done) {
                    if (pl_dij[jj] > 4.1 * R2) { done = true; sr_ok = true; }
                    else {
                        // clip_by_plane_fast (generic_RVD_polygon.h:257-366)
                        ++st_planes;
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
                            ++st_cuts;
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
                                    const double denom = 2.0 * (prev_l - l);
                                    double l1, l2;
                                    if (fabs(denom) < 1e-20) { l1 = 0.5; l2 = 0.5; }
                                    else { l1 = (d - 2.0 * l) / denom; l2 = 1.0 - l1; }
                                    if (m < CLIP_MAXV) {
#pragma unroll
                                        for (int c = 0; c < D; ++c)
                                            P[nxt][m][c] = l1 * P[cur][prev][c] + l2 * P[cur][k][c];
                                        if (WEIGHTED) Wt[nxt][m] = l1 * Wt[cur][prev] + l2 * Wt[cur][k];
                                    }
                                    ++m;
                                }
                                if (status > 0) {
                                    if (m < CLIP_MAXV) {
#pragma unroll
                                        for (int c = 0; c < D; ++c) P[nxt][m][c] = P[cur][k][c];
                                        if (WEIGHTED) Wt[nxt][m] = Wt[cur][k];
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
            // list used up before the radius test passed (generic_RVD.h:2179-2181)
            if (active && !sr_ok && nn > 0 && n > 0) lexh = true;

            // TriangleAction fan (generic_RVD.h:452-463) + integration (RVD.cpp:280-296, 575-604)
            if (n >= 3) {
                ++st_ne;
                const double* p1 = P[cur][0];
                double ea = sqrt(dist2<D>(p1, P[cur][1]));
                for (int i = 1; i + 1 < n; ++i) {
                    const double* p2 = P[cur][i];
                    const double* p3 = P[cur][i + 1];
                    ++st_tri;
                    const double eb = sqrt(dist2<D>(p2, p3));
                    const double ec = sqrt(dist2<D>(p3, p1));
                    const double sh = 0.5 * (ea + eb + ec);
                    const double A2 = sh * (sh - ea) * (sh - eb) * (sh - ec);
                    const double area = sqrt(fmax(A2, 0.0));
                    ea = ec;
                    if (a.mode == 0) {
                        if (!WEIGHTED) {
                            const double sc = area / 3.0;
                            acc_s += area;
#pragma unroll
                            for (int c = 0; c < D; ++c) acc_v[c] += sc * (p1[c] + p2[c] + p3[c]);
                        } else {
                            // Geom::triangle_centroid (geometry_nd.h:178-199)
                            const double wa = Wt[cur][0], wb = Wt[cur][i], wc = Wt[cur][i + 1];
                            const double abc = wa + wb + wc;
                            acc_s += area / 3.0 * abc;
                            const double wp = wa + abc, wq = wb + abc, wr = wc + abc;
                            const double sc = area / 12.0;
#pragma unroll
                            for (int c = 0; c < D; ++c) acc_v[c] += sc * (wp * p1[c] + wq * p2[c] + wr * p3[c]);
                        }
                    } else {
                        if (!WEIGHTED) {
                            double cur_f = 0.0;
#pragma unroll
                            for (int c = 0; c < D; ++c) {
                                const double u0 = pi[c] - p1[c];
                                const double u1 = pi[c] - p2[c];
                                const double u2 = pi[c] - p3[c];
                                cur_f += u0 * u0;
                                cur_f += u1 * (u0 + u1);
                                cur_f += u2 * (u0 + u1 + u2);
                            }
                            acc_s += area * cur_f / 6.0;
#pragma unroll
                            for (int c = 0; c < D; ++c) {
                                const double Gc = (1.0 / 3.0) * (p1[c] + p2[c] + p3[c]);
                                acc_v[c] += (2.0 * area) * (pi[c] - Gc);
                            }
                        } else {
                            const double rho0 = Wt[cur][0], rho1 = Wt[cur][i], rho2 = Wt[cur][i + 1];
                            const double Sp = rho0 + rho1 + rho2;
                            const double al0 = Sp + rho0, al1 = Sp + rho1, al2 = Sp + rho2;
                            double d00 = 0, d10 = 0, d11 = 0, d20 = 0, d21 = 0, d22 = 0;
#pragma unroll
                            for (int c = 0; c < D; ++c) {
                                const double sp0 = pi[c] - p1[c], sp1 = pi[c] - p2[c], sp2 = pi[c] - p3[c];
                                d00 += sp0 * sp0; d10 += sp1 * sp0; d11 += sp1 * sp1;
                                d20 += sp2 * sp0; d21 += sp2 * sp1; d22 += sp2 * sp2;
                            }
                            double cur_f = 0.0;
                            cur_f += (al0 + rho0) * d00;
                            cur_f += (al1 + rho0) * d10;
                            cur_f += (al1 + rho1) * d11;
                            cur_f += (al2 + rho0) * d20;
                            cur_f += (al2 + rho1) * d21;
                            cur_f += (al2 + rho2) * d22;
                            acc_s += area * cur_f / 30.0;
#pragma unroll
                            for (int c = 0; c < D; ++c)
                                acc_v[c] += (area / 6.0) * (4.0 * Sp * pi[c] - (al0 * p1[c] + al1 * p2[c] + al2 * p3[c]));
                        }
                    }
                }
            }
        }

        // warp reduction (fixed tree: deterministic for a given candidate list)
        acc_s = warp_sum(acc_s);
#pragma unroll
        for (int c = 0; c < D; ++c) acc_v[c] = warp_sum(acc_v[c]);
        const bool any_exh = __any_sync(B200_FULL, lexh);
        u32 fl = lflags;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) fl |= __shfl_xor_sync(B200_FULL, fl, m);
        if (a.stats) {
            st_planes = (unsigned long long)warp_sum((double)st_planes);
            st_cuts = (unsigned long long)warp_sum((double)st_cuts);
            st_tri = (unsigned long long)warp_sum((double)st_tri);
            st_ne = (unsigned long long)warp_sum((double)st_ne);
        }
        if (lane == 0) {
            a.out_s[s] = acc_s;
#pragma unroll
            for (int c = 0; c < D; ++c) a.out_v[(size_t)s * D + c] = acc_v[c];
            uint8_t f8 = (uint8_t)(a.flags[s] & ~(uint8_t)(1 | 4 | 8));
            f8 |= (uint8_t)fl;
            if (any_exh) {
                if (!a.check_SR) f8 |= 1;                 // Lloyd: cell truncated to the stored list
                else if (nn + 1 >= a.S) { }               // every other seed already used: exact
                else if (nn >= B200CVT_KMAX_DEV) f8 |= 8; // implementation cap
                else if (a.redo_list) { u32 pos = atomicAdd(a.redo_n, 1u); a.redo_list[pos] = s; }
            }
            a.flags[s] = f8;
            if (a.stats) {
                atomicAdd(&a.stats[0], st_planes); atomicAdd(&a.stats[1], st_cuts);
                atomicAdd(&a.stats[2], st_tri); atomicAdd(&a.stats[3], st_ne);
            }
        }
    }
}
