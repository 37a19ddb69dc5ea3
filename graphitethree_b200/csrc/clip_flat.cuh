// clip_flat.cuh — the main clip + integrate path: one thread per candidate (facet, seed) pair,
// pairs grouped by the number of bisectors that cut them.
//
// Same arithmetic as clip.cuh (clip_by_cell_SR generic_RVD.h:2134-2199, clip_by_plane_fast
// generic_RVD_polygon.h:241-366, integration RVD.cpp:254-369 / 543-724) — the polygons are
// bit-identical to the reference's — but a different mapping:
//
//   compact_pairs_kernel  : one warp per seed sorts its candidate row by facet id and writes it
//                           into the flat pair arrays (seed-major, facets ascending). The rows carry, for every
//                           pair, the bit mask of the bisectors that may cut the UNCLIPPED facet (facet_pairs.cuh:
//                           conservative FP32 scan); no bit = the cell contains the facet = the pair is integrated
//                           at once from the precomputed facet area.
//   clip_win_kernel       : a block takes a window of 512 consecutive pairs, counting-sorts it by the number of
//                           masked bisectors in shared memory (warps of pairs with equal work), then every thread
//                           applies exactly the reference's loop restricted to the masked bisectors (the others
//                           cannot change the polygon), radius test with the CURRENT polygon before each.
//                           The polygon lives in shared memory, lane-interleaved
//                           ([vertex][coord][lane]: lanes indexing different vertices never
//                           conflict), and is clipped IN PLACE: both intersection points of a cut
//                           are built first, then the vertices are moved in
//                           the reference's emission order (the write index is never more than one
//                           slot ahead of the read index; the next vertex is held in registers).
//   reduce_pairs_kernel   : one warp per seed sums its pairs' contributions in facet order with a fixed tree —
//                           deterministic, independent of the partition and of atomics order.
//
// Pairs the fast path cannot finish (more than CLIPF_MAXV vertices, more than two crossings of one
// plane = numerically non-convex) are marked PSTAT_SLOW; their seeds are re-evaluated by the
// warp-per-seed kernel of clip.cuh (local-memory ping-pong buffers).
#pragma once
#include "common.cuh"
#include "clip.cuh"

#define CLIPF_MAXV 11

#define PSTAT_EXHAUSTED 1u
#define PSTAT_SLOW 16u

// ---------------------------------------------------------------------------------------
// rows -> flat, sorted
// ---------------------------------------------------------------------------------------
struct CompactArgs {
    const u32* pair_cnt;      // [S] sorted order
    const u32* pair_facet;    // [S][cap]
    const u32* pair_mask;     // [S][cap] bisectors that may cut the pair | PMASK_SR_OK
    u32 cap;
    const u32* pair_off;      // [nown+1] exclusive scan of min(cnt, cap) over the owned range
    u32 qbegin, nown;
    u32* flat_seed;           // [npairs] sorted position of the seed
    u32* flat_facet;          // [npairs]
    u32* flat_mask;           // [npairs]
};

#define PCLASS_MAXCUT 7u

__global__ void __launch_bounds__(256)
compact_pairs_kernel(CompactArgs a) {
    const int lane = threadIdx.x & 31;
    const u32 wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 i = wid; i < a.nown; i += nwarps) {
        const u32 s = a.qbegin + i;
        const u32 n = min(a.pair_cnt[s], a.cap);
        if (n == 0) continue;
        const u32 off = a.pair_off[i];
        const u32* row = a.pair_facet + (size_t)s * a.cap;
        const u32* mrow = a.pair_mask + (size_t)s * a.cap;
        if (n <= 32) {
            u64 v = lane < n ? (((u64)row[lane] << 32) | mrow[lane]) : ~0ull;
#pragma unroll
            for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
                for (int j = k >> 1; j > 0; j >>= 1) {
                    u64 o = __shfl_xor_sync(B200_FULL, v, j);
                    bool up = ((lane & k) == 0), lower = ((lane & j) == 0);
                    v = (lower == up) ? min(v, o) : max(v, o);
                }
            if (lane < n) {
                const u32 m = (u32)v;
                a.flat_facet[off + lane] = (u32)(v >> 32); a.flat_seed[off + lane] = s; a.flat_mask[off + lane] = m;
            }
        } else {
            // rank by counting (facet ids of one seed are distinct)
            for (u32 t = lane; t < n; t += 32) {
                const u32 v = row[t], m = mrow[t];
                u32 r = 0;
                for (u32 u = 0; u < n; ++u) r += (row[u] < v) ? 1u : 0u;
                a.flat_facet[off + r] = v; a.flat_seed[off + r] = s; a.flat_mask[off + r] = m;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// integration of one fan triangle (p1, p2, p3) of seed pi, given its area:
// ComputeCentroids / ComputeCVTFuncGrad (+Weighted) — RVD.cpp:280-296, 322-369, 575-604, 640-724
// ---------------------------------------------------------------------------------------
template <int D, bool WEIGHTED>
__device__ __forceinline__ void integrate_triangle(const double* p1, const double* p2, const double* p3, double area,
                                                   const double* pi, int mode, double& acc_s, double* acc_v) {
    constexpr int VW = D + (WEIGHTED ? 1 : 0);
    if (mode == 0) {
        if (!WEIGHTED) {
            const double sc = area / 3.0;
            acc_s += area;
#pragma unroll
            for (int c = 0; c < D; ++c) acc_v[c] += sc * (p1[c] + p2[c] + p3[c]);
        } else {
            // Geom::triangle_centroid (geometry_nd.h:178-199)
            const double wa = p1[VW - 1], wb = p2[VW - 1], wc = p3[VW - 1];
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
            const double rho0 = p1[VW - 1], rho1 = p2[VW - 1], rho2 = p3[VW - 1];
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

// Geom::triangle_area (geometry_nd.h:143-156): Heron on nD edge lengths, clamped
template <int D>
__device__ __forceinline__ double heron_area(double ea, double eb, double ec) {
    const double sh = 0.5 * (ea + eb + ec);
    const double A2 = sh * (sh - ea) * (sh - eb) * (sh - ec);
    return sqrt(fmax(A2, 0.0));
}

// area of every mesh facet, once per mesh (the fan of an unclipped facet is the facet itself)
template <int D>
__global__ void facet_area_kernel(const double* tri, u32 T, double* area) {
    u32 f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= T) return;
    const double* t = tri + (size_t)f * 3 * D;
    double p[3][D];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int c = 0; c < D; ++c) p[i][c] = t[i * D + c];
    // same evaluation order as the fan loop: ea = |p1 p2|, eb = |p2 p3|, ec = |p3 p1|
    const double ea = sqrt(dist2<D>(p[0], p[1]));
    const double eb = sqrt(dist2<D>(p[1], p[2]));
    const double ec = sqrt(dist2<D>(p[2], p[0]));
    area[f] = heron_area<D>(ea, eb, ec);
}

// ---------------------------------------------------------------------------------------
// bisector table: for every seed i and every stored neighbour j (list order = increasing distance)
//   n = pi - pj, d = sum_c (pi[c] + pj[c]) * n[c]   (generic_RVD_polygon.h:257-274), dij = |pi - pj|^2
// computed once per evaluation instead of once per (facet, seed) pair. Row = PLANE_STRIDE doubles.
// ---------------------------------------------------------------------------------------
#define PLANE_STRIDE(D) ((D) == 3 ? 6 : (D) + 2)

template <int D>
__global__ void __launch_bounds__(256)
plane_table_kernel(const void* xs_, const u32* nbr, const u32* nbr_n, u32 kstride, const u32* seed_list, u32 qbegin, u32 nseeds,
                   const u32* nseeds_dev, double* planes, float* planes32) {
    constexpr int PS = PLANE_STRIDE(D);
    const SeedRec<D>* xs = (const SeedRec<D>*)xs_;
    if (nseeds_dev) nseeds = *nseeds_dev;
    const size_t total = (size_t)nseeds * kstride;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const u32 i = (u32)(e / kstride), jj = (u32)(e - (size_t)i * kstride);
        const u32 s = seed_list ? seed_list[i] : qbegin + i;
        if (jj >= min(nbr_n[s], kstride)) continue;
        double pi[D], pj[D];
        const SeedRec<D>* rj = xs + nbr[(size_t)s * kstride + jj];
#pragma unroll
        for (int c = 0; c < D; ++c) { pi[c] = xs[s].p[c]; pj[c] = rj->p[c]; }
        double* o = planes + ((size_t)s * kstride + jj) * PS;
        double d = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const double nc = pi[c] - pj[c];
            o[c] = nc;
            d += (pi[c] + pj[c]) * nc;
        }
        o[D] = d;
        const double dij = dist2<D>(pi, pj);
        o[D + 1] = dij;
        if (planes32) {
            constexpr int PS32 = (D == 3) ? 4 : 8;
            float* o32 = planes32 + ((size_t)s * kstride + jj) * PS32;
#pragma unroll
            for (int c = 0; c < D; ++c) o32[c] = (float)(pi[c] - pj[c]);
            o32[D] = (float)dij;
        }
    }
}

// ---------------------------------------------------------------------------------------
// classification: which bisectors can cut the facet at all
// ---------------------------------------------------------------------------------------
struct ClipFlatArgs {
    const void* xs;
    const u32* nbr; const u32* nbr_n; u32 kstride;
    const double* planes;       // [S][kstride][PLANE_STRIDE] bisector table
    const double* tri;          // [T][3][D]
    const double* triw;         // [T][3] or NULL
    const double* facet_area;   // [T]
    const u32* flat_seed; const u32* flat_facet;
    const u32* npairs_dev;      // device scalar: number of flat pairs
    int mode;                   // 0: m, mg   1: f_seed, g
    double* contrib;            // [(1+D)][cstride]
    size_t cstride;
    uint8_t* pstat;             // [npairs]
    const u32* flat_mask;       // [npairs] bisectors (positions in the neighbour list) that may cut the pair | PMASK_SR_OK
    unsigned long long* stats;  // optional
};

// ---------------------------------------------------------------------------------------
// clipping of the pairs with at least one masked bisector
// ---------------------------------------------------------------------------------------
struct ClipStats { unsigned long long planes = 0, cuts = 0, tri = 0, ne = 0; };

// one candidate pair t: clip facet f by the masked bisectors of seed s and integrate. P = this thread's polygon
// (lane-interleaved shared memory, stride 32 doubles between consecutive values)
template <int D, bool WEIGHTED>
__device__ __forceinline__ void clip_one_pair(const ClipFlatArgs& a, const u32 t, double* P, ClipStats& st) {
    constexpr int VW = D + (WEIGHTED ? 1 : 0);
    constexpr int PS = PLANE_STRIDE(D);
#define PV(k, c) P[((k) * VW + (c)) * 32]
    const SeedRec<D>* xs = (const SeedRec<D>*)a.xs;
    const u32 s = a.flat_seed[t];
    const u32 f = a.flat_facet[t];
    const u32 mask_in = a.flat_mask[t];
    u32 mask = mask_in & 0x7fffffffu;
    double pi[D];
#pragma unroll
    for (int c = 0; c < D; ++c) pi[c] = xs[s].p[c];
    const u32 nn = min(min(a.nbr_n[s], a.kstride), 32u);
    const double* prow = a.planes + (size_t)s * a.kstride * PS;
    int n = 3;
    double R2 = 0.0;
    {
        const double* tp = a.tri + (size_t)f * 3 * D;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double v[D];
#pragma unroll
            for (int c = 0; c < D; ++c) { v[c] = tp[i * D + c]; PV(i, c) = v[c]; }
            if (WEIGHTED) PV(i, D) = a.triw[(size_t)f * 3 + i];
            R2 = fmax(R2, dist2<D>(pi, v));
        }
    }
    // no masked bisector: the cell contains the facet. The classification (FP32, conservative) may have certified the radius
    // test on the unclipped facet; if it did not, the test is redone exactly below with the last neighbour
    bool sr_ok = (mask == 0) && (mask_in & 0x80000000u), slow = false, cut_any = false;
    int last_jj = -1;
    // clip_by_cell_SR (generic_RVD.h:2155-2177) over the masked bisectors, increasing distance
    while (mask) {
        const int jj = __ffs(mask) - 1;
        mask &= mask - 1;
        double2 rowbuf[PS / 2];
        {
            const double2* r2 = (const double2*)(prow + (size_t)jj * PS);
#pragma unroll
            for (int q = 0; q < PS / 2; ++q) rowbuf[q] = __ldg(r2 + q);
        }
        const double* pl = (const double*)rowbuf;
        if (pl[D + 1] > 4.1 * R2) { sr_ok = true; break; }
        last_jj = jj;
        ++st.planes;
        double nj[D];
#pragma unroll
        for (int c = 0; c < D; ++c) nj[c] = pl[c];
        const double d = pl[D];
        // pass 1: side of every vertex (generic_RVD_polygon.h:276-297)
        u32 pos = 0, neg = 0;
        for (int k = 0; k < n; ++k) {
            double l = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) l += PV(k, c) * nj[c];
            const double tk = 2.0 * l - d;
            pos |= (tk > 0.0 ? 1u : 0u) << k;
            neg |= (tk < 0.0 ? 1u : 0u) << k;
        }
        const u32 full = (1u << n) - 1u;
        if (pos == full) continue;                 // nothing to cut
        ++st.cuts;
        cut_any = true;
        // pass 2 (generic_RVD_polygon.h:299-362). A crossing is emitted where the side changes and the
        // previous vertex is not on the plane; a convex polygon has at most two.
        const u32 ppos = ((pos << 1) | (pos >> (n - 1))) & full;   // bit k = side of vertex k-1
        const u32 pneg = ((neg << 1) | (neg >> (n - 1))) & full;
        const u32 X = (ppos | pneg) & ((pos ^ ppos) | (neg ^ pneg));
        const int nx = __popc(X);
        if (nx > 2 || __popc(pos) + nx > CLIPF_MAXV) { slow = true; break; }
        double I[2][VW];
        int kx0 = -1, kx1 = -1;
        {
            u32 xr = X;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
#pragma unroll
                for (int c = 0; c < VW; ++c) I[q][c] = 0.0;
                if (xr) {
                    const int k = __ffs(xr) - 1;
                    xr &= xr - 1;
                    if (q == 0) kx0 = k; else kx1 = k;
                    const int kp = (k == 0) ? n - 1 : k - 1;
                    double vp[VW], vc[VW];
#pragma unroll
                    for (int c = 0; c < VW; ++c) { vp[c] = PV(kp, c); vc[c] = PV(k, c); }
                    double lp = 0.0, l = 0.0;
#pragma unroll
                    for (int c = 0; c < D; ++c) { lp += vp[c] * nj[c]; l += vc[c] * nj[c]; }
                    const double denom = 2.0 * (lp - l);
                    double l1, l2;
                    if (fabs(denom) < 1e-20) { l1 = 0.5; l2 = 0.5; }
                    else { l1 = (d - 2.0 * l) / denom; l2 = 1.0 - l1; }
#pragma unroll
                    for (int c = 0; c < VW; ++c) I[q][c] = l1 * vp[c] + l2 * vc[c];
                }
            }
        }
        int m = 0;
        double R2n = 0.0;
        double vc[VW], vn[VW];
#pragma unroll
        for (int c = 0; c < VW; ++c) { vc[c] = PV(0, c); vn[c] = 0.0; }
        for (int k = 0; k < n; ++k) {
            if (k + 1 < n) {
#pragma unroll
                for (int c = 0; c < VW; ++c) vn[c] = PV(k + 1, c);
            }
            if (k == kx0 || k == kx1) {
                // m <= k + 1: at most one crossing precedes without a dropped vertex
                double Iq[VW];
#pragma unroll
                for (int c = 0; c < VW; ++c) { Iq[c] = (k == kx0) ? I[0][c] : I[1][c]; PV(m, c) = Iq[c]; }
                R2n = fmax(R2n, dist2<D>(pi, Iq));
                ++m;
            }
            if ((pos >> k) & 1u) {
                if (m != k) {
#pragma unroll
                    for (int c = 0; c < VW; ++c) PV(m, c) = vc[c];
                }
                R2n = fmax(R2n, dist2<D>(pi, vc));
                ++m;
            }
#pragma unroll
            for (int c = 0; c < VW; ++c) vc[c] = vn[c];
        }
        n = m;
        R2 = R2n;
        if (n == 0) break;
    }

    double acc_s = 0.0, acc_v[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc_v[c] = 0.0;
    uint8_t ps = 0;
    if (slow) ps = PSTAT_SLOW;
    else {
        if (!sr_ok && n > 0 && nn > 0) {
            // the reference goes on testing the remaining neighbours (unmasked: they cannot cut);
            // the list is sorted, so the radius test passes for one of them iff it passes for the last
            if (last_jj < (int)nn - 1) sr_ok = prow[(size_t)(nn - 1) * PS + D + 1] > 4.1 * R2;
            // list used up before the radius test passed (generic_RVD.h:2179-2181)
            if (!sr_ok) ps = PSTAT_EXHAUSTED;
        }
        if (n >= 3) {
            ++st.ne;
            if (!cut_any) {
                double p1[VW], p2[VW], p3[VW];
#pragma unroll
                for (int c = 0; c < VW; ++c) { p1[c] = PV(0, c); p2[c] = PV(1, c); p3[c] = PV(2, c); }
                ++st.tri;
                integrate_triangle<D, WEIGHTED>(p1, p2, p3, a.facet_area[f], pi, a.mode, acc_s, acc_v);
            } else {
                // TriangleAction fan (generic_RVD.h:452-463)
                double p1[VW], p2[VW], p3[VW];
#pragma unroll
                for (int c = 0; c < VW; ++c) { p1[c] = PV(0, c); p3[c] = PV(1, c); }
                double ea = sqrt(dist2<D>(p1, p3));
                for (int i = 1; i + 1 < n; ++i) {
#pragma unroll
                    for (int c = 0; c < VW; ++c) { p2[c] = p3[c]; p3[c] = PV(i + 1, c); }
                    ++st.tri;
                    const double eb = sqrt(dist2<D>(p2, p3));
                    const double ec = sqrt(dist2<D>(p3, p1));
                    const double area = heron_area<D>(ea, eb, ec);
                    ea = ec;
                    integrate_triangle<D, WEIGHTED>(p1, p2, p3, area, pi, a.mode, acc_s, acc_v);
                }
            }
        }
    }
    a.contrib[t] = acc_s;
#pragma unroll
    for (int c = 0; c < D; ++c) a.contrib[(size_t)(c + 1) * a.cstride + t] = acc_v[c];
    a.pstat[t] = ps;
#undef PV
}

// Pairs are taken in windows of CLIPW_W consecutive entries of the seed-major pair list (= a few dozen
// Morton-neighbouring seeds and the facets around them: facet corners, bisector rows and the per-pair outputs of a
// window stay in L1/L2). Inside a window the pairs are counting-sorted by class (number of bisectors that may cut
// them) in shared memory, so that the lanes of a warp run the same number of plane iterations; each thread then
// clips CLIPW_ROUNDS pairs, one from each quarter of the sorted window (equal work per warp).
#define CLIPW_THREADS 128
#ifndef CLIPW_ROUNDS
#define CLIPW_ROUNDS 4          // measured at C2: 4 rounds 0.429 ms, 6 rounds 0.423 ms, 8 rounds 0.461 ms (shared memory)
#endif
#define CLIPW_W (CLIPW_THREADS * CLIPW_ROUNDS)
#define CLIPW_NW (CLIPW_THREADS / 32)
#define CLIPW_NCNT (8 * CLIPW_ROUNDS * CLIPW_NW)

#ifndef CLIPW_MINBLK
#define CLIPW_MINBLK 6
#endif
template <int D, bool WEIGHTED>
__global__ void __launch_bounds__(CLIPW_THREADS, (D == 3 && !WEIGHTED) ? CLIPW_MINBLK : 1)
clip_win_kernel(ClipFlatArgs a) {
    constexpr int VW = D + (WEIGHTED ? 1 : 0);
    extern __shared__ double s_dyn[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    double* P = s_dyn + (size_t)w * (CLIPF_MAXV * VW * 32) + lane;
    unsigned short* order = (unsigned short*)(s_dyn + (size_t)CLIPW_NW * CLIPF_MAXV * VW * 32);   // [CLIPW_W]
    u32* ccnt = (u32*)(order + CLIPW_W);                                                          // [class][round][warp]
    constexpr int CPL = CLIPW_NCNT / 32;      // counters per lane in the scan below
    static_assert(CLIPW_NCNT % 32 == 0 && CLIPW_ROUNDS <= 8, "counter scan by one warp; class/rank packing");
    const u32 npairs = *a.npairs_dev;
    const u32 lt = (1u << lane) - 1u;
    ClipStats st;
    for (u32 base = blockIdx.x * CLIPW_W; base < npairs; base += gridDim.x * CLIPW_W) {
        const u32 cnt = min((u32)CLIPW_W, npairs - base);
        for (int i = tid; i < CLIPW_NCNT; i += CLIPW_THREADS) ccnt[i] = 0;
        __syncthreads();
        // class and rank inside (class, round, warp) of every pair of the window
        u32 pcls = 0; u64 prank = 0;
#pragma unroll
        for (int r = 0; r < CLIPW_ROUNDS; ++r) {
            const u32 idx = r * CLIPW_THREADS + tid;
            u32 cls = 8;
            if (idx < cnt) cls = min((u32)__popc(a.flat_mask[base + idx] & 0x7fffffffu), PCLASS_MAXCUT);
            const u32 m = __match_any_sync(B200_FULL, cls);
            const u32 rank = __popc(m & lt);
            if (cls < 8 && rank == 0) ccnt[(cls * CLIPW_ROUNDS + r) * CLIPW_NW + w] = __popc(m);
            pcls |= cls << (4 * r); prank |= (u64)rank << (8 * r);
        }
        __syncthreads();
        if (w == 0) {
            // exclusive scan of the counters in (class, round, warp) order = stable counting sort
            u32 v[CPL], sum = 0;
#pragma unroll
            for (int i = 0; i < CPL; ++i) { v[i] = ccnt[lane * CPL + i]; sum += v[i]; }
            u32 incl = sum;
#pragma unroll
            for (int m = 1; m < 32; m <<= 1) {
                const u32 o = __shfl_up_sync(B200_FULL, incl, m);
                if (lane >= m) incl += o;
            }
            u32 run = incl - sum;
#pragma unroll
            for (int i = 0; i < CPL; ++i) { ccnt[lane * CPL + i] = run; run += v[i]; }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < CLIPW_ROUNDS; ++r) {
            const u32 cls = (pcls >> (4 * r)) & 15u, rank = (u32)(prank >> (8 * r)) & 255u;
            if (cls < 8) order[ccnt[(cls * CLIPW_ROUNDS + r) * CLIPW_NW + w] + rank] = (unsigned short)(r * CLIPW_THREADS + tid);
        }
        __syncthreads();
        for (int r = 0; r < CLIPW_ROUNDS; ++r) {
            const u32 p = r * CLIPW_THREADS + tid;
            if (p < cnt) clip_one_pair<D, WEIGHTED>(a, base + order[p], P, st);
        }
        __syncthreads();
    }
    if (a.stats) {
        st.planes = (unsigned long long)warp_sum((double)st.planes);
        st.cuts = (unsigned long long)warp_sum((double)st.cuts);
        st.tri = (unsigned long long)warp_sum((double)st.tri);
        st.ne = (unsigned long long)warp_sum((double)st.ne);
        if (lane == 0) {
            atomicAdd(&a.stats[8], st.planes); atomicAdd(&a.stats[1], st.cuts);
            atomicAdd(&a.stats[2], st.tri); atomicAdd(&a.stats[3], st.ne);
        }
    }
}

// ---------------------------------------------------------------------------------------
// per-seed sums, in facet order
// ---------------------------------------------------------------------------------------
struct ReduceArgs {
    const u32* pair_off; u32 qbegin, nown;
    const double* contrib; size_t cstride;
    const uint8_t* pstat;
    const u32* nbr_n; u32 kstride;
    int check_SR; u32 S;
    double* out_s; double* out_v; uint8_t* flags;
    u32* slow_list; u32* slow_n;     // seeds to re-evaluate with the warp-per-seed kernel
    u32* redo_list; u32* redo_n;     // seeds whose neighbour list must grow (check_SR)
};

// one warp per seed: lanes take the seed's pairs round-robin (in facet order), then a fixed xor tree
template <int D>
__global__ void __launch_bounds__(256)
reduce_pairs_kernel(ReduceArgs a) {
    const int lane = threadIdx.x & 31;
    const u32 i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= a.nown) return;
    const u32 s = a.qbegin + i;
    const u32 b = a.pair_off[i], e = a.pair_off[i + 1];
    double acc_s = 0.0, acc_v[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc_v[c] = 0.0;
    u32 ps = 0;
    for (u32 t = b + lane; t < e; t += 32) {
        acc_s += a.contrib[t];
#pragma unroll
        for (int c = 0; c < D; ++c) acc_v[c] += a.contrib[(size_t)(c + 1) * a.cstride + t];
        ps |= a.pstat[t];
    }
    acc_s = warp_sum(acc_s);
#pragma unroll
    for (int c = 0; c < D; ++c) acc_v[c] = warp_sum(acc_v[c]);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) ps |= __shfl_xor_sync(B200_FULL, ps, m);
    if (lane != 0) return;
    a.out_s[s] = acc_s;
#pragma unroll
    for (int c = 0; c < D; ++c) a.out_v[(size_t)s * D + c] = acc_v[c];
    uint8_t f8 = (uint8_t)(a.flags[s] & ~(uint8_t)(1 | 4 | 8));
    if (ps & PSTAT_SLOW) {
        u32 p = atomicAdd(a.slow_n, 1u);
        a.slow_list[p] = s;
    } else if (ps & PSTAT_EXHAUSTED) {
        const u32 nn = min(a.nbr_n[s], a.kstride);
        if (!a.check_SR) f8 |= 1;
        else if (nn + 1 >= a.S) { }
        else if (nn >= B200CVT_KMAX_DEV) f8 |= 8;
        else { u32 p = atomicAdd(a.redo_n, 1u); a.redo_list[p] = s; }
    }
    a.flags[s] = f8;
}
