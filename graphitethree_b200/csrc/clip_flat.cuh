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
// Pairs the fast path cannot finish (more than CLIPF_CAP vertices, more than two crossings of one
// plane = numerically non-convex) are marked PSTAT_SLOW; their seeds are re-evaluated by the
// warp-per-seed kernel of clip.cuh (local-memory ping-pong buffers).
#pragma once
#include "common.cuh"
#include "clip.cuh"


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
    // persistent warps; the row of the next seed is requested before the current one is sorted (the kernel is latency-bound)
    u32 n = 0, off = 0;
    u64 v = ~0ull;
    auto fetch = [&](u32 i, u32& n_, u32& off_, u64& v_) {
        n_ = 0; off_ = 0; v_ = ~0ull;
        if (i < a.nown) {
            const u32 s = a.qbegin + i;
            n_ = min(a.pair_cnt[s], a.cap);
            off_ = a.pair_off[i];
            if ((u32)lane < n_ && n_ <= 32) v_ = ((u64)a.pair_facet[(size_t)s * a.cap + lane] << 32) | a.pair_mask[(size_t)s * a.cap + lane];
        }
    };
    fetch(wid, n, off, v);
    for (u32 i = wid; i < a.nown; i += nwarps) {
        u32 n2, off2; u64 v2;
        fetch(i + nwarps, n2, off2, v2);
        const u32 s = a.qbegin + i;
        if (n > 0 && n <= 32) {
#pragma unroll
            for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
                for (int j = k >> 1; j > 0; j >>= 1) {
                    u64 o = __shfl_xor_sync(B200_FULL, v, j);
                    bool up = ((lane & k) == 0), lower = ((lane & j) == 0);
                    v = (lower == up) ? min(v, o) : max(v, o);
                }
            if ((u32)lane < n) {
                const u32 m = (u32)v;
                a.flat_facet[off + lane] = (u32)(v >> 32); a.flat_seed[off + lane] = s; a.flat_mask[off + lane] = m;
            }
        } else if (n > 32) {
            // rank by counting (facet ids of one seed are distinct)
            const u32* row = a.pair_facet + (size_t)s * a.cap;
            const u32* mrow = a.pair_mask + (size_t)s * a.cap;
            for (u32 t = lane; t < n; t += 32) {
                const u32 fv = row[t], m = mrow[t];
                u32 r = 0;
                for (u32 u = 0; u < n; ++u) r += (row[u] < fv) ? 1u : 0u;
                a.flat_facet[off + r] = fv; a.flat_seed[off + r] = s; a.flat_mask[off + r] = m;
            }
        }
        n = n2; off = off2; v = v2;
    }
}

// Correctly rounded x / N for the constants of the integration formulas (3, 6, 12, 30): q = RN(x y), y = RN(1/N),
// r = x - N q (exact, one FMA), result RN(q + r y) — Markstein's correction step; bit-equal to the IEEE division the
// reference executes (checked against x / N on 4e8 operands incl. boundary mantissas), 3 instructions instead of ~45.
template <int N>
__device__ __forceinline__ double div_exact(double x) {
    const double y = 1.0 / (double)N;
    const double q = x * y;
    const double r = __fma_rn(-(double)N, q, x);
    return __fma_rn(r, y, q);
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
            const double sc = div_exact<3>(area);
            acc_s += area;
#pragma unroll
            for (int c = 0; c < D; ++c) acc_v[c] += sc * (p1[c] + p2[c] + p3[c]);
        } else {
            // Geom::triangle_centroid (geometry_nd.h:178-199)
            const double wa = p1[VW - 1], wb = p2[VW - 1], wc = p3[VW - 1];
            const double abc = wa + wb + wc;
            acc_s += div_exact<3>(area) * abc;
            const double wp = wa + abc, wq = wb + abc, wr = wc + abc;
            const double sc = div_exact<12>(area);
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
            acc_s += div_exact<6>(area * cur_f);
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
            acc_s += div_exact<30>(area * cur_f);
#pragma unroll
            for (int c = 0; c < D; ++c)
                acc_v[c] += div_exact<6>(area) * (4.0 * Sp * pi[c] - (al0 * p1[c] + al1 * p2[c] + al2 * p3[c]));
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
    const float* planes32;      // [S][kstride][4 | 8] FP32 filter copy (n, |n|^2), facet_pairs.cuh
    double vmax2;               // max |v|^2 over the mesh corners (bound of the reference's FP64 rounding)
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

// ---- packed-permutation helpers (4 bits per entry, 8 entries) ----
__device__ __forceinline__ u32 nibmask(int L) { return L >= 8 ? 0xffffffffu : ((1u << (4 * L)) - 1u); }
__device__ __forceinline__ u32 shr4(u32 x, int m) { return m >= 8 ? 0u : (x >> (4 * m)); }
__device__ __forceinline__ u32 shl4(u32 x, int m) { return m >= 8 ? 0u : (x << (4 * m)); }
// entries of `perm` at the set positions of `bits`, in index order. The set must be at most two runs, one of them
// starting at position 0 (a cyclic interval of [0, n) or its complement); *ok is cleared otherwise.
__device__ __forceinline__ u32 perm_compress(u32 perm, u32 bits, bool* ok) {
    const int t = __ffs(~bits) - 1;                 // trailing ones
    const u32 low = perm & nibmask(t);
    const u32 rest = bits >> t;                      // bit 0 is clear
    if (rest == 0) return low;
    const int a = __ffs(rest) - 1;
    const u32 run = rest >> a;
    if (run & (run + 1u)) *ok = false;               // not one contiguous run
    const int L = __popc(rest);
    const u32 high = shr4(perm, t + a) & nibmask(L);
    return low | shl4(high, t);
}
__device__ __forceinline__ u32 perm_insert(u32 K, int m, u32 v) {
    return (K & nibmask(m)) | shl4(v, m) | shl4(shr4(K, m), m + 1);
}

// The polygon of one pair lives in CLIPF_CAP vertex slots of shared memory, lane-interleaved (lanes that index
// different slots never conflict):
//   P : FP64 coordinates (+ weight)            [slot][coord][lane]
//   Q : FP32 shadow q = v - p_seed, w = |q|^2   [slot][lane] float4 (two float4 for D = 6)
// Vertices never move. The cyclic order of the polygon is a packed permutation (4 bits per logical vertex -> slot) in a
// register, entry 0 = the reference's vertex 0 (the fan apex of the integration); a cut only writes the (at most two)
// new vertices into free slots and rebuilds the permutation in the reference's emission order
// (generic_RVD_polygon.h:299-362).
//
// Side tests (the reference's sgn(2 v.n - d), generic_RVD_polygon.h:276-297) and the radius test
// (generic_RVD.h:2155-2174) are decided by an FP32 filter on seed-local coordinates: 2 v.n - d = 2 q.n + |n|^2 with
// n = p_i - p_j (see facet_pairs.cuh: same filter table). A value beyond the margin has the sign of the reference's own
// FP64 evaluation (margin = FP32 evaluation error + the reference's FP64 rounding on global coordinates, both bounded
// above); anything inside the margin is re-evaluated with the reference's FP64 expression. Intersections, areas and
// integrals are FP64 and follow the reference operation by operation, so the polygons stay bit-identical.
// Measured at C2 (profiles/r2_clip_variants.md): the kernel is latency-bound, its time follows the number of resident warps
// (4 blocks 0.55 ms, 5 blocks 0.38 ms, 6 blocks 0.35 ms). 8 slots + 96 registers = 5 blocks per SM without spills; 7 slots
// + 80 registers = 6 blocks is 7 % faster but needs a second pass for the 8-vertex polygons that eats the gain; 6 slots +
// 72 registers = 7 blocks spills (0.40 ms). Polygons with more vertices than slots take the warp-per-seed path (clip.cuh).
#ifndef CLIPF_CAP
#define CLIPF_CAP 8
#endif
#define CLIPF_QW(D) ((D) == 3 ? 1 : 2)

template <int D, bool WEIGHTED, int CAP>
__device__ __forceinline__ void clip_one_pair(const ClipFlatArgs& a, const u32 t, const u32 s, const u32 f, const u32 mask_in,
                                              double* P, float4* Q, ClipStats& st) {
    constexpr int VW = D + (WEIGHTED ? 1 : 0);
    constexpr int PS = PLANE_STRIDE(D);
    constexpr int PS32 = (D == 3) ? 4 : 8;
    constexpr int QW = CLIPF_QW(D);
#define PV(k, c) P[((k) * VW + (c)) * 32]
#define QV(k, h) Q[((k) * QW + (h)) * 32]
#define SLOT(k) ((perm >> (4 * (k))) & 7u)
    const SeedRec<D>* xs = (const SeedRec<D>*)a.xs;
    u32 mask = mask_in & 0x7fffffffu;
    double pi[D];
    double pi2 = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) { pi[c] = xs[s].p[c]; pi2 += pi[c] * pi[c]; }
    const u32 nn = min(min(a.nbr_n[s], a.kstride), 32u);
    const double* prow = a.planes + (size_t)s * a.kstride * PS;
    const float* prow32 = a.planes32 + (size_t)s * a.kstride * PS32;
    // the reference's FP64 rounding on global coordinates (facet_pairs.cuh: 4e-15 (|v|^2 + |p_i|^2)), v anywhere on the mesh
    const float mextra = __double2float_ru(4e-15 * (a.vmax2 + pi2));
    // bisector rows of a masked neighbour (-DCLIP_ROWPIPE requests them one iteration ahead: measured slower, the 16 extra
    // registers cost more than the hidden latency gains)
    float4 nrow32[PS32 / 4];
    double2 nrow64[PS / 2];
    auto request_rows = [&](int jj) {
        const float4* r4 = (const float4*)(prow32 + (size_t)jj * PS32);
#pragma unroll
        for (int q = 0; q < PS32 / 4; ++q) nrow32[q] = __ldg(r4 + q);
        const double2* r2 = (const double2*)(prow + (size_t)jj * PS);
#pragma unroll
        for (int q = 0; q < PS / 2; ++q) nrow64[q] = __ldg(r2 + q);
    };
#ifdef CLIP_ROWPIPE
    if (mask) request_rows(__ffs(mask) - 1);
#endif
    int n = 3;
    u32 perm = 0x76543210u & nibmask(CAP);   // entries [0, n): the polygon in cyclic order; entries [n, 8): the free slots
    // FP32 radius^2 of the current polygon about the seed: exact (to FP32) while r2_valid, else R2lo <= radius^2 <= R2f
    float R2f = 0.0f, R2lo = 0.0f;
    bool r2_valid = true;
    {
        const double* tp = a.tri + (size_t)f * 3 * D;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float qf[D];
            float w = 0.0f;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const double v = tp[i * D + c];
                PV(i, c) = v;
                qf[c] = (float)(v - pi[c]);
                w = fmaf(qf[c], qf[c], w);
            }
            if (WEIGHTED) PV(i, D) = a.triw[(size_t)f * 3 + i];
            if (D == 3) QV(i, 0) = make_float4(qf[0], qf[1], qf[2], w);
            else { QV(i, 0) = make_float4(qf[0], qf[1], qf[2], qf[3]); QV(i, 1) = make_float4(qf[D - 2], qf[D - 1], w, 0.0f); }
            R2f = fmaxf(R2f, w);
        }
    }
    // exact radius^2 of the current polygon: max_k |p_i - P_k|^2 as the reference evaluates it
    auto exact_R2 = [&]() {
        double R2 = 0.0;
        for (int k = 0; k < n; ++k) {
            const u32 sl = SLOT(k);
            double v[D];
#pragma unroll
            for (int c = 0; c < D; ++c) v[c] = PV(sl, c);
            R2 = fmax(R2, dist2<D>(pi, v));
        }
        return R2;
    };
    // radius test dij > 4.1 R2 (generic_RVD.h:2170): FP32 filter, exact when the two sides are within 1e-5 of each other
    auto radius_test = [&](float dijf, u32 jj, float r2hi, float r2lo) {
        if (r2lo > 1e-30f && dijf < 1e30f) {
            if (dijf > 4.1f * r2hi * 1.00001f) return true;
            if (dijf < 4.1f * r2lo * 0.99999f) return false;
        }
        return prow[(size_t)jj * PS + D + 1] > 4.1 * exact_R2();
    };
    // no masked bisector: the cell contains the facet. The classification (FP32, conservative) may have certified the radius
    // test on the unclipped facet; if it did not, the test is redone below with the last neighbour
    bool sr_ok = (mask == 0) && (mask_in & 0x80000000u), slow = false, cut_any = false;
    int last_jj = -1;
    // clip_by_cell_SR (generic_RVD.h:2155-2177) over the masked bisectors, increasing distance
    while (mask) {
        const int jj = __ffs(mask) - 1;
        mask &= mask - 1;
        float nf[D], dijf;
        double nj[D], d;
#ifndef CLIP_ROWPIPE
        request_rows(jj);
#endif
        {
            const float4 r0 = nrow32[0];
            if (D == 3) { nf[0] = r0.x; nf[1] = r0.y; nf[2] = r0.z; dijf = r0.w; }
            else {
                const float4 r1 = nrow32[PS32 / 4 - 1];
                nf[0] = r0.x; nf[1] = r0.y; nf[2] = r0.z; nf[3] = r0.w; nf[D - 2] = r1.x; nf[D - 1] = r1.y; dijf = r1.z;
            }
            const double* pl = (const double*)nrow64;
#pragma unroll
            for (int c = 0; c < D; ++c) nj[c] = pl[c];
            d = pl[D];
        }
#ifdef CLIP_ROWPIPE
        if (mask) request_rows(__ffs(mask) - 1);
#endif
        // one pass over the vertices: FP32 side value against its margin, and the radius of the polygon
        u32 pos = 0, neg = 0;
        float r2f = 0.0f;
        u32 pp = perm, bit = 1u;
#ifndef CLIP_SIDE_UNROLL
#pragma unroll 1
#endif
        for (int k = 0; k < n; ++k, pp >>= 4, bit <<= 1) {
            const u32 sl = pp & 7u;
            float qf[D], w;
            const float4 q0 = QV(sl, 0);
            if (D == 3) { qf[0] = q0.x; qf[1] = q0.y; qf[2] = q0.z; w = q0.w; }
            else {
                const float4 q1 = QV(sl, 1);
                qf[0] = q0.x; qf[1] = q0.y; qf[2] = q0.z; qf[3] = q0.w; qf[D - 2] = q1.x; qf[D - 1] = q1.y; w = q1.z;
            }
            float l = 0.0f;
#pragma unroll
            for (int c = 0; c < D; ++c) l = fmaf(qf[c], nf[c], l);
            const float tk = fmaf(2.0f, l, dijf);
            const float margin = fmaf(2e-6f, w + dijf, mextra);
            r2f = fmaxf(r2f, w);
            if (tk > margin) pos |= bit;
            if (tk < -margin) neg |= bit;
        }
        R2f = r2f; r2_valid = true;
        if (radius_test(dijf, (u32)jj, r2f, r2f)) { sr_ok = true; break; }
        last_jj = jj;
        ++st.planes;
        const u32 full = (1u << n) - 1u;
        u32 unc = ~(pos | neg) & full;
        if (unc) {
            // inside the margin: the reference's own expression (generic_RVD_polygon.h:276-297)
            while (unc) {
                const int k = __ffs(unc) - 1;
                unc &= unc - 1;
                const u32 sl = SLOT(k);
                double l = 0.0;
#pragma unroll
                for (int c = 0; c < D; ++c) l += PV(sl, c) * nj[c];
                const double tk = 2.0 * l - d;
                pos |= (tk > 0.0 ? 1u : 0u) << k;
                neg |= (tk < 0.0 ? 1u : 0u) << k;
            }
        }
        if (pos == full) continue;                 // nothing to cut
        ++st.cuts;
        cut_any = true;
        // pass 2 (generic_RVD_polygon.h:299-362). A crossing is emitted where the side changes and the
        // previous vertex is not on the plane; a convex polygon has at most two.
        const u32 ppos = ((pos << 1) | (pos >> (n - 1))) & full;   // bit k = side of vertex k-1
        const u32 pneg = ((neg << 1) | (neg >> (n - 1))) & full;
        const u32 X = (ppos | pneg) & ((pos ^ ppos) | (neg ^ pneg));
        const int nx = __popc(X);
        if (nx > 2 || __popc(pos) + nx > CAP) { slow = true; break; }
        double I[2][VW];
        int kx0 = -1, kx1 = -1;
        float r2_new = 0.0f;
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
                    const u32 sp = SLOT(kp), sc = SLOT(k);
                    double vp[VW], vc[VW];
#pragma unroll
                    for (int c = 0; c < VW; ++c) { vp[c] = PV(sp, c); vc[c] = PV(sc, c); }
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
        // new cyclic order = the reference's emission order: (crossing on edge k-1 -> k), then (vertex k if kept).
        // The kept vertices of a convex polygon are a cyclic interval: at most two runs of the packed permutation.
        bool runs_ok = true;
        const int kc = __popc(pos);
        u32 K = perm_compress(perm, pos, &runs_ok);                       // kept slots, index order
        const u32 Dr = perm_compress(perm, ~pos & full, &runs_ok);         // dropped slots
        if (!runs_ok) { slow = true; break; }
        const u32 pool = Dr | shl4(shr4(perm, n), n - kc);                // dropped slots, then the old free slots
        const int m0 = __popc(pos & ((1u << (kx0 < 0 ? 0 : kx0)) - 1u));
        const int m1 = __popc(pos & ((1u << (kx1 < 0 ? 0 : kx1)) - 1u)) + 1;
        if (kx0 >= 0) K = perm_insert(K, m0, pool & 7u);
        if (kx1 >= 0) K = perm_insert(K, m1, (pool >> 4) & 7u);
        const int m = kc + nx;
        const u32 np = K | shl4(shr4(pool, nx), m);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if ((q == 0 ? kx0 : kx1) >= 0) {
                const u32 sl = (q == 0 ? pool : (pool >> 4)) & 7u;
                float qf[D];
                float w = 0.0f;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    PV(sl, c) = I[q][c];
                    qf[c] = (float)(I[q][c] - pi[c]);
                    w = fmaf(qf[c], qf[c], w);
                }
                if (WEIGHTED) PV(sl, D) = I[q][D];
                if (D == 3) QV(sl, 0) = make_float4(qf[0], qf[1], qf[2], w);
                else { QV(sl, 0) = make_float4(qf[0], qf[1], qf[2], qf[3]); QV(sl, 1) = make_float4(qf[D - 2], qf[D - 1], w, 0.0f); }
                r2_new = fmaxf(r2_new, w);
            }
        }
        perm = np;
        n = m;
        // radius of the new polygon: between the new vertices (it contains them) and the old radius (it shrank)
        R2lo = r2_new; r2_valid = false;
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
            if (last_jj < (int)nn - 1) {
                const float dl = __ldg(prow32 + (size_t)(nn - 1) * PS32 + D);
                sr_ok = radius_test(dl, nn - 1, R2f, r2_valid ? R2f : R2lo);
            }
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
                {
                    const u32 s0 = SLOT(0), s1 = SLOT(1);
#pragma unroll
                    for (int c = 0; c < VW; ++c) { p1[c] = PV(s0, c); p3[c] = PV(s1, c); }
                }
                double ea = sqrt(dist2<D>(p1, p3));
                for (int i = 1; i + 1 < n; ++i) {
                    const u32 sl = SLOT(i + 1);
#pragma unroll
                    for (int c = 0; c < VW; ++c) { p2[c] = p3[c]; p3[c] = PV(sl, c); }
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
#undef QV
#undef SLOT
}

// Pairs are taken in windows of CLIPW_W consecutive entries of the seed-major pair list (= a few dozen
// Morton-neighbouring seeds and the facets around them: facet corners, bisector rows and the per-pair outputs of a
// window stay in L1/L2). Inside a window the pairs are counting-sorted by class (number of bisectors that may cut
// them) in shared memory, so that the lanes of a warp run the same number of plane iterations; each thread then
// clips CLIPW_ROUNDS pairs, one from each quarter of the sorted window (equal work per warp).
#ifndef CLIPW_THREADS
#define CLIPW_THREADS 128
#endif
#ifndef CLIPW_ROUNDS
#define CLIPW_ROUNDS 4          // measured at C2: 4 rounds 0.429 ms, 6 rounds 0.423 ms, 8 rounds 0.461 ms (shared memory)
#endif
#define CLIPW_W (CLIPW_THREADS * CLIPW_ROUNDS)
#define CLIPW_NW (CLIPW_THREADS / 32)
#define CLIPW_NCNT (8 * CLIPW_ROUNDS * CLIPW_NW)

#ifndef CLIPW_MINBLK
#define CLIPW_MINBLK 5
#endif
template <int D, bool WEIGHTED>
__global__ void __launch_bounds__(CLIPW_THREADS, (D == 3 && !WEIGHTED) ? CLIPW_MINBLK : 1)
clip_win_kernel(ClipFlatArgs a) {
    constexpr int VW = D + (WEIGHTED ? 1 : 0);
    extern __shared__ double s_dyn[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int QW = CLIPF_QW(D);
    double* P = s_dyn + (size_t)w * (CLIPF_CAP * VW * 32) + lane;
    float4* Q = (float4*)(s_dyn + (size_t)CLIPW_NW * CLIPF_CAP * VW * 32) + (size_t)w * (CLIPF_CAP * QW * 32) + lane;
    unsigned short* order = (unsigned short*)((float4*)(s_dyn + (size_t)CLIPW_NW * CLIPF_CAP * VW * 32) +
                                              (size_t)CLIPW_NW * CLIPF_CAP * QW * 32);                  // [CLIPW_W]
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
        // Round r serves the r-th quarter of the sorted window; the warps rotate through the four 32-pair groups of a
        // quarter (warp w takes group (w + r) mod 4), so that every warp gets the same mix of cheap and expensive
        // groups and the barrier at the end of the window finds them together. The (pair, seed, facet, mask) record of
        // the next round is requested while the current pair is clipped.
        auto sorted_pos = [&](int r) { return (u32)(r * CLIPW_THREADS + (((w + r) & (CLIPW_NW - 1)) << 5) + lane); };
        u32 tn = 0, sn = 0, fn = 0, mn = 0;
        bool vn = sorted_pos(0) < cnt;
        if (vn) { tn = base + order[sorted_pos(0)]; sn = a.flat_seed[tn]; fn = a.flat_facet[tn]; mn = a.flat_mask[tn]; }
#pragma unroll 1
        for (int r = 0; r < CLIPW_ROUNDS; ++r) {
            const u32 t = tn, sd = sn, ft = fn, mk = mn;
            const bool v = vn;
            vn = (r + 1 < CLIPW_ROUNDS) && sorted_pos(r + 1) < cnt;
            if (vn) { tn = base + order[sorted_pos(r + 1)]; sn = a.flat_seed[tn]; fn = a.flat_facet[tn]; mn = a.flat_mask[tn]; }
            if (v) clip_one_pair<D, WEIGHTED, CLIPF_CAP>(a, t, sd, ft, mk, P, Q, st);
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

// Eight lanes per seed, four seeds per warp (the loads of four seeds are in flight together: the kernel is latency-bound).
// The sum is the one of the former warp-per-seed layout, bit for bit: position q = 0..31 takes the pairs q, q + 32, ... in
// facet order, then the xor tree 16, 8, 4, 2, 1; lane g of a group holds the positions g, g + 8, g + 16, g + 24, so the first
// two levels of the tree are local.
template <int D>
__global__ void __launch_bounds__(256)
reduce_pairs_kernel(ReduceArgs a) {
    const int lane = threadIdx.x & 31, g = lane & 7;
    const u32 i = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4u + (u32)(lane >> 3);
    const bool valid = i < a.nown;
    const u32 s = a.qbegin + i;
    u32 b = 0, e = 0;
    if (valid) { b = a.pair_off[i]; e = a.pair_off[i + 1]; }
    double acc[4][D + 1];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c <= D; ++c) acc[j][c] = 0.0;
    u32 ps = 0;
    for (u32 base = b; base < e; base += 32) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const u32 t = base + (u32)g + 8u * j;
            if (t < e) {
                acc[j][0] += a.contrib[t];
#pragma unroll
                for (int c = 0; c < D; ++c) acc[j][c + 1] += a.contrib[(size_t)(c + 1) * a.cstride + t];
                ps |= a.pstat[t];
            }
        }
    }
    double acc_s, acc_v[D];
    {
        double r[D + 1];
#pragma unroll
        for (int c = 0; c <= D; ++c) {
            const double lo = acc[0][c] + acc[2][c];      // level 16: q and q + 16
            const double hi = acc[1][c] + acc[3][c];
            double v = lo + hi;                           // level 8
            v += __shfl_xor_sync(B200_FULL, v, 4);
            v += __shfl_xor_sync(B200_FULL, v, 2);
            v += __shfl_xor_sync(B200_FULL, v, 1);
            r[c] = v;
        }
        acc_s = r[0];
#pragma unroll
        for (int c = 0; c < D; ++c) acc_v[c] = r[c + 1];
    }
#pragma unroll
    for (int m = 4; m > 0; m >>= 1) ps |= __shfl_xor_sync(B200_FULL, ps, m);
    if (g != 0 || !valid) return;
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
