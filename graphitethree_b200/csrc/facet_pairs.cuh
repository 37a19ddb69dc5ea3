// facet_pairs.cuh — candidate (facet, seed) pairs AND their classification, one thread per facet.
//
// Replaces the facet-driven double flood-fill of
//   GEOGen::RestrictedVoronoiDiagram::compute_surfacic_with_seeds_priority (generic_RVD.h:1318-1424)
// and its 1-NN start (find_seed_near_facet, generic_RVD.h:2010-2070) by a walk on the bisector
// table of the kNN graph (clip_flat.cuh: plane_table_kernel):
//
//  1. home seed s0 of the facet = a seed whose cell contains the facet centroid: start from the
//     seed found for this facet in the previous evaluation and hop to any listed neighbour that
//     is closer to the centroid (any s0 gives a correct candidate set; a near one a small set).
//  2. For a facet with corners c_i and ANY seed s0, a seed s whose Voronoi cell meets the facet
//     has |c_i - s| <= |c_i - s0| for some corner i (the half-space {x: |x-s| <= |x-s0|} meets the
//     triangle iff it contains a corner), i.e. corner i is NOT strictly on s0's side of the bisector
//     (s0, s); and then |s - s0| <= 2 max_i |c_i - s0|. So one scan of s0's bisectors in list order,
//     up to squared distance 4.1 max_i |c_i - s0|^2, with the side test of clip_by_plane_fast on the
//     three corners, yields at once (a) every candidate seed and (b) the classification of the
//     pair (facet, s0): the bisectors that may cut it (bit mask), "cell contains the facet" (no
//     bit), "facet outside the cell" (some bisector has all corners outside).
//  3. Each candidate s gets the same scan with its own bisectors (classification of (facet, s)).
//
// If s0's list ends before the distance bound, candidates beyond the list are possible: the
// facet falls back to the uniform-grid scan of every seed in the union of the corner balls.
// Pairs are appended to per-seed rows {facet, mask}; compact_pairs_kernel sorts each row.
#pragma once
#include "common.cuh"
#include "knn.cuh"
#include "clip_flat.cuh"

#define PMASK_SR_OK 0x80000000u   // the radius test passed on the unclipped facet before the list ended
#define PMASK_BITS 0x7fffffffu
// 8 resident blocks (64 registers): the walk is bound by the latency of dependent loads, measured 13 % faster than 4
#ifndef FACET_MINBLK
#define FACET_MINBLK 8
#endif

struct FacetPairArgs {
    const double* tri;        // [T][3][D] facet corners (facets Morton-sorted once per mesh)
    u32 T;
    const void* xs;
    const u32* nbr; const u32* nbr_n; u32 kstride;
    const float* planes32;    // [S][kstride][PLANE32_STRIDE] FP32 filter copy of the bisector table (n, |n|^2)
    const uint8_t* has_planes; // optional [S]: 1 if the seed's rows of nbr/planes are valid (sharded runs)
    const u32* facet_list; const u32* facet_list_n;   // optional: the facets to process (sharded runs), count on the device
    const uint2* cell_range;
    const u32* rank_of;
    u32* facet_guess;         // [T] original index of the last home seed (B200_NONE: none)
    u32 S;
    u32 qbegin, qend;         // owned sorted range
    u32* pair_cnt;            // [S] sorted order
    u32* pair_facet;          // [S][cap]
    u32* pair_mask;           // [S][cap]
    u32 cap;
    u32* max_cnt;             // device scalar: max row length seen (overflow detection)
    uint2* tasks; u32 task_cap; u32* task_n;   // candidate (seed, facet) tasks of kernel A for kernel B
    uint2* big_list; u32 big_cap; u32* big_n;  // (facet, home seed) of the facets left to facet_big_kernel
    unsigned long long* stats; // optional: [14] facets that took the grid fallback
    GridParams g;
};

// scan of seed s's bisectors against the facet corners. Returns the mask of bisectors that may cut
// (| PMASK_SR_OK); *empty = some bisector has all corners outside; *cand = bisectors with a corner
// not strictly inside (candidate neighbours); *hop = first bisector with the facet centroid outside
// (the neighbour is closer to the centroid than s), -1 if none.
// MODE 0: mask only; 1 (home): + candidates, stops at the first hop; 2 (walk): + candidates, no hop
//
// The side values only feed conservative decisions, so the scan runs in FP32 on seed-local coordinates: with
// q = c - p_s (FP64 difference, rounded once) the reference's side value 2 c.n - d (generic_RVD_polygon.h:257-297) is
// 2 q.n + |n|^2, n = p_s - p_j, and the filter table (PLANE32_STRIDE floats per bisector: n, |n|^2) holds n and |n|^2
// rounded to float. Rounding errors: <= 1e-6 (|q||n| + |n|^2) for the FP32 evaluation, plus the reference's own FP64
// rounding on global coordinates, <= 4e-15 (|c|^2 + |p_s|^2); both are covered by the margin. A corner is "inside" /
// "outside" only beyond the margin, the radius test passes only beyond a 1e-5 relative inflation, so every decision
// taken here also holds for the reference's FP64 values; the exact tests are redone by the clip kernels.
#define PLANE32_STRIDE(D) ((D) == 3 ? 4 : 8)

template <int D, int NC, int MODE>
__device__ __forceinline__ u32 classify_facet(const double (*v)[D], double vmax2, const double* pi, const float* prow, u32 nn,
                                              bool* empty, u32* cand, int* hop) {
    constexpr int PS = PLANE32_STRIDE(D);
    float q[NC][D];
    double R2d = 0.0, pi2 = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) pi2 += pi[c] * pi[c];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        double r = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const double dq = v[i][c] - pi[c];
            q[i][c] = (float)dq;
            r += dq * dq;
        }
        R2d = fmax(R2d, r);
    }
    const float R2 = __double2float_ru(R2d);
    const float R2lim = 4.1f * R2 * 1.00001f;
    const float mextra = __double2float_ru(4e-15 * (vmax2 + pi2));
    u32 mask = 0, cm = 0;
    bool emp = false;
    if (MODE == 1) *hop = -1;
    // one bisector row = PS floats, fetched as 16-byte vectors; the next row is requested before this one is used
    float4 rowbuf[PS / 4], nextbuf[PS / 4];
    if (nn > 0) {
        const float4* r4 = (const float4*)prow;
#pragma unroll
        for (int k = 0; k < PS / 4; ++k) nextbuf[k] = __ldg(r4 + k);
    }
    for (u32 jj = 0; jj < nn; ++jj) {
#pragma unroll
        for (int k = 0; k < PS / 4; ++k) rowbuf[k] = nextbuf[k];
        if (jj + 1 < nn) {
            const float4* r4 = (const float4*)(prow + (size_t)(jj + 1) * PS);
#pragma unroll
            for (int k = 0; k < PS / 4; ++k) nextbuf[k] = __ldg(r4 + k);
        }
        const float* pl = (const float*)rowbuf;
        const float dij = pl[D];
        // radius test on the unclipped facet (generic_RVD.h:2155-2174): clipping only shrinks R2, so every
        // bisector the reference tests is visited
        if (dij > R2lim) { mask |= PMASK_SR_OK; break; }
        const float margin = 2e-6f * (R2 + dij) + mextra;
        float tsum = 0.0f;
        bool all_in = true, all_out = true;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            float l = 0.0f;
#pragma unroll
            for (int c = 0; c < D; ++c) l = fmaf(q[i][c], pl[c], l);
            const float tk = fmaf(2.0f, l, dij);
            tsum += tk;
            all_in = all_in && (tk > margin);
            all_out = all_out && (tk < -margin);
        }
        if (MODE == 1 && tsum < 0.0f) { *hop = (int)jj; break; }
        if (all_in) continue;   // clearly inside: cannot touch any clipped polygon / cell
        cm |= 1u << jj;
        if (all_out) emp = true;   // clearly outside: removes the element
        else mask |= 1u << jj;
    }
    *empty = emp;
    if (MODE >= 1) *cand = cm;
    return mask;
}

template <int D>
__device__ __forceinline__ void emit_pair(const FacetPairArgs& a, u32 s, u32 f, u32 mask) {
    if (s < a.qbegin || s >= a.qend) return;
    const u32 slot = atomicAdd(&a.pair_cnt[s], 1u);
    if (slot < a.cap) {
        a.pair_facet[(size_t)s * a.cap + slot] = f;
        a.pair_mask[(size_t)s * a.cap + slot] = mask;
    } else atomicMax(a.max_cnt, slot + 1);
}

// every owned seed inside the union of the balls B(c_i, |c_i - s0|), from the uniform grid
template <int D, int NC>
__device__ __noinline__ void grid_candidates(const FacetPairArgs& a, const double (*v)[D], double vmax2, u32 s0, u32 f) {
    constexpr int PS = PLANE32_STRIDE(D);
    const SeedRec<D>* xs = (const SeedRec<D>*)a.xs;
    double r2[NC], lo3[3], hi3[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) { lo3[ax] = 1e300; hi3[ax] = -1e300; }
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        r2[i] = dist2<D>(v[i], xs[s0].p) * (1.0 + 1e-12);
        const double rr = sqrt(r2[i]) * (1.0 + 1e-12);
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) { lo3[ax] = fmin(lo3[ax], v[i][ax] - rr); hi3[ax] = fmax(hi3[ax], v[i][ax] + rr); }
    }
    int lo[3], hi[3];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) { lo[ax] = grid_coord(a.g, lo3[ax], ax); hi[ax] = grid_coord(a.g, hi3[ax], ax); }
    for (int cz = lo[2]; cz <= hi[2]; ++cz)
        for (int cy = lo[1]; cy <= hi[1]; ++cy)
            for (int cx = lo[0]; cx <= hi[0]; ++cx) {
                const uint2 rg = a.cell_range[morton_encode(a.g, cx, cy, cz)];
                const u32 sb = max(rg.x, a.qbegin), se = min(rg.y, a.qend);
                for (u32 s = sb; s < se; ++s) {
                    double ps[D];
#pragma unroll
                    for (int c = 0; c < D; ++c) ps[c] = xs[s].p[c];
                    bool in = false;
#pragma unroll
                    for (int i = 0; i < NC; ++i) in = in || dist2<D>(v[i], ps) <= r2[i];
                    if (!in) continue;
                    const u32 nns = min(min(a.nbr_n[s], a.kstride), 31u);
                    bool empty = false;
                    const u32 mask = classify_facet<D, NC, 0>(v, vmax2, ps, a.planes32 + (size_t)s * a.kstride * PS, nns, &empty, nullptr, nullptr);
                    if (!empty) emit_pair<D>(a, s, f, mask);
                }
            }
}

// Candidates of a FACET whose home list does not reach the distance bound because the facet is much larger than the seed
// spacing (crease facets of an anisotropic 6D surface span tens of cells; their corner balls hold 10^4 seeds). The facet
// is covered by pieces: a piece whose home list reaches its distance bound is handled by the one-level rule of
// facet_home_kernel (home seed of the piece by hopping, then every listed neighbour with a piece corner not strictly
// inside its bisector); any other piece is bisected along its longest edge, down to half a seed spacing, where the union
// of the piece's corner balls is scanned on the grid (a few cells). A cell that meets the facet meets a piece. Purely
// geometric, hence complete whatever the neighbour lists are. facet_home_kernel only lists such facets; here one WARP
// takes one facet: the pieces live on a per-warp stack in shared memory, every round each lane pops one piece and either
// settles it or pushes its two halves; the seeds found go into a per-warp hash set; the lanes then classify the set
// against the whole facet with each seed's own bisectors and emit every pair once.
#define BIG_WARPS 2
#define BIG_HASH 1024            // slots per warp
#define BIG_HASH_MAX 768         // distinct seeds per facet before giving up
#define BIG_STACK 160            // pieces per warp
#define BIG_MAX_ROUNDS 256

template <int D> struct BigPiece { double w[3][D]; u32 home; u32 pad; };

template <int D>
__global__ void __launch_bounds__(BIG_WARPS * 32)
facet_big_kernel(const __grid_constant__ FacetPairArgs a) {
    constexpr int PS = PLANE32_STRIDE(D);
    extern __shared__ double s_big[];
    __shared__ u32 s_tab[BIG_WARPS][BIG_HASH];
    __shared__ u32 s_cnt[BIG_WARPS][2];      // [0] distinct seeds [1] stack height
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const SeedRec<D>* xs = (const SeedRec<D>*)a.xs;
    u32* tab = s_tab[w];
    BigPiece<D>* stack = (BigPiece<D>*)s_big + (size_t)w * BIG_STACK;
    const u32 nbig = min(*a.big_n, a.big_cap);
    const double tau = a.g.h * (1.0 / 6.0);       // half a seed spacing (the grid cell is ~3 spacings wide)
    const double tau2 = tau * tau;
    for (u32 e = blockIdx.x * BIG_WARPS + w; e < nbig; e += gridDim.x * BIG_WARPS) {
        const uint2 ent = a.big_list[e];
        const u32 f = ent.x, s0 = ent.y;
        for (int i = lane; i < BIG_HASH; i += 32) tab[i] = B200_NONE;
        double v[3][D];
        const double* t = a.tri + (size_t)f * 3 * D;
        double vmax2 = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double q2 = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) { v[i][c] = t[i * D + c]; q2 += v[i][c] * v[i][c]; }
            vmax2 = fmax(vmax2, q2);
        }
        if (lane == 0) { s_cnt[w][0] = 0; s_cnt[w][1] = 0; }
        __syncwarp();
        auto collect = [&](u32 sd) -> bool {
            u32 slot = (sd * 2654435761u) >> 22;
            for (int probe = 0; probe < BIG_HASH; ++probe) {
                const u32 old = atomicCAS(&tab[slot], B200_NONE, sd);
                if (old == sd) return true;
                if (old == B200_NONE) return atomicAdd(&s_cnt[w][0], 1u) < BIG_HASH_MAX;
                slot = (slot + 1) & (BIG_HASH - 1);
            }
            return false;
        };
        bool ok = true;
        for (int round = 0; round < BIG_MAX_ROUNDS; ++round) {
            double pw[3][D];
            u32 home = s0;
            bool active;
            if (round == 0) {
                // first round: 32 pieces = five levels of longest-edge bisection, lane bit b picks the half at level b
                active = true;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int c = 0; c < D; ++c) pw[i][c] = v[i][c];
                for (int level = 0; level < 5; ++level) {
                    double e2[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) e2[i] = dist2<D>(pw[i], pw[(i + 1) % 3]);
                    int le = 0;
                    if (e2[1] > e2[le]) le = 1;
                    if (e2[2] > e2[le]) le = 2;
                    const int i0 = le, i1 = (le + 1) % 3;
                    const int repl = ((lane >> level) & 1) ? i0 : i1;      // the corner replaced by the midpoint
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        const double mid = 0.5 * (pw[i0][c] + pw[i1][c]);
#pragma unroll
                        for (int i = 0; i < 3; ++i) if (i == repl) pw[i][c] = mid;
                    }
                }
            } else {
                const u32 height = s_cnt[w][1];
                if (height == 0) break;
                // the counter runs past the stack when a split found it full (the pieces were not stored): give up on the
                // bisection before anything is read from beyond the stack; the facet takes the whole-facet grid scan below
                if (height > BIG_STACK) { ok = false; break; }
                const u32 take = min(height, 32u);
                active = (u32)lane < take;
                if (active) {
                    const BigPiece<D>& pc = stack[height - 1 - lane];
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int c = 0; c < D; ++c) pw[i][c] = pc.w[i][c];
                    home = pc.home;
                }
                __syncwarp();
                if (lane == 0) s_cnt[w][1] = height - take;
                __syncwarp();
            }
            if (!__all_sync(B200_FULL, ok)) break;
            if (active) {
                double wmax2 = 0.0;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    double q2 = 0.0;
#pragma unroll
                    for (int c = 0; c < D; ++c) q2 += pw[i][c] * pw[i][c];
                    wmax2 = fmax(wmax2, q2);
                }
                // home seed of the piece, starting from its parent's
                bool certified = false, empty0 = false;
                u32 cand = 0;
                for (int it = 0; it < 64; ++it) {
                    if (a.has_planes && !a.has_planes[home]) break;
                    double p0[D];
#pragma unroll
                    for (int c = 0; c < D; ++c) p0[c] = xs[home].p[c];
                    const u32 nn0 = min(min(a.nbr_n[home], a.kstride), 31u);
                    int hop = -1;
                    const u32 m0 = classify_facet<D, 3, 1>(pw, wmax2, p0, a.planes32 + (size_t)home * a.kstride * PS, nn0, &empty0, &cand, &hop);
                    if (hop >= 0) { home = a.nbr[(size_t)home * a.kstride + hop]; cand = 0; continue; }
                    certified = (m0 & PMASK_SR_OK) || (nn0 + 1 >= a.S);
                    break;
                }
                bool split = false;
                if (!certified) {
                    double e2[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) e2[i] = dist2<D>(pw[i], pw[(i + 1) % 3]);
                    int le = 0;
                    if (e2[1] > e2[le]) le = 1;
                    if (e2[2] > e2[le]) le = 2;
                    if (e2[le] > tau2) {
                        split = true;
                        const u32 pos = atomicAdd(&s_cnt[w][1], 2u);
                        if (pos + 2 > BIG_STACK) ok = false;
                        else {
                            // (mid, w[le+1], w[le+2]) and (w[le], mid, w[le+2])
                            const int i0 = le, i1 = (le + 1) % 3, i2 = (le + 2) % 3;
                            BigPiece<D>& c0 = stack[pos];
                            BigPiece<D>& c1 = stack[pos + 1];
#pragma unroll
                            for (int c = 0; c < D; ++c) {
                                const double mid = 0.5 * (pw[i0][c] + pw[i1][c]);
                                c0.w[0][c] = mid; c0.w[1][c] = pw[i1][c]; c0.w[2][c] = pw[i2][c];
                                c1.w[0][c] = pw[i0][c]; c1.w[1][c] = mid; c1.w[2][c] = pw[i2][c];
                            }
                            c0.home = home; c1.home = home;
                        }
                    }
                }
                if (!split) {
                    if (a.stats) { atomicAdd(&a.stats[9], 1ull); if (!certified) atomicAdd(&a.stats[10], 1ull); }
                    if (certified) {
                        ok = ok && collect(home);
                        while (cand && ok) {
                            const int jj = __ffs(cand) - 1;
                            cand &= cand - 1;
                            ok = collect(a.nbr[(size_t)home * a.kstride + jj]);
                        }
                    } else {
                        // exact nearest seed of the piece centroid, then the union of the corner balls B(w_i, |w_i - home|)
                        double gc[D];
#pragma unroll
                        for (int c = 0; c < D; ++c) gc[c] = (pw[0][c] + pw[1][c] + pw[2][c]) * (1.0 / 3.0);
                        home = grid_nearest<D>(xs, a.cell_range, a.g, gc, nullptr);
                        double r2[3], lo3[3], hi3[3];
#pragma unroll
                        for (int ax = 0; ax < 3; ++ax) { lo3[ax] = 1e300; hi3[ax] = -1e300; }
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            r2[i] = dist2<D>(pw[i], xs[home].p) * (1.0 + 1e-12);
                            const double rr = sqrt(r2[i]) * (1.0 + 1e-12);
#pragma unroll
                            for (int ax = 0; ax < 3; ++ax) { lo3[ax] = fmin(lo3[ax], pw[i][ax] - rr); hi3[ax] = fmax(hi3[ax], pw[i][ax] + rr); }
                        }
                        int lo[3], hi[3];
#pragma unroll
                        for (int ax = 0; ax < 3; ++ax) { lo[ax] = grid_coord(a.g, lo3[ax], ax); hi[ax] = grid_coord(a.g, hi3[ax], ax); }
                        for (int cz = lo[2]; cz <= hi[2] && ok; ++cz)
                            for (int cy = lo[1]; cy <= hi[1] && ok; ++cy)
                                for (int cx = lo[0]; cx <= hi[0] && ok; ++cx) {
                                    const uint2 rg = a.cell_range[morton_encode(a.g, cx, cy, cz)];
                                    for (u32 sd = rg.x; sd < rg.y && ok; ++sd) {
                                        double ps[D];
#pragma unroll
                                        for (int c = 0; c < D; ++c) ps[c] = xs[sd].p[c];
                                        bool in = false;
#pragma unroll
                                        for (int i = 0; i < 3; ++i) in = in || dist2<D>(pw[i], ps) <= r2[i];
                                        if (in) ok = collect(sd);
                                    }
                                }
                    }
                }
            }
            __syncwarp();
        }
        const bool all_ok = __all_sync(B200_FULL, ok) && s_cnt[w][1] == 0;
        __syncwarp();
        if (a.stats && lane == 0) atomicAdd(&a.stats[15], all_ok ? 1ull : (1ull << 32));
        if (all_ok) {
            // classification of (facet, seed) for every collected seed, with the seed's own bisectors: the set is first
            // compacted into the (now empty) stack area so that every lane has a seed
            u32* dense = (u32*)stack;
            u32 nd = 0;
            for (int i0 = 0; i0 < BIG_HASH; i0 += 32) {
                const u32 sd = tab[i0 + lane];
                const bool keep = sd != B200_NONE && sd >= a.qbegin && sd < a.qend && !(a.has_planes && !a.has_planes[sd]);
                const u32 m = __ballot_sync(B200_FULL, keep);
                if (keep) dense[nd + __popc(m & ((1u << lane) - 1u))] = sd;
                nd += __popc(m);
            }
            __syncwarp();
            for (u32 i = lane; i < nd; i += 32) {
                const u32 sd = dense[i];
                double ps[D];
#pragma unroll
                for (int c = 0; c < D; ++c) ps[c] = xs[sd].p[c];
                const u32 nns = min(min(a.nbr_n[sd], a.kstride), 31u);
                bool empty = false;
                const u32 mask = classify_facet<D, 3, 0>(v, vmax2, ps, a.planes32 + (size_t)sd * a.kstride * PS, nns, &empty, nullptr, nullptr);
                if (!empty) emit_pair<D>(a, sd, f, mask);
            }
        } else if (lane == 0) {
            grid_candidates<D, 3>(a, v, vmax2, s0, f);      // budgets exceeded: corner balls of the whole facet
        }
        __syncwarp();
    }
}

// kernel A: one thread per facet — home seed, classification of (facet, home), candidate tasks
template <int D, int NC>
__global__ void __launch_bounds__(128, FACET_MINBLK)
facet_home_kernel(const __grid_constant__ FacetPairArgs a) {
    constexpr int PS = PLANE32_STRIDE(D);
    const u32 e = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const SeedRec<D>* xs = (const SeedRec<D>*)a.xs;
    u32 cand = 0, s0 = 0;
    const u32 nf = a.facet_list ? *a.facet_list_n : a.T;
    const u32 f = (e < nf) ? (a.facet_list ? a.facet_list[e] : e) : 0u;
    if (e < nf) {
        double v[NC][D];
        const double* t = a.tri + (size_t)f * NC * D;
        double vmax2 = 0.0;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            double q2 = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) { v[i][c] = t[i * D + c]; q2 += v[i][c] * v[i][c]; }
            vmax2 = fmax(vmax2, q2);
        }
        // home seed: previous answer (or a grid search), then hops while a listed neighbour is closer to the centroid
        const bool relevant = true;
        {
            double gc[D];
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double sc = 0.0;
#pragma unroll
                for (int i = 0; i < NC; ++i) sc += v[i][c];
                gc[c] = sc * (1.0 / NC);
            }
            const u32 guess = a.facet_guess[f];
            if (guess != B200_NONE) s0 = a.rank_of[guess];
            else s0 = grid_nearest<D>(xs, a.cell_range, a.g, gc, nullptr);
        }
        bool certified = false, empty0 = false;
        u32 mask0 = 0;
        bool skip = false;
        for (int it = 0; relevant && it < 32; ++it) {
            if (a.has_planes && !a.has_planes[s0]) {
                // sharded run: s0 lies more than two grid cells away from every owned seed. A seed whose cell meets the
                // facet is within 2 rho + delta of the centroid g (rho = facet radius about g, delta = |g - s0|), hence
                // within 2 rho + 2 delta of s0: if that is at most two cells, no owned seed can meet this facet
                double gc[D], rho2 = 0.0;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    double sc = 0.0;
#pragma unroll
                    for (int i = 0; i < NC; ++i) sc += v[i][c];
                    gc[c] = sc * (1.0 / NC);
                }
#pragma unroll
                for (int i = 0; i < NC; ++i) rho2 = fmax(rho2, dist2<D>(gc, v[i]));
                const double delta = sqrt(dist2<D>(gc, xs[s0].p));
                // (D > 3: the grid only sees the first three coordinates, the bound does not carry over: no skip)
                if (D == 3 && (2.0 * sqrt(rho2) + 2.0 * delta) * (1.0 + 1e-9) <= 2.0 * a.g.h) {
                    skip = true;
                    a.facet_guess[f] = (u32)xs[s0].orig;      // a seed near the centroid: lets the facet filter drop this facet next time
                    if (a.stats) atomicAdd(&a.stats[13], 1ull);
                }
                break;
            }
            double p0[D];
#pragma unroll
            for (int c = 0; c < D; ++c) p0[c] = xs[s0].p[c];
            const u32 nn0 = min(min(a.nbr_n[s0], a.kstride), 31u);
            int hop = -1;
            mask0 = classify_facet<D, NC, 1>(v, vmax2, p0, a.planes32 + (size_t)s0 * a.kstride * PS, nn0, &empty0, &cand, &hop);
            if (hop >= 0) { s0 = a.nbr[(size_t)s0 * a.kstride + hop]; cand = 0; continue; }
            // the scan reached the distance bound (or the list holds every other seed): s0 is the nearest seed of
            // the centroid and every candidate is in the list
            certified = (mask0 & PMASK_SR_OK) || (nn0 + 1 >= a.S);
            break;
        }
        if (!relevant || skip) {
            cand = 0;
        } else if (certified) {
            a.facet_guess[f] = (u32)xs[s0].orig;
            if (!empty0) emit_pair<D>(a, s0, f, mask0);
        } else {
            // s0's list ends inside the distance bound (or s0 has no table): exact nearest seed from the grid,
            // candidates from the grid
            cand = 0;
            double gc[D];
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double sc = 0.0;
#pragma unroll
                for (int i = 0; i < NC; ++i) sc += v[i][c];
                gc[c] = sc * (1.0 / NC);
            }
            s0 = grid_nearest<D>(xs, a.cell_range, a.g, gc, nullptr);
            a.facet_guess[f] = (u32)xs[s0].orig;
            if (a.stats) atomicAdd(&a.stats[14], 1ull);
            double vv[NC][D];
#pragma unroll
            for (int i = 0; i < NC; ++i)
#pragma unroll
                for (int c = 0; c < D; ++c) vv[i][c] = v[i][c];
            // facets much larger than the seed spacing go to facet_big_kernel (the corner balls of the whole facet would
            // hold thousands of seeds); everything else, and tets, scan the grid here
            bool done = false;
            if (NC == 3 && a.big_list) {
                double dm2 = 0.0;
#pragma unroll
                for (int i = 0; i < 3; ++i) dm2 = fmax(dm2, dist2<D>(vv[i], vv[(i + 1) % 3]));
                if (dm2 > a.g.h * a.g.h) {
                    const u32 pos = atomicAdd(a.big_n, 1u);
                    if (pos < a.big_cap) { a.big_list[pos] = make_uint2(f, s0); done = true; }
                }
            }
            if (!done) grid_candidates<D, NC>(a, vv, vmax2, s0, f);
        }
    }
    // candidate tasks (seed, facet), appended with one atomic per warp (one per block measured the same: 0.334 vs 0.325 ms)
    const u32 nc = __popc(cand);
    u32 incl = nc;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
        const u32 o = __shfl_up_sync(B200_FULL, incl, m);
        if (lane >= m) incl += o;
    }
    const u32 total = __shfl_sync(B200_FULL, incl, 31);
    if (total == 0) return;
    u32 base = 0;
    if (lane == 31) base = atomicAdd(a.task_n, total);
    base = __shfl_sync(B200_FULL, base, 31);
    u32 pos = base + incl - nc;
    while (cand) {
        const int jj = __ffs(cand) - 1;
        cand &= cand - 1;
        if (pos < a.task_cap) a.tasks[pos] = make_uint2(a.nbr[(size_t)s0 * a.kstride + jj], f);
        ++pos;
    }
}

// kernel B: one thread per candidate task — classification of (facet, candidate) with the candidate's bisectors
template <int D, int NC>
__global__ void __launch_bounds__(128, FACET_MINBLK)
facet_task_kernel(const __grid_constant__ FacetPairArgs a) {
    constexpr int PS = PLANE32_STRIDE(D);
    const SeedRec<D>* xs = (const SeedRec<D>*)a.xs;
    const u32 ntask = min(*a.task_n, a.task_cap);
    for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < ntask; e += gridDim.x * blockDim.x) {
        const uint2 tk = a.tasks[e];
        const u32 s = tk.x, f = tk.y;
        if (s < a.qbegin || s >= a.qend) continue;
        double v[NC][D];
        const double* t = a.tri + (size_t)f * NC * D;
        double vmax2 = 0.0;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            double q2 = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) { v[i][c] = t[i * D + c]; q2 += v[i][c] * v[i][c]; }
            vmax2 = fmax(vmax2, q2);
        }
        double ps[D];
#pragma unroll
        for (int c = 0; c < D; ++c) ps[c] = xs[s].p[c];
        const u32 nns = min(min(a.nbr_n[s], a.kstride), 31u);
        bool empty = false;
        const u32 mask = classify_facet<D, NC, 0>(v, vmax2, ps, a.planes32 + (size_t)s * a.kstride * PS, nns, &empty, nullptr, nullptr);
        if (!empty) emit_pair<D>(a, s, f, mask);
    }
}

// ---------------------------------------------------------------------------------------
// sharded runs: which facets can meet the cell of an owned seed at all
// ---------------------------------------------------------------------------------------
// once per mesh: bounding ball of every facet about its centroid, in float (radius inflated for the rounding)
template <int D, int NC>
__global__ void facet_ball_kernel(const double* tri, u32 T, float4* ball, float* rad) {
    const u32 f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= T) return;
    const double* t = tri + (size_t)f * NC * D;
    double v[NC][D], gc[D];
#pragma unroll
    for (int i = 0; i < NC; ++i)
#pragma unroll
        for (int c = 0; c < D; ++c) v[i][c] = t[i * D + c];
    double gmax = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        double sc = 0.0;
#pragma unroll
        for (int i = 0; i < NC; ++i) sc += v[i][c];
        gc[c] = sc * (1.0 / NC); gmax = fmax(gmax, fabs(gc[c]));
    }
    // the grid lives in the first three coordinates; distances in D dimensions are not smaller
    double rho2 = 0.0;
#pragma unroll
    for (int i = 0; i < NC; ++i) rho2 = fmax(rho2, dist2<D>(gc, v[i]));
    double rho = sqrt(rho2);
    rho = rho * (1.0 + 1e-6) + 4e-7 * gmax + 1e-30;
    ball[f] = make_float4((float)gc[0], (float)gc[1], (float)gc[2], __double2float_ru(rho));
    rad[f] = __double2float_ru(rho);
}

// once per grid: the grid cell of every facet centroid
template <int D, int NC>
__global__ void facet_cell_kernel(const double* tri, u32 T, GridParams g, u32* facet_cell) {
    const u32 f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= T) return;
    const double* t = tri + (size_t)f * NC * D;
    double gc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double sc = 0.0;
#pragma unroll
        for (int i = 0; i < NC; ++i) sc += t[i * D + c];
        gc[c] = sc * (1.0 / NC);
    }
    facet_cell[f] = morton_encode(g, grid_coord(g, gc[0], 0), grid_coord(g, gc[1], 1), grid_coord(g, gc[2], 2));
}

struct FacetFilterArgs {
    const float4* ball; const float* rad; const u32* facet_cell; u32 T;   // rad[f] = ball[f].w (the early exit reads 9 bytes per facet)
    const uint8_t* cellflag;      // [ncells] CELLF_* bits: an owned seed within 1 / 2 / 3 cells, cell occupied
    const uint2* cell_range;
    const u32* facet_guess; const u32* rank_of; const void* xs;
    GridParams g;
    u32* list; u32* list_n;
};

// A seed s whose cell meets facet f lies within R = 2 rho + delta of the centroid g (rho = facet radius about g,
// delta = distance from g to ANY seed). A ball of radius R <= m h about g stays within m cells of g's cell, so
// the facet is irrelevant to this rank if no owned seed lies within m cells. Level 1 uses static data only
// (a non-empty home cell bounds delta by sqrt(3) h); level 2 uses the facet's previous home seed.
// Most facets of a sharded run are far from the rank's seeds: they leave after three coalesced loads (cell id, ball,
// one status byte); only the rest follows the previous home seed.
#define FFILT_PER_THREAD 4       // facets per thread: their loads are in flight together (the kernel is latency-bound)
template <int D>
__global__ void __launch_bounds__(256)
facet_filter_kernel(FacetFilterArgs a) {
    const int lane = threadIdx.x & 31;
    const u32 base_f = blockIdx.x * (256u * FFILT_PER_THREAD) + threadIdx.x;
    u32 cid[FFILT_PER_THREAD], fl[FFILT_PER_THREAD];
    float rw[FFILT_PER_THREAD];
#pragma unroll
    for (int k = 0; k < FFILT_PER_THREAD; ++k) {
        const u32 f = base_f + 256u * k;
        cid[k] = f < a.T ? a.facet_cell[f] : 0u;
        rw[k] = f < a.T ? a.rad[f] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < FFILT_PER_THREAD; ++k) fl[k] = a.cellflag[cid[k]];
    const double h = a.g.h;
    u32 relmask = 0;
#pragma unroll
    for (int k = 0; k < FFILT_PER_THREAD; ++k) {
        const u32 f = base_f + 256u * k;
        bool rel = false;
        if (f < a.T) {
            // (D > 3: the grid only sees the first three coordinates and the float ball only stores those, so neither
            // bound on delta holds in the facet's own space: every facet stays relevant)
            if (D > 3) rel = true;
            else if (!(fl[k] & CELLF_WITHIN3) && (double)rw[k] <= 0.6 * h && (fl[k] & CELLF_OCCUPIED)) rel = false;
            else if (fl[k] & CELLF_WITHIN1) rel = true;      // an owned seed next to the facet: relevant whatever R is
            else {
                rel = true;
                const u32 guess = a.facet_guess[f];
                if (guess != B200_NONE) {
                    const float4 bb = a.ball[f];
                    const SeedRec<D>* xs = (const SeedRec<D>*)a.xs;
                    const SeedRec<D>* r = xs + a.rank_of[guess];
                    const double qx = r->p[0] - (double)bb.x, qy = r->p[1] - (double)bb.y, qz = r->p[2] - (double)bb.z;
                    const double d2 = qx * qx + qy * qy + qz * qz;
                    const double R = (2.0 * (double)bb.w + sqrt(d2)) * (1.0 + 1e-6);
                    if (R <= h) rel = (fl[k] & CELLF_WITHIN1) != 0;
                    else if (R <= 2.0 * h) rel = (fl[k] & CELLF_WITHIN2) != 0;
                    else if (R <= 3.0 * h) rel = (fl[k] & CELLF_WITHIN3) != 0;
                }
            }
        }
        relmask |= (rel ? 1u : 0u) << k;
    }
    // one append per BLOCK: half of the facets of a two-rank run are relevant, and one atomic per warp on the same counter
    // was the whole cost of this kernel
    __shared__ u32 s_cnt[8], s_base;
    const int w = threadIdx.x >> 5;
    const u32 mine = (u32)__popc(relmask);
    u32 incl = mine;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
        const u32 o = __shfl_up_sync(B200_FULL, incl, m);
        if (lane >= m) incl += o;
    }
    if (lane == 31) s_cnt[w] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 tot = 0;
        for (int i = 0; i < 8; ++i) { const u32 c = s_cnt[i]; s_cnt[i] = tot; tot += c; }
        s_base = tot ? atomicAdd(a.list_n, tot) : 0u;
    }
    __syncthreads();
    u32 pos = s_base + s_cnt[w] + incl - mine;
#pragma unroll
    for (int k = 0; k < FFILT_PER_THREAD; ++k)
        if ((relmask >> k) & 1u) a.list[pos++] = base_f + 256u * k;
}
