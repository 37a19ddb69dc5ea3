// vcell.cuh — volumetric mode, cell-first path: the Voronoi cell of a seed is built ONCE, cooperatively by a warp, and
// integrated directly when it provably lies inside the tetrahedralised domain; only the cells that may touch the domain
// boundary go through the (tetrahedron, seed) clipping of clip_tet.cuh.
//
// What the reference computes for a seed i in volumetric mode (generic_RVD.h:1464-1597, RVD.cpp:428-527, 791-910) is a
// sum over the tets t that meet its cell of integrals over  t ∩ cell(i):  the mass and first moment (Lloyd), or, from the
// faces of t ∩ cell(i) that are NOT shared with another tet, pyramids with apex p_i (CVT energy and gradient). The tets
// partition the domain, the faces between two tets are skipped by the reference itself (visit_inner_tets = false) or cancel
// (volumes add), so for a cell that does not reach the domain boundary the sum is the integral over the cell and the
// tets are irrelevant: ~35 (tet, seed) clips of 8-20 planes each collapse into one cell of ~15 faces. The bisectors are
// applied in the reference's order (increasing distance) with the reference's security-radius rule (4.1 R^2,
// generic_RVD.h:2295-2328), now on the whole cell, every new vertex is the reference's interpolation on a cell edge
// (Vertex::intersect_geom, generic_RVD_vertex.h:976-1030), the conflict zone is the reference's flood fill from the
// furthest vertex (ConvexCell::clip_by_plane, generic_RVD_cell.h:886-1160). Sums differ from the reference's by the order
// of the additions only (1e-14 relative; the parity tolerance is 1e-9).
//
// Cell = dual form, as the reference: a cell vertex is a triangle of three plane ids with three adjacent cell vertices and a
// point; 64 slots per warp in shared memory. A clip is data-parallel over the cell: every lane tests its vertices (two
// slots per lane), ballots give the conflict zone, every zone edge that leads to a kept vertex creates one new vertex in a
// free slot (prefix sum over the lanes), and the ring of new vertices is closed through a table indexed by plane id
// (a new vertex (P, v1, v2) is followed by the one whose v1 is its v2). No sequential walk anywhere.
//
// "Provably inside": a static grid over the mesh (vgrid_mark_kernel, per mesh) flags the grid cells a boundary face may
// touch and the grid cells whose centre lies in a tet; a Voronoi cell all of whose overlapped grid cells are inside and
// untouched is inside the domain. Everything else (hull cells, cells near the boundary, seeds outside the domain, cells
// that overflow the slots) is appended to a list for clip_tet_kernel.
#pragma once
#include "common.cuh"
#include "clip.cuh"

#define VC_WARPS 8
#define VC_SLOTS 64
#define VG_BND 1u      // a boundary face of the tet mesh may touch the grid cell
#define VG_IN 2u       // the centre of the grid cell lies in a tet
#define VG_MAXCHECK 1728u

struct VGrid {
    double lo[3];
    double h, inv_h;
    int res[3];
    u32* cells;        // one byte per grid cell, four per word
    u32* need;         // per evaluation: one bit per grid cell, set where a cell of the tet path may lie (need[0] bit 31 of the
                       // last word is not special; need_all[0] != 0 means "everywhere")
    u32* need_all;
};

__device__ __forceinline__ u32 vg_byte(const VGrid& g, int cx, int cy, int cz) {
    const u32 idx = ((u32)cz * (u32)g.res[1] + (u32)cy) * (u32)g.res[0] + (u32)cx;
    return (__ldg(g.cells + (idx >> 2)) >> ((idx & 3u) * 8u)) & 0xffu;
}
__device__ __forceinline__ void vg_or(const VGrid& g, int cx, int cy, int cz, u32 bit) {
    const u32 idx = ((u32)cz * (u32)g.res[1] + (u32)cy) * (u32)g.res[0] + (u32)cx;
    const u32 m = bit << ((idx & 3u) * 8u);
    if ((g.cells[idx >> 2] & m) != m) atomicOr(g.cells + (idx >> 2), m);
}

// per mesh: thread per (tet, z-slab of its bounding box)
__global__ void vgrid_mark_kernel(const double* __restrict__ tet, const uint8_t* __restrict__ inner, u32 T, VGrid g, int nslab, uint2* __restrict__ tet_gbox) {
    const u32 f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= T) return;
    const int slab = blockIdx.y;
    double p[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) p[k][c] = tet[(size_t)f * 12 + k * 3 + c];
    const u32 in = inner[f];
    // boundary faces: every grid cell their bounding box touches
    for (int lf = 0; lf < 4; ++lf) {
        if ((in >> lf) & 1u) continue;
        int c0[3], c1[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double mn = 1e300, mx = -1e300;
            for (int k = 0; k < 4; ++k) if (k != lf) { mn = fmin(mn, p[k][c]); mx = fmax(mx, p[k][c]); }
            c0[c] = max(0, (int)floor((mn - g.lo[c]) * g.inv_h - 1e-6));
            c1[c] = min(g.res[c] - 1, (int)floor((mx - g.lo[c]) * g.inv_h + 1e-6));
        }
        for (int z = c0[2] + slab; z <= c1[2]; z += nslab)
            for (int y = c0[1]; y <= c1[1]; ++y)
                for (int x = c0[0]; x <= c1[0]; ++x) vg_or(g, x, y, z, VG_BND);
    }
    // grid cell centres inside the tet (closed: a centre on a face shared by two tets is marked by both)
    double U[3], V[3], W[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { U[c] = p[1][c] - p[0][c]; V[c] = p[2][c] - p[0][c]; W[c] = p[3][c] - p[0][c]; }
    const double det = U[0] * (V[1] * W[2] - V[2] * W[1]) - U[1] * (V[0] * W[2] - V[2] * W[0]) + U[2] * (V[0] * W[1] - V[1] * W[0]);
    int c0[3], c1[3];
    u32 b0[3], b1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double mn = fmin(fmin(p[0][c], p[1][c]), fmin(p[2][c], p[3][c]));
        const double mx = fmax(fmax(p[0][c], p[1][c]), fmax(p[2][c], p[3][c]));
        c0[c] = max(0, (int)ceil((mn - g.lo[c]) * g.inv_h - 0.5 - 1e-9));
        c1[c] = min(g.res[c] - 1, (int)floor((mx - g.lo[c]) * g.inv_h - 0.5 + 1e-9));
        b0[c] = (u32)max(0, min(g.res[c] - 1, (int)floor((mn - g.lo[c]) * g.inv_h - 1e-6)));
        b1[c] = (u32)max(0, min(g.res[c] - 1, (int)floor((mx - g.lo[c]) * g.inv_h + 1e-6)));
    }
    // the grid cells the tet's bounding box touches (8 bits per bound: the grid has at most 250 cells per axis)
    if (slab == 0) tet_gbox[f] = make_uint2(b0[0] | (b0[1] << 8) | (b0[2] << 16), b1[0] | (b1[1] << 8) | (b1[2] << 16));
    if (det == 0.0) return;
    const double tol = 1e-9 * fabs(det);
    const double sg = det > 0.0 ? 1.0 : -1.0;
    for (int z = c0[2] + slab; z <= c1[2]; z += nslab)
        for (int y = c0[1]; y <= c1[1]; ++y)
            for (int x = c0[0]; x <= c1[0]; ++x) {
                double q[3] = {g.lo[0] + (x + 0.5) * g.h - p[0][0], g.lo[1] + (y + 0.5) * g.h - p[0][1], g.lo[2] + (z + 0.5) * g.h - p[0][2]};
                // barycentric coordinates times det (Cramer)
                const double b1 = q[0] * (V[1] * W[2] - V[2] * W[1]) - q[1] * (V[0] * W[2] - V[2] * W[0]) + q[2] * (V[0] * W[1] - V[1] * W[0]);
                const double b2 = U[0] * (q[1] * W[2] - q[2] * W[1]) - U[1] * (q[0] * W[2] - q[2] * W[0]) + U[2] * (q[0] * W[1] - q[1] * W[0]);
                const double b3 = U[0] * (V[1] * q[2] - V[2] * q[1]) - U[1] * (V[0] * q[2] - V[2] * q[0]) + U[2] * (V[0] * q[1] - V[1] * q[0]);
                const double s1 = b1 * sg, s2 = b2 * sg, s3 = b3 * sg, s0 = fabs(det) - s1 - s2 - s3;
                if (s0 >= -tol && s1 >= -tol && s2 >= -tol && s3 >= -tol) vg_or(g, x, y, z, VG_IN);
            }
}

// the tets whose bounding box touches a grid cell where a cell of the tet path may lie (need bits set by vcell_kernel)
__global__ void __launch_bounds__(256)
vtet_filter_kernel(const uint2* __restrict__ tet_gbox, const u32* __restrict__ in_list, const u32* __restrict__ in_n, u32 T, VGrid g,
                   u32* __restrict__ list, u32* __restrict__ list_n) {
    const int lane = threadIdx.x & 31;
    const u32 n = in_list ? *in_n : T;
    const u32 e = blockIdx.x * 256u + threadIdx.x;
    bool rel = false;
    u32 f = 0;
    if (e < n) {
        f = in_list ? in_list[e] : e;
        if (*g.need_all) rel = true;
        else {
            const uint2 b = tet_gbox[f];
            const u32 x0 = b.x & 255u, y0 = (b.x >> 8) & 255u, z0 = (b.x >> 16) & 255u;
            const u32 x1 = b.y & 255u, y1 = (b.y >> 8) & 255u, z1 = (b.y >> 16) & 255u;
            for (u32 z = z0; z <= z1 && !rel; ++z)
                for (u32 y = y0; y <= y1 && !rel; ++y) {
                    const u32 row = (z * (u32)g.res[1] + y) * (u32)g.res[0];
                    for (u32 x = x0; x <= x1; ++x) {
                        const u32 idx = row + x;
                        if ((__ldg(g.need + (idx >> 5)) >> (idx & 31u)) & 1u) { rel = true; break; }
                    }
                }
        }
    }
    __shared__ u32 s_cnt[8], s_base;
    const int w = threadIdx.x >> 5;
    const u32 bal = __ballot_sync(B200_FULL, rel);
    if (lane == 0) s_cnt[w] = (u32)__popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 tot = 0;
        for (int i = 0; i < 8; ++i) { const u32 c = s_cnt[i]; s_cnt[i] = tot; tot += c; }
        s_base = tot ? atomicAdd(list_n, tot) : 0u;
    }
    __syncthreads();
    if (rel) list[s_base + s_cnt[w] + (u32)__popc(bal & ((1u << lane) - 1u))] = f;
}

struct VCellArgs {
    const void* xs;
    const u32* nbr; const u32* nbr_n; u32 kstride;
    int nbr_by_slot;
    const u32* seed_list; u32 nseeds; const u32* nseeds_dev; u32 qbegin;
    int mode, check_SR;
    u32 S;
    double box_lo[3], box_hi[3];
    VGrid vg;
    double* out_s; double* out_v; uint8_t* flags;
    u32* redo_list; u32* redo_n;     // inside the domain, neighbour list used up before the radius test passed (check_SR)
    u32* bnd_list; u32* bnd_n;       // not provably inside the domain: (tet, seed) path
    unsigned long long* stats;       // volumetric handles: [9] cells integrated here, [10] cells sent to the tet path, [12] bisectors applied
    // mode 3 (restricted Delaunay tets, PrimalTetrahedronAction generic_RVD.h:1036-1058): rows of four ORIGINAL seed indices
    uint4* tets; unsigned long long* tet_n; unsigned long long tet_cap;
};

__device__ __forceinline__ bool vc_in(u32 lo, u32 hi, u32 t) { return ((((t & 32u) ? hi : lo) >> (t & 31u)) & 1u) != 0u; }

__device__ __forceinline__ double vc_warp_max_nonneg(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const u32 hi = (u32)(b >> 32), lo = (u32)b;
    const u32 mh = __reduce_max_sync(B200_FULL, hi);
    const u32 ml = __reduce_max_sync(B200_FULL, hi == mh ? lo : 0u);
    return __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
}

struct VcState {
    u32 used_lo, used_hi, np;
    bool overflow;
};

// max squared distance of the cell's vertices to the seed
__device__ __forceinline__ double vc_radius2(const double (*P)[VC_SLOTS], const VcState& st, double pix, double piy, double piz, int lane) {
    double r = 0.0;
    if ((st.used_lo >> lane) & 1u) {
        const double dx = P[0][lane] - pix, dy = P[1][lane] - piy, dz = P[2][lane] - piz;
        double d = dx * dx; d += dy * dy; d += dz * dz;
        r = d;
    }
    if ((st.used_hi >> lane) & 1u) {
        const double dx = P[0][lane + 32] - pix, dy = P[1][lane + 32] - piy, dz = P[2][lane + 32] - piz;
        double d = dx * dx; d += dy * dy; d += dz * dz;
        r = fmax(r, d);
    }
    return vc_warp_max_nonneg(r);
}

// ConvexCell::clip_by_plane, the whole warp on one cell. Returns true if the plane cut.
// PLANE = false: the bisector of (pi, pj = (ex, ey, ez)); id = position of pj in the seed's neighbour list; the plane gets the next id.
// PLANE = true : the half-space ex x + ey y + ez z + ew >= 0 (a face of a tetrahedron); id = the plane id it gets.
template <bool PLANE>
__device__ __forceinline__ bool vc_clip(double (*P)[VC_SLOTS], uchar4* V, uchar4* T, u32* B, unsigned short* IT, unsigned char* FS, unsigned char* PLN, VcState& st,
                                        double pix, double piy, double piz, double ex, double ey, double ez, double ew, u32 id, int lane) {
    const double pjx = ex, pjy = ey, pjz = ez;
    const bool uA = (st.used_lo >> lane) & 1u, uB = (st.used_hi >> lane) & 1u;
    double ax = 0.0, ay = 0.0, az = 0.0, bx = 0.0, by = 0.0, bz = 0.0, rA = 0.0, rB = 0.0;
    if (uA) {
        ax = P[0][lane]; ay = P[1][lane]; az = P[2][lane];
        if (PLANE) { rA += ax * ex; rA += ay * ey; rA += az * ez; rA += ew; }
        else {
            rA += (pjx - ax) * (pjx - ax); rA -= (pix - ax) * (pix - ax);
            rA += (pjy - ay) * (pjy - ay); rA -= (piy - ay) * (piy - ay);
            rA += (pjz - az) * (pjz - az); rA -= (piz - az) * (piz - az);
        }
    }
    if (uB) {
        bx = P[0][lane + 32]; by = P[1][lane + 32]; bz = P[2][lane + 32];
        if (PLANE) { rB += bx * ex; rB += by * ey; rB += bz * ez; rB += ew; }
        else {
            rB += (pjx - bx) * (pjx - bx); rB -= (pix - bx) * (pix - bx);
            rB += (pjy - by) * (pjy - by); rB -= (piy - by) * (piy - by);
            rB += (pjz - bz) * (pjz - bz); rB -= (piz - bz) * (piz - bz);
        }
    }
    const bool cA = uA && rA < 0.0, cB = uB && rB < 0.0;
    const u32 c0lo = __ballot_sync(B200_FULL, cA);
    const u32 c0hi = st.used_hi ? __ballot_sync(B200_FULL, cB) : 0u;
    if ((c0lo | c0hi) == 0u) return false;
    if (!PLANE && st.np >= 250u) { st.overflow = true; return false; }     // ids 250 .. 253 are the faces of a tetrahedron
    const uchar4 tA = uA ? T[lane] : make_uchar4(0, 0, 0, 0);
    const uchar4 tB = uB ? T[lane + 32] : make_uchar4(0, 0, 0, 0);
    // conflict zone = the connected part of the negative vertices that holds the furthest one (flood fill)
    u32 klo = c0lo, khi = c0hi;
    if (__popc(c0lo) + __popc(c0hi) > 1) {
        if (c0lo) { klo = c0lo & (0u - c0lo); khi = 0u; } else { klo = 0u; khi = c0hi & (0u - c0hi); }
        for (int pass = 0; pass < 2; ++pass) {
            for (;;) {
                const bool jA = cA && (((klo >> lane) & 1u) || vc_in(klo, khi, tA.x) || vc_in(klo, khi, tA.y) || vc_in(klo, khi, tA.z));
                const bool jB = cB && (((khi >> lane) & 1u) || vc_in(klo, khi, tB.x) || vc_in(klo, khi, tB.y) || vc_in(klo, khi, tB.z));
                const u32 nlo = __ballot_sync(B200_FULL, jA);
                const u32 nhi = c0hi ? __ballot_sync(B200_FULL, jB) : 0u;
                if (nlo == klo && nhi == khi) break;
                klo = nlo; khi = nhi;
            }
            if ((klo == c0lo && khi == c0hi) || pass == 1) break;
            // the negative vertices are not connected (rounding): restart from the furthest one, as the reference does
            double best = cA ? rA : 0.0; u32 bslot = cA ? (u32)lane : 0xffu;
            if (cB && (!(cA) || rB < best)) { best = rB; bslot = (u32)lane + 32u; }
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                const double ob = __shfl_xor_sync(B200_FULL, best, m);
                const u32 os = __shfl_xor_sync(B200_FULL, bslot, m);
                if (os != 0xffu && (bslot == 0xffu || ob < best || (ob == best && os < bslot))) { best = ob; bslot = os; }
            }
            klo = bslot < 32u ? (1u << bslot) : 0u;
            khi = bslot < 32u ? 0u : (1u << (bslot - 32u));
        }
    }
    const u32 new_v = PLANE ? id : st.np++;
    if (!PLANE && lane == 0) PLN[new_v] = (unsigned char)id;      // which neighbour of the list the plane is the bisector of
    const bool zA = (klo >> lane) & 1u, zB = (khi >> lane) & 1u;
    u32 eA = 0u, eB = 0u;       // bit e: edge e of the zone vertex leads to a kept vertex
    if (zA) eA = (vc_in(klo, khi, tA.x) ? 0u : 1u) | (vc_in(klo, khi, tA.y) ? 0u : 2u) | (vc_in(klo, khi, tA.z) ? 0u : 4u);
    if (zB) eB = (vc_in(klo, khi, tB.x) ? 0u : 1u) | (vc_in(klo, khi, tB.y) ? 0u : 2u) | (vc_in(klo, khi, tB.z) ? 0u : 4u);
    const u32 cnt = __popc(eA) + __popc(eB);
    const u32 lt = (1u << lane) - 1u;
    const u32 q0 = __ballot_sync(B200_FULL, cnt & 1u), q1 = __ballot_sync(B200_FULL, cnt & 2u), q2 = __ballot_sync(B200_FULL, cnt & 4u);
    u32 k = __popc(q0 & lt) + 2u * __popc(q1 & lt) + 4u * __popc(q2 & lt);
    const u32 total = __popc(q0) + 2u * __popc(q1) + 4u * __popc(q2);
    const u32 free_lo = ~st.used_lo, free_hi = ~st.used_hi;
    const u32 nfl = __popc(free_lo);
    if (total > 32u || total > nfl + __popc(free_hi)) { st.overflow = true; return true; }
    if (total == 0u) { st.used_lo = 0u; st.used_hi = 0u; return true; }     // everything removed
    // The zone edges are found by the few lanes of the zone; the work per edge (a new vertex) is spread over the warp:
    // the zone lanes publish (slot, edge) items, every lane publishes the free slot of its rank, lane i builds item i.
    if (eA | eB) {
#pragma unroll
        for (int e = 0; e < 3; ++e) if ((eA >> e) & 1u) IT[k++] = (unsigned short)((u32)lane | ((u32)e << 8));
#pragma unroll
        for (int e = 0; e < 3; ++e) if ((eB >> e) & 1u) IT[k++] = (unsigned short)(((u32)lane + 32u) | ((u32)e << 8));
    }
    const bool fA = (free_lo >> lane) & 1u;
    const u32 rkA = __popc(free_lo & lt);
    if (fA) FS[rkA] = (unsigned char)lane;
    const bool takeA = fA && rkA < total;
    bool takeB = false;
    if (total > nfl) {
        const bool fB = (free_hi >> lane) & 1u;
        const u32 rkB = nfl + __popc(free_hi & lt);
        if (fB) FS[rkB] = (unsigned char)(lane + 32);
        takeB = fB && rkB < total;
    }
    __syncwarp();
    // phase A: lane i reads what new vertex i needs (nothing is written yet)
    u32 N = 0u, v1 = 0u, v2 = 0u, nb = 0u, idx = 0u;
    double nx_ = 0.0, ny_ = 0.0, nz_ = 0.0;
    const bool mine = (u32)lane < total;
    if (mine) {
        // bisector (ConvexCell::clip_by_plane / intersect_geom): n = pi - pj, d = -(n . (pi + pj)) / 2
        double nx = ex, ny = ey, nz = ez, dd = ew;
        if (!PLANE) {
            nx = pix - pjx; ny = piy - pjy; nz = piz - pjz;
            dd = 0.0;
            dd -= nx * (pjx + pix); dd -= ny * (pjy + piy); dd -= nz * (pjz + piz);
            dd = 0.5 * dd;
        }
        const u32 it = IT[lane];
        const u32 t = it & 0xffu, e = it >> 8;
        N = FS[lane];
        const uchar4 tt = T[t], vt = V[t];
        nb = e == 0 ? tt.x : (e == 1 ? tt.y : tt.z);
        v1 = e == 0 ? vt.y : (e == 1 ? vt.z : vt.x);
        v2 = e == 0 ? vt.z : (e == 1 ? vt.x : vt.y);
        const double tx = P[0][t], ty = P[1][t], tz = P[2][t];
        const double kx = P[0][nb], ky = P[1][nb], kz = P[2][nb];
        double l1 = 0.0, l2 = 0.0;
        l1 += kx * nx; l1 += ky * ny; l1 += kz * nz;
        l2 += tx * nx; l2 += ty * ny; l2 += tz * nz;
        l1 = fabs(l1 + dd); l2 = fabs(l2 + dd);
        const double l12 = l1 + l2;
        if (l12 > 1e-30) { l1 /= l12; l2 /= l12; } else { l1 = 0.5; l2 = 0.5; }
        nx_ = l1 * tx + l2 * kx; ny_ = l1 * ty + l2 * ky; nz_ = l1 * tz + l2 * kz;
        // the kept vertex will see the new one where it saw the zone vertex
        const uchar4 tn = T[nb];
        idx = (tn.y == t ? 1u : 0u) | (tn.z == t ? 2u : 0u);
    }
    __syncwarp();
    // phase B: the new vertices (free slots), the links of the kept vertices (one byte each, all distinct), the ring table
    uint8_t* Tb = (uint8_t*)T;
    if (mine) {
        P[0][N] = nx_; P[1][N] = ny_; P[2][N] = nz_;
        V[N] = make_uchar4((unsigned char)new_v, (unsigned char)v1, (unsigned char)v2, 0);
        Tb[N * 4 + 0] = (unsigned char)nb;
        Tb[nb * 4 + idx] = (unsigned char)N;
        B[v1] = N;
    }
    __syncwarp();
    // phase C: close the ring: the new vertex (P, v1, v2) is followed by the new vertex whose v1 is v2
    u32 nxn = 0u;
    if (mine) { nxn = B[v2]; Tb[N * 4 + 1] = (unsigned char)nxn; }
    __syncwarp();
    if (mine) Tb[nxn * 4 + 2] = (unsigned char)N;
    const u32 add_lo = __ballot_sync(B200_FULL, takeA);
    const u32 add_hi = total > nfl ? __ballot_sync(B200_FULL, takeB) : 0u;
    st.used_lo = (st.used_lo & ~klo) | add_lo;
    st.used_hi = (st.used_hi & ~khi) | add_hi;
    __syncwarp();
    return true;
}

// Geom::tetra_volume<3> (geometry.h:483-524)
__device__ __forceinline__ double vc_tet_volume(double ax, double ay, double az, double bx, double by, double bz,
                                                double cx, double cy, double cz, double dx, double dy, double dz) {
    const double U0 = bx - ax, U1 = by - ay, U2 = bz - az;
    const double V0 = cx - ax, V1 = cy - ay, V2 = cz - az;
    const double W0 = dx - ax, W1 = dy - ay, W2 = dz - az;
    const double x = V1 * W2 - V2 * W1;
    const double y = V2 * W0 - V0 * W2;
    const double z = V0 * W1 - V1 * W0;
    return fabs((U0 * x + U1 * y + U2 * z) / 6.0);
}

// Integration of one cell (per-lane partial sums are ADDED to acc_*; the caller reduces over the warp). Every face (plane id) is
// fanned from its lowest-numbered vertex; one fan triangle per (vertex, face) corner, all corners in parallel.
// inner: bit lf set = plane 250 + lf is a face shared with another tetrahedron (skipped by the func/grad action, visit_inner_tets = false).
__device__ __forceinline__ void vc_integrate(double (*P)[VC_SLOTS], const uchar4* V, const uchar4* T, u32* B, const VcState& st,
                                             double pix, double piy, double piz, int mode, u32 inner, int lane,
                                             double& acc_s, double& acc_x, double& acc_y, double& acc_z) {
    const bool uA = (st.used_lo >> lane) & 1u, uB = (st.used_hi >> lane) & 1u;
    double ax = 0.0, ay = 0.0, az = 0.0, bx = 0.0, by = 0.0, bz = 0.0;
    uchar4 vA = make_uchar4(0, 0, 0, 0), vB = vA, tA = vA, tB = vA;
    if (uA) { ax = P[0][lane]; ay = P[1][lane]; az = P[2][lane]; vA = V[lane]; tA = T[lane]; }
    if (uB) { bx = P[0][lane + 32]; by = P[1][lane + 32]; bz = P[2][lane + 32]; vB = V[lane + 32]; tB = T[lane + 32]; }
    for (u32 i = lane; i < 256u; i += 32) B[i] = 0xffffffffu;
    __syncwarp();
    if (uA) { atomicMin(&B[vA.x], (u32)lane); atomicMin(&B[vA.y], (u32)lane); atomicMin(&B[vA.z], (u32)lane); }
    if (uB) { atomicMin(&B[vB.x], (u32)lane + 32u); atomicMin(&B[vB.y], (u32)lane + 32u); atomicMin(&B[vB.z], (u32)lane + 32u); }
    __syncwarp();
    const u32 t0 = st.used_lo ? (u32)__ffs((int)st.used_lo) - 1u : 32u + (u32)__ffs((int)st.used_hi) - 1u;
    const double q0x = P[0][t0], q0y = P[1][t0], q0z = P[2][t0];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        if (!(half ? uB : uA)) continue;
        const u32 t = half ? (u32)lane + 32u : (u32)lane;
        const uchar4 vt = half ? vB : vA, tt = half ? tB : tA;
        const double px = half ? bx : ax, py = half ? by : ay, pz = half ? bz : az;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const u32 cv = i == 0 ? vt.x : (i == 1 ? vt.y : vt.z);
            const u32 nx = i == 0 ? tt.y : (i == 1 ? tt.z : tt.x);       // next vertex around the face (move_to_next_around_vertex)
            const u32 A = B[cv];
            if (t == A || nx == A) continue;
            const double Ax = P[0][A], Ay = P[1][A], Az = P[2][A];
            const double Nx = P[0][nx], Ny = P[1][nx], Nz = P[2][nx];
            if (mode == 0) {
                // ComputeCentroidsVolumetric over TetrahedronAction (generic_RVD.h:901-978, RVD.cpp:428-497): tets from the cell's first vertex
                if (A == t0) continue;
                const double m = vc_tet_volume(q0x, q0y, q0z, Ax, Ay, Az, px, py, pz, Nx, Ny, Nz);
                const double sc = m / 4.0;
                acc_s += m;
                acc_x += sc * (q0x + Ax + px + Nx); acc_y += sc * (q0y + Ay + py + Ny); acc_z += sc * (q0z + Az + pz + Nz);
            } else {
                // ComputeCVTFuncGradVolumetric (RVD.cpp:791-876): pyramid of the face triangle with apex p_i
                if (cv >= 250u && ((inner >> (cv - 250u)) & 1u)) continue;
                const double mi = vc_tet_volume(pix, piy, piz, Ax, Ay, Az, px, py, pz, Nx, Ny, Nz);
                double fi = 0.0;
                {
                    const double Uc = Ax - pix, Vc = px - pix, Wc = Nx - pix;
                    fi += Uc * Uc + Vc * Vc + Wc * Wc; fi += (Uc * Vc + Vc * Wc + Wc * Uc);
                }
                {
                    const double Uc = Ay - piy, Vc = py - piy, Wc = Ny - piy;
                    fi += Uc * Uc + Vc * Vc + Wc * Wc; fi += (Uc * Vc + Vc * Wc + Wc * Uc);
                }
                {
                    const double Uc = Az - piz, Vc = pz - piz, Wc = Nz - piz;
                    fi += Uc * Uc + Vc * Vc + Wc * Wc; fi += (Uc * Vc + Vc * Wc + Wc * Uc);
                }
                fi *= (mi / 10.0);
                acc_s += fi;
                acc_x += 2.0 * mi * (0.75 * pix - 0.25 * Ax - 0.25 * px - 0.25 * Nx);
                acc_y += 2.0 * mi * (0.75 * piy - 0.25 * Ay - 0.25 * py - 0.25 * Ny);
                acc_z += 2.0 * mi * (0.75 * piz - 0.25 * Az - 0.25 * pz - 0.25 * Nz);
            }
        }
    }
    __syncwarp();
}

#ifndef VC_MINBLK
#define VC_MINBLK 3
#endif
__global__ void __launch_bounds__(VC_WARPS * 32, VC_MINBLK) vcell_kernel(VCellArgs a) {
    __shared__ double sP[VC_WARPS][3][VC_SLOTS];
    __shared__ uchar4 sV[VC_WARPS][VC_SLOTS];
    __shared__ uchar4 sT[VC_WARPS][VC_SLOTS];
    __shared__ u32 sB[VC_WARPS][256];
    __shared__ unsigned short sIT[VC_WARPS][32];
    __shared__ unsigned char sFS[VC_WARPS][VC_SLOTS];
    __shared__ unsigned char sPLN[VC_WARPS][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double (*P)[VC_SLOTS] = sP[w];
    uchar4* V = sV[w];
    uchar4* T = sT[w];
    u32* B = sB[w];
    unsigned short* IT = sIT[w];
    unsigned char* FS = sFS[w];
    unsigned char* PLN = sPLN[w];
    const SeedRec<3>* xs = (const SeedRec<3>*)a.xs;
    const u32 nseeds = a.nseeds_dev ? *a.nseeds_dev : a.nseeds;
    unsigned long long st_cells = 0, st_bnd = 0, st_clips = 0;
    for (u32 si = blockIdx.x * VC_WARPS + w; si < nseeds; si += gridDim.x * VC_WARPS) {
        const u32 s = a.seed_list ? a.seed_list[si] : a.qbegin + si;
        const double pix = xs[s].p[0], piy = xs[s].p[1], piz = xs[s].p[2];
        const size_t nrow = a.nbr_by_slot ? (size_t)si : (size_t)s;
        const u32 nn = min(a.nbr_n[nrow], a.kstride);
        const u32* nrowp = a.nbr + nrow * a.kstride;
        __syncwarp();
        if (lane < 8) {
            const int b0 = lane & 1, b1 = (lane >> 1) & 1, b2 = (lane >> 2) & 1;
            P[0][lane] = b0 ? a.box_hi[0] : a.box_lo[0];
            P[1][lane] = b1 ? a.box_hi[1] : a.box_lo[1];
            P[2][lane] = b2 ? a.box_hi[2] : a.box_lo[2];
            uchar4 v = make_uchar4((unsigned char)b0, (unsigned char)(2 + b1), (unsigned char)(4 + b2), 0);
            uchar4 t = make_uchar4((unsigned char)(lane ^ 1), (unsigned char)(lane ^ 2), (unsigned char)(lane ^ 4), 0);
            if ((b0 + b1 + b2) & 1) {
                unsigned char x = v.y; v.y = v.z; v.z = x;
                x = t.y; t.y = t.z; t.z = x;
            }
            V[lane] = v; T[lane] = t;
        }
        __syncwarp();
        VcState st;
        st.used_lo = 0xffu; st.used_hi = 0u; st.np = 6u; st.overflow = false;
        double R2 = vc_radius2(P, st, pix, piy, piz, lane);
        bool sr_ok = false, done = false;
        for (u32 base = 0; base < nn && !done; base += 32) {
            double qx = 0.0, qy = 0.0, qz = 0.0, qd = 0.0;
            if (base + lane < nn) {
                const SeedRec<3>* r = xs + nrowp[base + lane];
                qx = r->p[0]; qy = r->p[1]; qz = r->p[2];
                const double dx = qx - pix, dy = qy - piy, dz = qz - piz;
                qd = dx * dx; qd += dy * dy; qd += dz * dz;
            }
            const u32 cnt = min(32u, nn - base);
            for (u32 l = 0; l < cnt; ++l) {
                const double dj = shfl_d(qd, (int)l);
                if (dj > 4.1 * R2) { sr_ok = true; done = true; break; }
                ++st_clips;
                const bool cut = vc_clip<false>(P, V, T, B, IT, FS, PLN, st, pix, piy, piz, shfl_d(qx, (int)l), shfl_d(qy, (int)l), shfl_d(qz, (int)l), 0.0, base + l, lane);
                if (st.overflow || (st.used_lo | st.used_hi) == 0u) { done = true; break; }
                if (cut) R2 = vc_radius2(P, st, pix, piy, piz, lane);
            }
        }
        const bool uA = (st.used_lo >> lane) & 1u, uB = (st.used_hi >> lane) & 1u;
        double ax = 0.0, ay = 0.0, az = 0.0, bx = 0.0, by = 0.0, bz = 0.0;
        if (uA) { ax = P[0][lane]; ay = P[1][lane]; az = P[2][lane]; }
        if (uB) { bx = P[0][lane + 32]; by = P[1][lane + 32]; bz = P[2][lane + 32]; }
        // inside the domain?  every grid cell the cell's bounding box overlaps must be inside and free of boundary faces
        const bool nonempty = (st.used_lo | st.used_hi) != 0u;
        bool interior = !st.overflow && nonempty;
        int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, mx[3] = {-0x7fffffff, -0x7fffffff, -0x7fffffff};
        if (nonempty) {
            const double big = 1e9;
            if (uA) {
                const double g0 = fmin(fmax((ax - a.vg.lo[0]) * a.vg.inv_h, -big), big), g1 = fmin(fmax((ay - a.vg.lo[1]) * a.vg.inv_h, -big), big),
                             g2 = fmin(fmax((az - a.vg.lo[2]) * a.vg.inv_h, -big), big);
                mn[0] = (int)floor(g0 - 1e-6); mx[0] = (int)floor(g0 + 1e-6);
                mn[1] = (int)floor(g1 - 1e-6); mx[1] = (int)floor(g1 + 1e-6);
                mn[2] = (int)floor(g2 - 1e-6); mx[2] = (int)floor(g2 + 1e-6);
            }
            if (uB) {
                const double g0 = fmin(fmax((bx - a.vg.lo[0]) * a.vg.inv_h, -big), big), g1 = fmin(fmax((by - a.vg.lo[1]) * a.vg.inv_h, -big), big),
                             g2 = fmin(fmax((bz - a.vg.lo[2]) * a.vg.inv_h, -big), big);
                mn[0] = min(mn[0], (int)floor(g0 - 1e-6)); mx[0] = max(mx[0], (int)floor(g0 + 1e-6));
                mn[1] = min(mn[1], (int)floor(g1 - 1e-6)); mx[1] = max(mx[1], (int)floor(g1 + 1e-6));
                mn[2] = min(mn[2], (int)floor(g2 - 1e-6)); mx[2] = max(mx[2], (int)floor(g2 + 1e-6));
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) { mn[c] = __reduce_min_sync(B200_FULL, mn[c]); mx[c] = __reduce_max_sync(B200_FULL, mx[c]); }
        }
        if (interior) {
            if (mn[0] < 0 || mn[1] < 0 || mn[2] < 0 || mx[0] >= a.vg.res[0] || mx[1] >= a.vg.res[1] || mx[2] >= a.vg.res[2]) interior = false;
            else {
                const u32 ex = (u32)(mx[0] - mn[0] + 1), ey = (u32)(mx[1] - mn[1] + 1), ez = (u32)(mx[2] - mn[2] + 1);
                const u32 nc = ex * ey * ez;
                if (nc > VG_MAXCHECK) interior = false;
                else {
                    bool ok = true;
                    for (u32 i = lane; i < nc; i += 32) {
                        const u32 x = i % ex, y = (i / ex) % ey, z = i / (ex * ey);
                        ok = ok && ((vg_byte(a.vg, mn[0] + (int)x, mn[1] + (int)y, mn[2] + (int)z) & (VG_BND | VG_IN)) == VG_IN);
                    }
                    interior = __all_sync(B200_FULL, ok);
                }
            }
        }
        const bool exhausted = !sr_ok && nn > 0;
        if (!interior) {
            if (lane == 0) { const u32 pos = atomicAdd(a.bnd_n, 1u); a.bnd_list[pos] = s; }
            ++st_bnd;
            // the tets this cell can meet lie in the grid cells its bounding box touches (a cell only shrinks from here on)
            if (nonempty && a.vg.need) {
                const int x0 = max(mn[0], 0), y0 = max(mn[1], 0), z0 = max(mn[2], 0);
                const int x1 = min(mx[0], a.vg.res[0] - 1), y1 = min(mx[1], a.vg.res[1] - 1), z1 = min(mx[2], a.vg.res[2] - 1);
                if (x0 <= x1 && y0 <= y1 && z0 <= z1) {
                    const u32 ex = (u32)(x1 - x0 + 1), ey = (u32)(y1 - y0 + 1), ez = (u32)(z1 - z0 + 1);
                    const u32 nc = ex * ey * ez;
                    if (nc > 32768u) { if (lane == 0) *a.vg.need_all = 1u; }
                    else
                        for (u32 i = lane; i < nc; i += 32) {
                            const u32 x = i % ex, y = (i / ex) % ey, z = i / (ex * ey);
                            const u32 idx = (((u32)z0 + z) * (u32)a.vg.res[1] + (u32)y0 + y) * (u32)a.vg.res[0] + (u32)x0 + x;
                            const u32 m = 1u << (idx & 31u);
                            if (!(a.vg.need[idx >> 5] & m)) atomicOr(a.vg.need + (idx >> 5), m);
                        }
                }
            }
            continue;
        }
        if (exhausted && a.check_SR && nn + 1 < a.S && nn < B200CVT_KMAX_DEV && a.redo_list) {
            if (lane == 0) { const u32 pos = atomicAdd(a.redo_n, 1u); a.redo_list[pos] = s; }
            continue;
        }
        ++st_cells;
        if (a.mode == 3) {
            // every vertex of the cell on three bisectors is a Voronoi vertex inside the domain: its Delaunay tet, once (smallest seed)
            __syncwarp();
            const u32 me = (u32)xs[s].orig;
            uint4 rowA = make_uint4(0, 0, 0, 0), rowB = rowA;
            bool eA = false, eB = false;
            if (uA) {
                const uchar4 v = V[lane];
                if (v.x >= 6 && v.y >= 6 && v.z >= 6) {
                    rowA = make_uint4(me, (u32)xs[nrowp[PLN[v.x]]].orig, (u32)xs[nrowp[PLN[v.y]]].orig, (u32)xs[nrowp[PLN[v.z]]].orig);
                    eA = me < rowA.y && me < rowA.z && me < rowA.w;
                }
            }
            if (uB) {
                const uchar4 v = V[lane + 32];
                if (v.x >= 6 && v.y >= 6 && v.z >= 6) {
                    rowB = make_uint4(me, (u32)xs[nrowp[PLN[v.x]]].orig, (u32)xs[nrowp[PLN[v.y]]].orig, (u32)xs[nrowp[PLN[v.z]]].orig);
                    eB = me < rowB.y && me < rowB.z && me < rowB.w;
                }
            }
            const u32 mA = __ballot_sync(B200_FULL, eA), mB = __ballot_sync(B200_FULL, eB);
            const u32 tot = (u32)__popc(mA) + (u32)__popc(mB);
            unsigned long long base_row = 0;
            if (lane == 0 && tot) base_row = atomicAdd(a.tet_n, (unsigned long long)tot);
            base_row = __shfl_sync(B200_FULL, base_row, 0);
            const u32 ltm = (1u << lane) - 1u;
            if (eA) { const unsigned long long r = base_row + __popc(mA & ltm); if (r < a.tet_cap) a.tets[r] = rowA; }
            if (eB) { const unsigned long long r = base_row + __popc(mA) + __popc(mB & ltm); if (r < a.tet_cap) a.tets[r] = rowB; }
            continue;
        }
        double acc_s = 0.0, acc_x = 0.0, acc_y = 0.0, acc_z = 0.0;
        vc_integrate(P, V, T, B, st, pix, piy, piz, a.mode, 0u, lane, acc_s, acc_x, acc_y, acc_z);
        acc_s = warp_sum(acc_s); acc_x = warp_sum(acc_x); acc_y = warp_sum(acc_y); acc_z = warp_sum(acc_z);
        if (lane == 0) {
            a.out_s[s] = acc_s;
            a.out_v[(size_t)s * 3 + 0] = acc_x; a.out_v[(size_t)s * 3 + 1] = acc_y; a.out_v[(size_t)s * 3 + 2] = acc_z;
            uint8_t f8 = (uint8_t)(a.flags[s] & ~(uint8_t)(1 | 4 | 8));
            if (exhausted) {
                if (!a.check_SR) f8 |= 1;
                else if (nn + 1 >= a.S) { }
                else if (nn >= B200CVT_KMAX_DEV) f8 |= 8;
            }
            a.flags[s] = f8;
        }
    }
    if (a.stats && lane == 0) {
        if (st_cells) atomicAdd(&a.stats[9], st_cells);
        if (st_bnd) atomicAdd(&a.stats[10], st_bnd);
        if (st_clips) atomicAdd(&a.stats[12], st_clips);
    }
}

// ---------------------------------------------------------------------------------------
// The cells that may reach the domain boundary: cell ∩ tet for every candidate tet, cooperatively.
// The (tet, seed) kernel of clip_tet.cuh starts from the tet and applies up to 20 bisectors to a 4..10-vertex piece, one piece
// per lane (7.9 of 32 lanes). Here the warp builds the seed's cell once (as vcell_kernel), then, tet by tet, clips a COPY of
// the cell by the four face planes of the tet (vc_clip<true>) and integrates the piece — the same set tet ∩ cell(i), summed over
// the candidate tets in ascending id. A piece is exact when the last neighbour of the list is farther than 2.02 R_piece (then
// no seed outside the list can cut it): otherwise the cell is flagged (Lloyd mode) or run again with a longer list (check_SR),
// exactly the two outcomes of clip_by_cell_SR (generic_RVD.h:2295-2346). Cells that overflow the slots go to clip_tet_kernel.
// ---------------------------------------------------------------------------------------
struct VCellTetArgs {
    VCellArgs c;                       // seed_list = the cells of the tet path; redo_list: list used up (check_SR); bnd_list unused
    const double* tet; const uint8_t* tet_inner;
    const u32* pair_cnt; u32* pair_facet; u32 cap;
    u32* ovf_list; u32* ovf_n;         // cells this kernel gives up on
};

__device__ __forceinline__ double vc_warp_min_d(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmin(v, __shfl_xor_sync(B200_FULL, v, m));
    return v;
}
__device__ __forceinline__ double vc_warp_max_d(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmax(v, __shfl_xor_sync(B200_FULL, v, m));
    return v;
}

#ifndef VCT_MINBLK
#define VCT_MINBLK 2
#endif
__global__ void __launch_bounds__(VC_WARPS * 32, VCT_MINBLK) vcell_tet_kernel(VCellTetArgs ta) {
    const VCellArgs& a = ta.c;
    __shared__ double sP[VC_WARPS][3][VC_SLOTS];
    __shared__ uchar4 sV[VC_WARPS][VC_SLOTS];
    __shared__ uchar4 sT[VC_WARPS][VC_SLOTS];
    __shared__ double sP2[VC_WARPS][3][VC_SLOTS];
    __shared__ uchar4 sV2[VC_WARPS][VC_SLOTS];
    __shared__ uchar4 sT2[VC_WARPS][VC_SLOTS];
    __shared__ u32 sB[VC_WARPS][256];
    __shared__ unsigned short sIT[VC_WARPS][32];
    __shared__ unsigned char sFS[VC_WARPS][VC_SLOTS];
    __shared__ unsigned char sPLN[VC_WARPS][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double (*P)[VC_SLOTS] = sP[w];
    uchar4* V = sV[w];
    uchar4* T = sT[w];
    double (*P2)[VC_SLOTS] = sP2[w];
    uchar4* V2 = sV2[w];
    uchar4* T2 = sT2[w];
    u32* B = sB[w];
    unsigned short* IT = sIT[w];
    unsigned char* FS = sFS[w];
    unsigned char* PLN = sPLN[w];
    const SeedRec<3>* xs = (const SeedRec<3>*)a.xs;
    const u32 nseeds = a.nseeds_dev ? *a.nseeds_dev : a.nseeds;
    for (u32 si = blockIdx.x * VC_WARPS + w; si < nseeds; si += gridDim.x * VC_WARPS) {
        const u32 s = a.seed_list ? a.seed_list[si] : a.qbegin + si;
        const double pix = xs[s].p[0], piy = xs[s].p[1], piz = xs[s].p[2];
        const size_t nrow = a.nbr_by_slot ? (size_t)si : (size_t)s;
        const u32 nn = min(a.nbr_n[nrow], a.kstride);
        const u32* nrowp = a.nbr + nrow * a.kstride;
        __syncwarp();
        if (lane < 8) {
            const int b0 = lane & 1, b1 = (lane >> 1) & 1, b2 = (lane >> 2) & 1;
            P[0][lane] = b0 ? a.box_hi[0] : a.box_lo[0];
            P[1][lane] = b1 ? a.box_hi[1] : a.box_lo[1];
            P[2][lane] = b2 ? a.box_hi[2] : a.box_lo[2];
            uchar4 v = make_uchar4((unsigned char)b0, (unsigned char)(2 + b1), (unsigned char)(4 + b2), 0);
            uchar4 t = make_uchar4((unsigned char)(lane ^ 1), (unsigned char)(lane ^ 2), (unsigned char)(lane ^ 4), 0);
            if ((b0 + b1 + b2) & 1) {
                unsigned char x = v.y; v.y = v.z; v.z = x;
                x = t.y; t.y = t.z; t.z = x;
            }
            V[lane] = v; T[lane] = t;
        }
        __syncwarp();
        VcState st;
        st.used_lo = 0xffu; st.used_hi = 0u; st.np = 6u; st.overflow = false;
        double R2 = vc_radius2(P, st, pix, piy, piz, lane);
        double last_d = 0.0;
        bool done = false;
        for (u32 base = 0; base < nn && !done; base += 32) {
            double qx = 0.0, qy = 0.0, qz = 0.0, qd = 0.0;
            if (base + lane < nn) {
                const SeedRec<3>* r = xs + nrowp[base + lane];
                qx = r->p[0]; qy = r->p[1]; qz = r->p[2];
                const double dx = qx - pix, dy = qy - piy, dz = qz - piz;
                qd = dx * dx; qd += dy * dy; qd += dz * dz;
            }
            const u32 cnt = min(32u, nn - base);
            if (base + cnt == nn) last_d = shfl_d(qd, (int)cnt - 1);
            for (u32 l = 0; l < cnt; ++l) {
                const double dj = shfl_d(qd, (int)l);
                if (dj > 4.1 * R2) { done = true; break; }
                const bool cut = vc_clip<false>(P, V, T, B, IT, FS, PLN, st, pix, piy, piz, shfl_d(qx, (int)l), shfl_d(qy, (int)l), shfl_d(qz, (int)l), 0.0, base + l, lane);
                if (st.overflow || (st.used_lo | st.used_hi) == 0u) { done = true; break; }
                if (cut) R2 = vc_radius2(P, st, pix, piy, piz, lane);
            }
        }
        if (nn > 0 && last_d == 0.0) {
            // the loop left before the last chunk was loaded: the distance of the last neighbour of the list
            const SeedRec<3>* r = xs + nrowp[nn - 1];
            const double dx = r->p[0] - pix, dy = r->p[1] - piy, dz = r->p[2] - piz;
            last_d = dx * dx; last_d += dy * dy; last_d += dz * dz;
        }
        if (st.overflow) {
            if (lane == 0) { const u32 pos = atomicAdd(ta.ovf_n, 1u); ta.ovf_list[pos] = s; }
            continue;
        }
        double acc_s = 0.0, acc_x = 0.0, acc_y = 0.0, acc_z = 0.0;
        bool any_exh = false, gave_up = false;
        if ((st.used_lo | st.used_hi) != 0u) {
            // bounding box of the cell: tets outside of it are skipped without a clip
            const bool uA = (st.used_lo >> lane) & 1u, uB = (st.used_hi >> lane) & 1u;
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (uA) { lo[c] = P[c][lane]; hi[c] = P[c][lane]; }
                if (uB) { lo[c] = fmin(lo[c], P[c][lane + 32]); hi[c] = fmax(hi[c], P[c][lane + 32]); }
                lo[c] = vc_warp_min_d(lo[c]); hi[c] = vc_warp_max_d(hi[c]);
            }
            // candidate tets in ascending id (rows are filled through atomics)
            const u32 npairs = min(ta.pair_cnt[s], ta.cap);
            u32* row = ta.pair_facet + (size_t)s * ta.cap;
            if (npairs > 1) {
                u32 n2 = 32; while (n2 < npairs) n2 <<= 1;
                n2 = min(n2, ta.cap);
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
            for (u32 pidx = 0; pidx < npairs && !gave_up; ++pidx) {
                const u32 f = row[pidx];
                const double* tp = ta.tet + (size_t)f * 12;
                double c[4][3];
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int q = 0; q < 3; ++q) c[k][q] = __ldg(tp + k * 3 + q);
                bool apart = false;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const double tmin = fmin(fmin(c[0][q], c[1][q]), fmin(c[2][q], c[3][q]));
                    const double tmax = fmax(fmax(c[0][q], c[1][q]), fmax(c[2][q], c[3][q]));
                    apart = apart || tmin > hi[q] || tmax < lo[q];
                }
                if (apart) continue;
                // work copy of the cell
                __syncwarp();
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int sl = lane + 32 * h2;
                    if (((h2 ? st.used_hi : st.used_lo) >> lane) & 1u) {
                        P2[0][sl] = P[0][sl]; P2[1][sl] = P[1][sl]; P2[2][sl] = P[2][sl];
                        V2[sl] = V[sl]; T2[sl] = T[sl];
                    }
                }
                __syncwarp();
                VcState s2 = st;
                bool empty = false;
#pragma unroll 1
                for (int lf = 0; lf < 4 && !empty; ++lf) {
                    const int i0 = (lf + 1) & 3, i1 = (lf + 2) & 3, i2 = (lf + 3) & 3;
                    const double ux = c[i1][0] - c[i0][0], uy = c[i1][1] - c[i0][1], uz = c[i1][2] - c[i0][2];
                    const double vx = c[i2][0] - c[i0][0], vy = c[i2][1] - c[i0][1], vz = c[i2][2] - c[i0][2];
                    double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
                    double sg = 0.0;
                    sg += nx * (c[lf][0] - c[i0][0]); sg += ny * (c[lf][1] - c[i0][1]); sg += nz * (c[lf][2] - c[i0][2]);
                    if (sg == 0.0) { empty = true; break; }               // flat tet: no volume
                    if (sg < 0.0) { nx = -nx; ny = -ny; nz = -nz; }
                    double dd = 0.0;
                    dd -= nx * c[i0][0]; dd -= ny * c[i0][1]; dd -= nz * c[i0][2];
                    vc_clip<true>(P2, V2, T2, B, IT, FS, PLN, s2, pix, piy, piz, nx, ny, nz, dd, 250u + (u32)lf, lane);
                    if (s2.overflow) { gave_up = true; break; }
                    if ((s2.used_lo | s2.used_hi) == 0u) empty = true;
                }
                if (gave_up || empty) continue;
                if (nn > 0) {
                    const double R2p = vc_radius2(P2, s2, pix, piy, piz, lane);
                    if (!(last_d > 4.1 * R2p)) any_exh = true;
                }
                vc_integrate(P2, V2, T2, B, s2, pix, piy, piz, a.mode, (u32)ta.tet_inner[f], lane, acc_s, acc_x, acc_y, acc_z);
            }
        }
        if (gave_up) {
            if (lane == 0) { const u32 pos = atomicAdd(ta.ovf_n, 1u); ta.ovf_list[pos] = s; }
            continue;
        }
        if (any_exh && a.check_SR && nn + 1 < a.S && nn < B200CVT_KMAX_DEV && a.redo_list) {
            if (lane == 0) { const u32 pos = atomicAdd(a.redo_n, 1u); a.redo_list[pos] = s; }
            continue;
        }
        acc_s = warp_sum(acc_s); acc_x = warp_sum(acc_x); acc_y = warp_sum(acc_y); acc_z = warp_sum(acc_z);
        if (lane == 0) {
            a.out_s[s] = acc_s;
            a.out_v[(size_t)s * 3 + 0] = acc_x; a.out_v[(size_t)s * 3 + 1] = acc_y; a.out_v[(size_t)s * 3 + 2] = acc_z;
            uint8_t f8 = (uint8_t)(a.flags[s] & ~(uint8_t)(1 | 4 | 8));
            if (any_exh) {
                if (!a.check_SR) f8 |= 1;
                else if (nn + 1 >= a.S) { }
                else if (nn >= B200CVT_KMAX_DEV) f8 |= 8;
            }
            a.flags[s] = f8;
        }
    }
}
