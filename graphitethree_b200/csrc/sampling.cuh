// sampling.cuh — initial sampling of the seeds.
//
// Replaces RestrictedVoronoiDiagram::compute_initial_sampling_on_surface / _in_volume (geogram/voronoi/RVD.cpp:1658-1698)
//   = mesh_generate_random_samples_on_surface<DIM> / _in_volume<DIM> (geogram/mesh/mesh_sampling.h:119-199, 280-360):
//   sorted uniforms s_i against the running sum of (element mass / total mass) pick the element of every sample, two
//   (three) more uniforms its barycentric coordinates (Geom::random_point_in_triangle / _in_tetra).
//
// What is parallel runs here: the element masses (one thread per element: a square root each) and the sample points
// (one thread per sample: corner gather + barycentric combination, written straight into the seed array of the handle).
// What is a recurrence stays on the host, because the result has to be the reference's bit for bit: the mt19937_64 stream
// (Numeric::random_float64, reset at every call), the sort of the uniforms, and the running sum of the masses in the
// caller's element order (a parallel scan would round differently and could move a sample across an element boundary).
#pragma once
#include "common.cuh"

// mesh_facet_mass<DIM> (mesh_sampling.h:67-96) / mesh_tetra_mass<3> (:213-262), written at the caller's element index.
// soup: sorted elements [T][PER][D]; w: per-corner weights [T][3] or NULL (surfaces only); perm: sorted index -> caller's
template <int D, int PER>
__global__ void sampling_mass_kernel(const double* soup, const double* w, u32 T, const u32* perm, double* mass_orig) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const double* t = soup + (size_t)i * PER * D;
    double p[PER][D];
#pragma unroll
    for (int k = 0; k < PER; ++k)
#pragma unroll
        for (int c = 0; c < D; ++c) p[k][c] = t[k * D + c];
    double m;
    if (PER == 4) {
        // Geom::tetra_volume (geometry.h:483-525): |dot(p2 - p1, cross(p3 - p1, p4 - p1)) / 6|
        double a[3], b[3], c3[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { a[k] = p[1][k] - p[0][k]; b[k] = p[2][k] - p[0][k]; c3[k] = p[PER - 1][k] - p[0][k]; }
        const double cx = b[1] * c3[2] - c3[1] * b[2], cy = b[2] * c3[0] - c3[2] * b[0], cz = b[0] * c3[1] - c3[0] * b[1];
        m = fabs((a[0] * cx + a[1] * cy + a[2] * cz) / 6.0);
    } else if (D == 3 && !w) {
        // Geom::triangle_area(vec3, vec3, vec3) (geometry.h:346-372)
        const double Ux = p[1][0] - p[0][0], Uy = p[1][1] - p[0][1], Uz = p[1][2] - p[0][2];
        const double Vx = p[2][0] - p[0][0], Vy = p[2][1] - p[0][1], Vz = p[2][2] - p[0][2];
        const double Nx = Uy * Vz - Uz * Vy, Ny = Uz * Vx - Ux * Vz, Nz = Ux * Vy - Uy * Vx;
        m = 0.5 * sqrt(Nx * Nx + Ny * Ny + Nz * Nz);
    } else {
        // Geom::triangle_area, nD (geometry_nd.h:143-156): Heron
        const double ea = sqrt(dist2<D>(p[0], p[1])), eb = sqrt(dist2<D>(p[1], p[2])), ec = sqrt(dist2<D>(p[2], p[0]));
        const double sh = 0.5 * (ea + eb + ec);
        const double A2 = sh * (sh - ea) * (sh - eb) * (sh - ec);
        m = sqrt(fmax(A2, 0.0));
        // Geom::triangle_mass (geometry_nd.h:237-252)
        if (w) m = m / 3.0 * (sqrt(fabs(w[(size_t)i * 3])) + sqrt(fabs(w[(size_t)i * 3 + 1])) + sqrt(fabs(w[(size_t)i * 3 + 2])));
    }
    mass_orig[perm[i]] = m;
}

// one thread per sample: lam[i] = the barycentric weights of the element's corners in corner order
// (Geom::random_point_in_triangle: geometry.h:602-619 for DIM = 3, geometry_nd.h:337-349 otherwise; _in_tetra :363-385)
template <int D, int PER>
__global__ void sampling_points_kernel(const double* soup, const u32* inv_perm, const u32* elem, const double* lam, u32 S, double* x) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const double* t = soup + (size_t)inv_perm[elem[i]] * PER * D;
    const double* l = lam + (size_t)i * 4;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        double v;
        if (PER == 3 && D == 3) v = l[0] * t[c] + l[1] * t[D + c] + l[2] * t[2 * D + c];
        else if (PER == 3) v = l[0] * t[c] + l[1] * t[D + c] + l[2] * t[2 * D + c];
        else v = l[0] * t[c] + l[1] * t[D + c] + l[2] * t[2 * D + c] + l[3] * t[3 * D + c];
        x[(size_t)i * D + c] = v;
    }
}
