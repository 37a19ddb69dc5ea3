// mesh_prep.cuh — b200cvt_set_mesh on the device: validation, bounding box and total measure, Morton order of the elements
// and the corner "soup" the evaluation kernels read.
//
// Replaces the host-side preparation of the reference's borrowed GEO::Mesh for this path: RVD_Nd_Impl's constructor
// (geogram/voronoi/RVD.cpp:120-170: vertex stride, optional "weight" attribute) and the Hilbert reordering of
// create_threads() (RVD.cpp:2390-2395, geogram/mesh/mesh_partition.cpp:59-72) — here a Morton order of the DEVICE copy
// only, the caller's mesh is never touched. At C3 (20 M triangles) the host version of this step took 2.7 s against
// 38 ms for a whole Lloyd iteration.
#pragma once
#include "common.cuh"

#define MESHPREP_BLOCKS 1024
#define MESHPREP_THREADS 256

struct MeshPartial { double lo[3], hi[3], measure; unsigned int bad; unsigned int pad; };

// pass 1: index validation, bounding box of the referenced vertices, total area (PER = 3) or volume (PER = 4);
// one partial per block, added by the host in block order (deterministic)
template <int PER>
__global__ void __launch_bounds__(MESHPREP_THREADS)
mesh_bounds_kernel(const double* vert, u32 nv, u32 stride, const u32* elems, u32 ne, MeshPartial* out) {
    __shared__ double sm[MESHPREP_THREADS / 32][8];
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, measure = 0.0;
    unsigned int bad = 0;
    for (u32 f = blockIdx.x * blockDim.x + threadIdx.x; f < ne; f += gridDim.x * blockDim.x) {
        double p[PER][3];
        bool ok = true;
#pragma unroll
        for (int lv = 0; lv < PER; ++lv) {
            const u32 v = elems[(size_t)f * PER + lv];
            if (v >= nv) { ok = false; bad = 1; }
            const double* q = vert + (size_t)(ok ? v : 0) * stride;
#pragma unroll
            for (int a = 0; a < 3; ++a) p[lv][a] = q[a];
        }
        if (!ok) continue;
#pragma unroll
        for (int lv = 0; lv < PER; ++lv)
#pragma unroll
            for (int a = 0; a < 3; ++a) { lo[a] = fmin(lo[a], p[lv][a]); hi[a] = fmax(hi[a], p[lv][a]); }
        double e1[3], e2[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) { e1[a] = p[1][a] - p[0][a]; e2[a] = p[2][a] - p[0][a]; }
        const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
        if (PER == 3) measure += 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
        else measure += fabs(cx * (p[PER - 1][0] - p[0][0]) + cy * (p[PER - 1][1] - p[0][1]) + cz * (p[PER - 1][2] - p[0][2])) / 6.0;
    }
    // block reduction: fixed tree
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double vals[7] = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], measure};
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) vals[k] = fmin(vals[k], __shfl_xor_sync(B200_FULL, vals[k], m));
#pragma unroll
        for (int k = 3; k < 6; ++k) vals[k] = fmax(vals[k], __shfl_xor_sync(B200_FULL, vals[k], m));
        vals[6] += __shfl_xor_sync(B200_FULL, vals[6], m);
        bad |= __shfl_xor_sync(B200_FULL, bad, m);
    }
    __shared__ unsigned int sbad[MESHPREP_THREADS / 32];
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) sm[w][k] = vals[k];
        sbad[w] = bad;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        MeshPartial r;
        for (int a = 0; a < 3; ++a) { r.lo[a] = 1e300; r.hi[a] = -1e300; }
        r.measure = 0.0; r.bad = 0; r.pad = 0;
        for (int i = 0; i < MESHPREP_THREADS / 32; ++i) {
            for (int a = 0; a < 3; ++a) { r.lo[a] = fmin(r.lo[a], sm[i][a]); r.hi[a] = fmax(r.hi[a], sm[i][3 + a]); }
            r.measure += sm[i][6];
            r.bad |= sbad[i];
        }
        out[blockIdx.x] = r;
    }
}

// pass 2: 30-bit Morton code of every element centroid (10 bits per axis over the bounding box), value = element index
template <int PER>
__global__ void mesh_codes_kernel(const double* vert, u32 stride, const u32* elems, u32 ne, double lox, double loy, double loz,
                                  double sc, u32* keys, u32* vals) {
    const u32 f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= ne) return;
    const double lo[3] = {lox, loy, loz};
    u32 q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double c = 0.0;
#pragma unroll
        for (int lv = 0; lv < PER; ++lv) c += vert[(size_t)elems[(size_t)f * PER + lv] * stride + a];
        const double t = (c * (1.0 / PER) - lo[a]) * sc;
        q[a] = (u32)fmin(1023.0, fmax(0.0, t));
    }
    u32 code = 0;
#pragma unroll
    for (int b = 0; b < 10; ++b)
        code |= (((q[0] >> b) & 1u) << (3 * b)) | (((q[1] >> b) & 1u) << (3 * b + 1)) | (((q[2] >> b) & 1u) << (3 * b + 2));
    keys[f] = code;
    vals[f] = f;
}

// pass 3: corner coordinates (and weights) of the elements in sorted order
template <int D, int PER>
__global__ void mesh_soup_kernel(const double* vert, u32 stride, const u32* elems, const u32* order, u32 ne, const double* weights,
                                 double* soup, double* soup_w) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ne) return;
    const u32 f = order[i];
#pragma unroll
    for (int lv = 0; lv < PER; ++lv) {
        const u32 v = elems[(size_t)f * PER + lv];
        const double* p = vert + (size_t)v * stride;
#pragma unroll
        for (int c = 0; c < D; ++c) soup[((size_t)i * PER + lv) * D + c] = p[c];
        if (soup_w && lv < 3) soup_w[(size_t)i * 3 + lv] = weights[v];
    }
}

// max |v|^2 over the corners of the soup (all D coordinates): bounds the reference's FP64 rounding on global
// coordinates in the FP32 side-test filter of the clip kernel (clip_flat.cuh). Positive doubles order like their bits.
__global__ void soup_vmax2_kernel(const double* soup, size_t ncorners, int D, unsigned long long* out) {
    double mx = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < ncorners; i += (size_t)gridDim.x * blockDim.x) {
        double q = 0.0;
        for (int c = 0; c < D; ++c) { const double v = soup[i * D + c]; q += v * v; }
        mx = fmax(mx, q);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) mx = fmax(mx, __shfl_xor_sync(B200_FULL, mx, m));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(mx));
}
