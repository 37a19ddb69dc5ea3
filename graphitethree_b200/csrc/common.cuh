// common.cuh — shared types and device helpers for the B200 CVT/RVD kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <stdexcept>

typedef uint32_t u32;
typedef uint64_t u64;

#define B200_FULL 0xffffffffu
#define B200_NONE 0xffffffffu

struct CudaError : public std::runtime_error {
    explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};

#define CUDA_CHECK(expr)                                                                      \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
            throw CudaError(std::string(#expr) + ": " + cudaGetErrorString(e_) + " at " +     \
                            __FILE__ + ":" + std::to_string(__LINE__));                       \
    } while (0)

// Uniform grid over the first three coordinates of the seeds. Cells are cubes of edge h;
// the cell table is indexed by a Morton code with per-axis bit counts, so that sorting the
// seeds by cell id is a Morton sort (contiguous ranges = compact regions, used for sharding).
struct GridParams {
    double lo[3];
    double h, inv_h;
    int res[3];
    int bits[3];
    int total_bits;
    u32 ncells;
    const u32* mtab;   // device table [3][1024]: the Morton bits of each axis coordinate (NULL: compute)
};

__host__ __device__ inline u32 morton_encode(const GridParams& g, int cx, int cy, int cz) {
#ifdef __CUDA_ARCH__
    if (g.mtab) return __ldg(g.mtab + cx) | __ldg(g.mtab + 1024 + cy) | __ldg(g.mtab + 2048 + cz);
#endif
    u32 code = 0;
    int out = 0;
    u32 x = (u32)cx, y = (u32)cy, z = (u32)cz;
#pragma unroll 1
    for (int b = 0; b < 11; ++b) {
        if (b < g.bits[0]) { code |= ((x >> b) & 1u) << out; ++out; }
        if (b < g.bits[1]) { code |= ((y >> b) & 1u) << out; ++out; }
        if (b < g.bits[2]) { code |= ((z >> b) & 1u) << out; ++out; }
    }
    return code;
}

__host__ __device__ inline int grid_coord(const GridParams& g, double v, int a) {
    double t = floor((v - g.lo[a]) * g.inv_h);
    int c = (t < 0.0) ? 0 : ((t >= (double)g.res[a]) ? g.res[a] - 1 : (int)t);
    return c;
}

// per-cell status byte of sharded runs (b200cvt.cu: mark_cells_kernel; facet_pairs.cuh: facet_filter_kernel)
#define CELLF_WITHIN1 1u      // an owned seed within one cell
#define CELLF_WITHIN2 2u
#define CELLF_WITHIN3 4u
#define CELLF_OCCUPIED 8u     // some seed lies in the cell

// One sorted seed record: D coordinates + original index. 32 B for D=3, 64 B for D=6.
template <int D> struct SeedRec;
template <> struct __align__(32) SeedRec<3> { double p[3]; long long orig; };
template <> struct __align__(64) SeedRec<6> { double p[6]; long long orig; long long pad; };

// Sum over coordinates, in order, of (b-a)^2 — bit-equal to Geom::distance2
// (geogram/basic/geometry_nd.h:65-74); this file is compiled with -fmad=false.
template <int D>
__device__ __forceinline__ double dist2(const double* a, const double* b) {
    double r = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        double d = b[c] - a[c];
        r += d * d;
    }
    return r;
}

__device__ __forceinline__ double shfl_xor_d(double v, int m) {
    return __shfl_xor_sync(B200_FULL, v, m);
}
__device__ __forceinline__ double shfl_d(double v, int src) {
    return __shfl_sync(B200_FULL, v, src);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(B200_FULL, v, m);
    return v;
}
