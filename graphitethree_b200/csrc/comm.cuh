// comm.cuh — multi-GPU plumbing inside the library: one NCCL communicator per handle for the bulk
// exchanges (all-gather of seed positions, reduce-scatter of gradients over NVLink) and peer-memory mailboxes for the
// scalar reductions of the L-BFGS kernels (a few doubles per dot product, summed inside the cooperative kernel by
// stores into every peer's HBM over NVLink/NVSwitch — no launch, no host).
//
// NCCL is loaded with dlopen on first use: a single-GPU caller never needs it, and a process that already holds a
// libnccl.so.2 (PyTorch's bundled copy) shares it.
#pragma once
#include "common.cuh"
#include <nccl.h>
#include <dlfcn.h>

#define B200_MAX_RANKS 16

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; names[i] && !api.lib; ++i) api.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
            api.CommInitAll = (decltype(api.CommInitAll))dlsym(api.lib, "ncclCommInitAll");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
            api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
            api.ReduceScatter = (decltype(api.ReduceScatter))dlsym(api.lib, "ncclReduceScatter");
            api.GroupStart = (decltype(api.GroupStart))dlsym(api.lib, "ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.lib, "ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
        }
    }
    if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.ReduceScatter || !api.CommDestroy)
        throw std::runtime_error("NCCL (libnccl.so.2) is not available: multi-GPU runs need it");
    return api;
}

#define NCCL_CHECK(expr)                                                                                  \
    do {                                                                                                  \
        ncclResult_t r_ = (expr);                                                                         \
        if (r_ != ncclSuccess)                                                                            \
            throw std::runtime_error(std::string(#expr) + ": " +                                          \
                                     (nccl_api().GetErrorString ? nccl_api().GetErrorString(r_) : "NCCL error") + " at " + \
                                     __FILE__ + ":" + std::to_string(__LINE__));                          \
    } while (0)

// One mailbox entry: written by exactly one peer (its rank selects the entry), read by the owner. 384 bytes.
#define B200_BOX_DOUBLES 46
struct __align__(128) PeerBox {
    double v[B200_BOX_DOUBLES];
    unsigned long long seq;
    unsigned long long pad;
};

// Device view of the communicator, passed by value to the L-BFGS kernels.
struct PeerComm {
    int rank, nranks;
    PeerBox* boxes[B200_MAX_RANKS];      // boxes[p]: rank p's mailbox array [2][B200_MAX_RANKS] in THIS process's address space
    unsigned long long* seq;              // local: number of cross-rank reductions done so far (identical on every rank)
    double* gtot;                         // local: [2][8] the last totals, published to the other blocks of the grid
    unsigned int* error;                  // local: set when a peer did not answer within the timeout
};

// Sum of K doubles over all ranks; called by the first warp of ONE block with the same `vals` in every lane. Every rank
// pushes its values into everybody's mailbox (NVLink stores), waits for the N entries of its own mailbox and adds them in
// rank order, so all ranks hold bit-identical totals. Two entry sets alternate: a rank can be at most one reduction ahead
// of its slowest peer (it needs that peer's entry to finish the current one).
template <int K>
__device__ __forceinline__ void peer_allreduce(const PeerComm& pc, const double* vals, double* result) {
    static_assert(K <= B200_BOX_DOUBLES, "mailbox entry too small");
    const int lane = threadIdx.x & 31;
    const unsigned long long seq = *((volatile unsigned long long*)pc.seq) + 1ull;
    const int par = (int)(seq & 1ull);
    if (lane < pc.nranks) {
        volatile PeerBox* b = pc.boxes[lane] + par * B200_MAX_RANKS + pc.rank;
#pragma unroll
        for (int j = 0; j < K; ++j) b->v[j] = vals[j];
        __threadfence_system();
        b->seq = seq;
    }
    double got[K];
#pragma unroll
    for (int j = 0; j < K; ++j) got[j] = 0.0;
    if (lane < pc.nranks) {
        volatile PeerBox* m = pc.boxes[pc.rank] + par * B200_MAX_RANKS + lane;
        const long long t0 = clock64();
        bool ok = true;
        while (m->seq != seq) {
            __nanosleep(64);
            if (clock64() - t0 > 40000000000ll) { ok = false; break; }      // ~20 s: a peer is gone
        }
        __threadfence_system();
        if (ok) {
#pragma unroll
            for (int j = 0; j < K; ++j) got[j] = m->v[j];
        } else atomicExch(pc.error, 1u);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double t = 0.0;
        for (int r = 0; r < pc.nranks; ++r) t += __shfl_sync(B200_FULL, got[j], r);
        result[j] = t;
    }
    if (lane == 0) *((volatile unsigned long long*)pc.seq) = seq;
}

// The same for n <= B200_BOX_DOUBLES values held in memory (vals, result: local global / shared memory; may alias): the
// whole first warp of ONE block calls it. Lane r < nranks stores the n values into rank r's mailbox and waits for rank r's
// entry in its own; then lane j adds the j-th values of all entries in rank order.
__device__ __forceinline__ void peer_allreduce_n(const PeerComm& pc, const double* vals, int n, double* result) {
    const int lane = threadIdx.x & 31;
    const unsigned long long seq = *((volatile unsigned long long*)pc.seq) + 1ull;
    const int par = (int)(seq & 1ull);
    if (lane < pc.nranks) {
        volatile PeerBox* b = pc.boxes[lane] + par * B200_MAX_RANKS + pc.rank;
        for (int j = 0; j < n; ++j) b->v[j] = vals[j];
        __threadfence_system();
        b->seq = seq;
        volatile PeerBox* m = pc.boxes[pc.rank] + par * B200_MAX_RANKS + lane;
        const long long t0 = clock64();
        while (m->seq != seq) {
            __nanosleep(32);
            if (clock64() - t0 > 40000000000ll) { atomicExch(pc.error, 1u); break; }      // ~20 s: a peer is gone
        }
        __threadfence_system();
    }
    __syncwarp();
    for (int j = lane; j < n; j += 32) {
        double t = 0.0;
        for (int r = 0; r < pc.nranks; ++r) t += ((volatile PeerBox*)(pc.boxes[pc.rank] + par * B200_MAX_RANKS + r))->v[j];
        result[j] = t;
    }
    __syncwarp();
    if (lane == 0) *((volatile unsigned long long*)pc.seq) = seq;
}

// ---------------------------------------------------------------------------------------
// Bulk exchanges of the sharded Newton loop through peer memory (the NCCL calls remain as the fallback path):
//   the gradient of a seed is STORED by the rank that evaluated it straight into the L-BFGS slice of the rank that owns it,
//   the new trial point of a slice is stored by its owner into every rank's seed array,
// each followed by a mailbox barrier. No staging buffer, no zero fill, no collective launch.
// ---------------------------------------------------------------------------------------
struct PeerBufs {
    double* x[B200_MAX_RANKS];       // every rank's seed array (original order, padded to nranks * L rows)
    double* g[B200_MAX_RANKS];       // every rank's L-BFGS gradient slice (L rows)
};

// barrier over the ranks = a mailbox reduction of nothing; everything this GPU stored before it (earlier kernels of the
// stream included) is visible to a peer once the peer has passed its own barrier
__global__ void peer_barrier_kernel(PeerComm pc) {
    __threadfence_system();
    double z = 0.0, r;
    peer_allreduce<1>(pc, &z, &r);
    __threadfence_system();
}

// this rank's slice of the seed array -> the same rows of every other rank's seed array
__global__ void peer_push_slice_kernel(PeerBufs pb, int rank, int nranks, size_t off, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double v = pb.x[rank][off + i];
        for (int p = 0; p < nranks; ++p)
            if (p != rank) pb.x[p][off + i] = v;
    }
}
