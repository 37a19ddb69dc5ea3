"""Host-side sharding logic of the multi-GPU path, shared by bench.py, the
torch.distributed harness and the CPU (gloo) tests.

Seeds are Morton-sorted; rank r owns sorted positions [r*L, min((r+1)*L, S)) with
L = ceil(S / nranks). The one exchange per evaluation is an all-gather of each rank's
chunk: L*dim values (updated positions, or gradient) followed by L scalars (per-seed
energy), zero padded. These functions mirror pack_slice_kernel / unpack_all_kernel in
graphitethree_b200/csrc/b200cvt.cu.
"""
import numpy as np


def slice_len(S, nranks):
    return (S + nranks - 1) // nranks


def owned_range(S, rank, nranks):
    L = slice_len(S, nranks)
    b = min(rank * L, S)
    return b, min(b + L, S)


def chunk_doubles(dim, S, nranks):
    return slice_len(S, nranks) * (dim + 1)


def pack_slice(vec_sorted, scal_sorted, rank, nranks):
    """vec_sorted [S, dim], scal_sorted [S] (sorted order) -> chunk [L*(dim+1)]."""
    S, dim = vec_sorted.shape
    L = slice_len(S, nranks)
    b, e = owned_range(S, rank, nranks)
    chunk = np.zeros(L * (dim + 1))
    chunk[:(e - b) * dim] = vec_sorted[b:e].reshape(-1)
    chunk[L * dim:L * dim + (e - b)] = scal_sorted[b:e]
    return chunk


def unpack_all(all_chunks, S, dim, nranks, order):
    """all_chunks [nranks*L*(dim+1)] -> (vec in ORIGINAL order [S, dim], scal in sorted order [S]);
    order[i] = original index of sorted position i."""
    L = slice_len(S, nranks)
    c = all_chunks.reshape(nranks, L * (dim + 1))
    vec_sorted = c[:, :L * dim].reshape(nranks * L, dim)[:S]
    scal_sorted = c[:, L * dim:].reshape(nranks * L)[:S]
    vec = np.empty((S, dim))
    vec[order] = vec_sorted
    return vec, scal_sorted.copy()


class TorchExchange:
    """The all-gather of the sharded path over torch.distributed (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, dim, S, rank, nranks, device, sync=True):
        """sync=False: the library runs on torch's current stream (Handle.set_stream), so the collective is
        ordered on the device and the host does not wait for it."""
        import torch
        self.sync = sync
        self.torch = torch
        self.chunk = chunk_doubles(dim, S, nranks)
        self.slice = torch.zeros(self.chunk, dtype=torch.float64, device=device)
        self.all = torch.zeros(self.chunk * nranks, dtype=torch.float64, device=device)
        self.device = device
        self.count = 0

    def __call__(self):
        import torch.distributed as dist
        dist.all_gather_into_tensor(self.all, self.slice)
        if self.all.is_cuda and self.sync:
            self.torch.cuda.current_stream().synchronize()
        self.count += 1
        return 0
