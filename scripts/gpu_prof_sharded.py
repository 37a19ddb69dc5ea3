"""Per-kernel times of the sharded path on ONE GPU: N handles (threads + a barrier play the all-gather, as in
tests/test_gpu_parity.py::test_sharded_lloyd_two_partitions_one_gpu) on the N-times larger weak-scaling workload.
Run under `ncu --metrics gpu__time_duration.sum`; every kernel appears once per rank and evaluation."""
import sys, os, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graphitethree_b200 import capi, sharding
import bench
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
V, F, X = bench.workload(world, False)
S, dim = X.shape
chunk = sharding.chunk_doubles(dim, S, world)
shared = torch.zeros(chunk * world, dtype=torch.float64, device="cuda")
barrier = threading.Barrier(world)

def worker(rank):
    h = capi.Handle(dim)
    h.set_mesh(V, F)
    h.set_partition(rank, world)
    sl = torch.zeros(chunk, dtype=torch.float64, device="cuda")
    al = torch.zeros(chunk * world, dtype=torch.float64, device="cuda")
    def exchange():
        shared[rank * chunk:(rank + 1) * chunk].copy_(sl)
        torch.cuda.synchronize(); barrier.wait()
        al.copy_(shared)
        torch.cuda.synchronize(); barrier.wait()
        return 0
    h.set_exchange(sl.data_ptr(), al.data_ptr(), chunk, exchange)
    xd = torch.from_numpy(X).cuda()
    h.set_seeds_device(xd.data_ptr(), S)
    h.lloyd_device(iters - 1)
    h.stats()
    h.lloyd_device(1)
    st = h.stats()
    print(rank, {k: st[k] for k in ('candidate_pairs', 'facets_uncertified', 'facets_skipped_far_from_owned_seeds', 'facets_subdivided')}, flush=True)
    c = h.cumulative()
    print(rank, {k: round(c[k] / max(c["evals"], 1), 3) for k in ("sort", "knn", "pairs", "clip")}, flush=True)
    h.close()
ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
[t.start() for t in ts]
[t.join() for t in ts]
