"""BASELINE.json config C3 as stated: procedural trefoil tube, 20 M triangles, 5 M seeds, Lloyd iterations, seeds sharded by
Morton range over N GPUs (strong scaling: the total size is fixed). Launch with torchrun, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/gpu_sharded_configs.py [--small]
Rank 0 prints one JSON line: Lloyd iterations/s (device time, max over ranks) and seed-iterations/s."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from graphitethree_b200 import capi, shapes, sharding

rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
small = "--small" in sys.argv
c5 = "--c5" in sys.argv          # volumetric config C5 instead: 10.1 M Kuhn tets of the unit cube, 1.01 M seeds
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
if c5:
    V, F = shapes.kuhn_cube(20 if small else 119)
    S = F.shape[0] // 10
    X = 0.01 + 0.98 * np.random.default_rng(5).random((S, 3))
else:
    V, F = shapes.trefoil_tube(1000 if small else 10000, 100 if small else 1000)
    S = 50000 if small else 5000000
    X = shapes.sample_surface(V, F, S, 1)
stream = torch.cuda.Stream()
h = capi.Handle(3, volumetric=c5, device=local_rank)
h.set_stream(stream.cuda_stream)
t0 = time.time(); h.set_mesh(V, F); t_mesh = time.time() - t0
h.set_partition(rank, world)
if world > 1:
    with torch.cuda.stream(stream):
        ex = sharding.TorchExchange(3, S, rank, world, torch.device("cuda", local_rank), sync=False)

        def exchange():
            with torch.cuda.stream(stream):
                return ex()
    h.set_exchange(ex.slice.data_ptr(), ex.all.data_ptr(), ex.chunk, exchange)
xd = torch.from_numpy(X).cuda()
with torch.cuda.stream(stream):
    h.set_seeds_device(xd.data_ptr(), S)
    h.lloyd_device(4)                      # relaxation + warm-up (buffers, neighbour-list coherence)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
h.cumulative(reset=True)
iters = 5 if c5 else 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record(stream)
    h.lloyd_device(iters)
    e1.record(stream)
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
c = h.cumulative()
x = h.get_seeds()
if rank == 0:
    t = float(ms.item()) * 1e-3
    print(json.dumps({"config": "C5" if c5 else "C3", "n_gpus": world, "elements": int(F.shape[0]), "seeds": S, "lloyd_iterations": iters,
                      "ms_per_iteration": 1e3 * t / iters, "lloyd_iterations_per_s": iters / t, "seed_iterations_per_s": S * iters / t,
                      "set_mesh_s": round(t_mesh, 2), "rank0_phase_ms": {k: round(c[k] / max(c["evals"], 1), 3) for k in ("sort", "knn", "pairs", "clip", "clip_kernel")},
                      "seeds_finite": bool(np.isfinite(x).all())}), flush=True)
h.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
