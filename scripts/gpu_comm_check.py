"""Multi-GPU parity of the in-library communicator (run under gpurun --gpus N).

  torchrun --nproc-per-node N scripts/gpu_comm_check.py   one process per GPU: b200cvt_comm_init (NCCL + CUDA-IPC mailboxes)
  python scripts/gpu_comm_check.py --group N              one process, N GPUs: b200cvt_group_* (threads + peer access)

Both compare N-GPU Lloyd (bit-identical seeds) and Newton (same iteration / evaluation counts, seeds within 1e-10 of the
single-GPU run: the dot products are summed in a different order) against one unsharded handle on GPU 0."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphitethree_b200 import capi, shapes


def inputs(small):
    V, F = shapes.noise_sphere(60 if small else 160)
    X = shapes.sample_surface(V, F, 6000 if small else 60000, 1)
    return V, F, X


def single(V, F, X, nl, nn):
    h = capi.Handle(3, device=0)
    h.set_mesh(V, F)
    x = h.lloyd(X, nl)
    x2, info = h.newton(x, nn, 7)
    h.close()
    return x, x2, info


def check(tag, xl, xn, info, ref):
    rl, rn, rinfo = ref
    ok_l = np.array_equal(xl, rl)
    dn = float(np.abs(xn - rn).max())
    ok_n = info["iters"] == rinfo["iters"] and info["nfev"] == rinfo["nfev"] and dn <= 1e-10
    print("%s: lloyd bit-identical=%s  newton iters %d/%d nfev %d/%d max|dx|=%.3g  -> %s" % (
        tag, ok_l, info["iters"], rinfo["iters"], info["nfev"], rinfo["nfev"], dn, "OK" if ok_l and ok_n else "FAIL"), flush=True)
    return ok_l and ok_n


def main():
    small = "--small" in sys.argv
    nl, nn = 3, 6
    V, F, X = inputs(small)
    if "--group" in sys.argv:
        n = int(sys.argv[sys.argv.index("--group") + 1])
        ref = single(V, F, X, nl, nn)
        g = capi.Group(n, 3)
        g.set_mesh(V, F)
        t0 = time.time()
        xl = g.lloyd(X, nl)
        xn, info = g.newton(xl, nn, 7)
        dt = time.time() - t0
        ok = check("group x%d (%.2f s)" % (g.size, dt), xl, xn, info, ref)
        # cancel: every rank must stop at the same iteration
        calls = []
        try:
            g.lloyd(X, 5, callback=lambda user, it, f, gn: (calls.append(it), 1 if it == 2 else 0)[1])
            ok = False
            print("cancel: not raised")
        except capi.B200CVTError as e:
            print("cancel after iteration %s: code %d" % (calls, e.code))
            ok = ok and e.code == 5 and calls == [1, 2]
        g.close()
        sys.exit(0 if ok else 1)
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cid = capi.comm_unique_id() if rank == 0 else np.zeros(capi.COMM_ID_BYTES, dtype=np.uint8)
    t = torch.from_numpy(cid).cuda()
    dist.broadcast(t, 0)
    h = capi.Handle(3, device=local)
    h.set_mesh(V, F)
    h.comm_init(t.cpu().numpy(), rank, world)
    xl = h.lloyd(X, nl)
    xn, info = h.newton(xl, nn, 7)
    ok = True
    if rank == 0:
        ok = check("comm x%d" % world, xl, xn, info, single(V, F, X, nl, nn))
    # every rank ends with the same seeds
    mine = torch.from_numpy(xn).cuda()
    root = mine.clone()
    dist.broadcast(root, 0)
    same = bool(torch.equal(mine, root))
    if not same:
        print("rank %d: seeds differ from rank 0" % rank, flush=True)
    h.close()
    flag = torch.tensor([0 if (ok and same) else 1], device="cuda")
    dist.all_reduce(flag)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
