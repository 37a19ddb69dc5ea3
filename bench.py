#!/usr/bin/env python
"""bench.py — CVT seed-iterations/s of the B200 path next to the reference CPU RVD.

Contract: python bench.py --gpus N --steps K --warmup W [--impl reference]; rank 0 prints ONE JSON line.

Workload (BASELINE.json configs[1], weak-scaled with N): noise-displaced sphere with ~2 M triangles
per GPU (frequency round(316*sqrt(N)) geodesic icosphere, 3-octave value noise, amplitude 0.1) and
200 000 seeds per GPU drawn uniformly by area (fixed RNG seed). One STEP = one CVT job on that input:
10 Lloyd iterations + Newton_iterations(30, m=7), restarted from the same initial seeds every step.
metric = seed-iterations/s = S * (Lloyd iterations + Newton function evaluations) / seconds.

  value : seeds resident in HBM when the timed region starts (device-resident loops)
  e2e   : the same job through the host-pointer C-ABI calls a geogram adapter makes
          (b200cvt_lloyd / b200cvt_newton with pinned host seeds: H2D + D2H inside the timed region)
  N > 1 : one process per GPU (torchrun), seeds sharded by Morton range, mesh replicated, one NCCL
          all-gather per evaluation through torch.distributed (scaling "weak")
  --impl reference : the unmodified reference (oracle/_ref, built from /root/reference by
          oracle/Makefile.ref) on the host cores, THE SAME JOB per step (10 Lloyd + Newton_iterations(30, 7) from the
          same initial seeds); the number of steps actually run is bounded by a time budget and stated in the line
  "c3"  : second timed leg, BASELINE.json configs[2] as stated: trefoil tube, 20 M triangles, 5 M seeds, Lloyd
          iterations, seeds sharded by Morton range over the N GPUs (strong scaling), Lloyd iterations/s
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

# The contract is ONE JSON line on stdout. Libraries loaded later write banners to file descriptor 1 (NCCL prints its
# version there from C): everything but the final line is sent to stderr, the line itself goes to the original stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LLOYD_ITERS, NEWTON_ITERS, NEWTON_M = 10, 30, 7
SEEDS_PER_GPU = 200000
BASE_FREQ = 316


def workload(n_gpus, small=False):
    from graphitethree_b200 import shapes
    if small:
        V, F = shapes.noise_sphere(60)
        S = 8000 * n_gpus
    else:
        V, F = shapes.noise_sphere(int(round(BASE_FREQ * n_gpus ** 0.5)))
        S = SEEDS_PER_GPU * n_gpus
    X = shapes.sample_surface(V, F, S, 1)
    return V, F, X


def workload_name(n_gpus, T, S):
    return ("noise-displaced sphere %d triangles, %d seeds, %d Lloyd + %d Newton (HLBFGS m=%d) iterations per step"
            % (T, S, LLOYD_ITERS, NEWTON_ITERS, NEWTON_M))


def workload_config(T, S):
    """The workload, identical in both arms (the arm-specific facts go under "run")."""
    return {"workload": workload_name(0, T, S), "seeds": int(S), "triangles": int(T), "lloyd_iterations": LLOYD_ITERS,
            "newton_iterations": NEWTON_ITERS, "newton_m": NEWTON_M,
            "l2": "inputs larger than L2 (facet table %d MB, seeds re-sorted every evaluation)" % (T * 72 // 2 ** 20)}


REF_BUDGET_S = 150.0     # wall-clock budget of the timed steps of the reference arm


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [t.strip() for t in line.split(",")]
                if len(p) < 8:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for n, v in zip(names, p[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
        out["reasons"] = sorted(reasons)
        return out


def algorithmic_flops_per_seed_iteration(V, F, states, func_grad):
    """F = 17 N_planes + 16 N_plane*vertex + 18 N_intersections + C N_triangles, counted on
    the oracle for the same inputs (C = 52 for Lloyd, 100 for energy+gradient), per seed."""
    from oracle import port
    tot = 0.0
    for x in states:
        e = port.surface_eval(V, F, x, 1 if func_grad else 0, bool(func_grad))
        c = e.counters
        Fl = 17.0 * c["planes"] + 16.0 * c["plane_vertex"] + 18.0 * c["intersections"] + (100.0 if func_grad else 52.0) * c["triangles"]
        tot += Fl / x.shape[0]
    return tot / len(states)


def run_reference(args, rank, world):
    """The reference's own CPU implementation on the host cores (all threads): the same job as the B200 arm per step."""
    if rank != 0:
        return
    from oracle import ref
    V, F, X = workload(args.gpus, args.small)
    S = X.shape[0]
    line = {"impl": "reference", "metric": "CVT seed-iterations/sec", "unit": "seed-iterations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(F.shape[0], S)}
    if ref.available():
        kind = "reference"
        r = ref.RefCVT(V, F, multithread=True)
        cores = ref.RefCVT.nb_threads()

        def step():
            r.set_points(X)
            c0 = r.counters()["funcgrad"]
            t = r.lloyd(LLOYD_ITERS) + r.newton(NEWTON_ITERS, NEWTON_M)
            return t, LLOYD_ITERS + (r.counters()["funcgrad"] - c0)
    else:
        from oracle import port
        kind, cores = "port", 1

        def step():
            t0 = time.time()
            x, _ = port.lloyd(V, F, X, LLOYD_ITERS)
            x, info = port.newton(V, F, x, NEWTON_ITERS, NEWTON_M)
            return time.time() - t0, LLOYD_ITERS + info["nfev"]
    # one untimed job at most (thread pool, Hilbert partition of the mesh); a CPU has no clocks to warm
    n_warm = min(args.warmup, 1)
    t_first = 0.0
    for _ in range(n_warm):
        t_first, _ = step()
    tt, ev, done = 0.0, 0, 0
    while done < args.steps:
        t, e = step()
        tt += t; ev += e; done += 1
        if tt + t > REF_BUDGET_S:        # the next step would overrun the budget
            break
    value = S * ev / tt
    sample = ("the full job (%d Lloyd + Newton_iterations(%d, m=%d) = %d evaluations) on the full %d-seed / %d-triangle input; "
              "%d of the %d requested steps run inside a %.0f s budget, %d untimed warm-up job(s)" % (
                  LLOYD_ITERS, NEWTON_ITERS, NEWTON_M, ev // max(done, 1), S, F.shape[0], done, args.steps, REF_BUDGET_S, n_warm))
    line.update({"value": value, "ms_per_step": 1e3 * tt / max(done, 1), "steps_timed": done,
                 "run": {"evaluations_per_step": ev // max(done, 1), "parallelism": "%d host threads" % cores, "wall_s": tt},
                 "cpu_baseline": {"value": value, "unit": "seed-iterations/s", "cores": cores, "kind": kind, "sample": sample},
                 "e2e": {"value": value, "unit": "seed-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    emit(line)


def comm_setup(h, rank, world, torch, dist, capi):
    """N > 1: the handle becomes one rank of an in-library communicator (NCCL all-gather / reduce-scatter + peer
    mailboxes). The 128-byte id is created by rank 0 and broadcast over the torch.distributed process group."""
    if world == 1:
        return
    cid = capi.comm_unique_id() if rank == 0 else np.zeros(capi.COMM_ID_BYTES, dtype=np.uint8)
    t = torch.from_numpy(cid).cuda()
    dist.broadcast(t, 0)
    h.comm_init(t.cpu().numpy(), rank, world)


def c3_leg(args, rank, local_rank, world, torch, dist, capi, sharding):
    """BASELINE.json configs[2] as stated: procedural trefoil-knot tube, 20 M triangles, 5 M seeds, Lloyd iterations, seeds
    sharded by Morton range over the N GPUs (STRONG scaling: total size fixed). Device time, max over ranks."""
    from graphitethree_b200 import shapes
    small = args.small
    V, F = shapes.trefoil_tube(1000 if small else 10000, 100 if small else 1000)
    S = 50000 if small else 5000000
    X = shapes.sample_surface(V, F, S, 1)
    stream = torch.cuda.Stream()
    h = capi.Handle(3, device=local_rank)
    h.set_stream(stream.cuda_stream)
    t0 = time.time()
    h.set_mesh(V, F)
    t_mesh = time.time() - t0
    comm_setup(h, rank, world, torch, dist, capi)
    xd = torch.from_numpy(X).cuda()
    warm, iters = 4, 20
    with torch.cuda.stream(stream):
        h.set_seeds_device(xd.data_ptr(), S)
        h.lloyd_device(warm)                   # relaxation + warm-up (buffers, neighbour-list coherence)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    h.cumulative(reset=True)
    l0 = h.launch_count()
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
    launches = h.launch_count() - l0
    x = h.get_seeds()
    t = float(ms.item()) * 1e-3
    out = {"workload": "trefoil-knot tube %d triangles, %d seeds, Lloyd iterations, Morton-range x%d (strong scaling)" % (F.shape[0], S, world),
           "n_gpus": world, "triangles": int(F.shape[0]), "seeds": S, "lloyd_iterations_timed": iters, "warmup_iterations": warm,
           "ms_per_iteration": 1e3 * t / iters, "lloyd_iterations_per_s": iters / t, "seed_iterations_per_s": S * iters / t,
           "target_lloyd_iterations_per_s_at_8_gpus": 50.0, "set_mesh_s": round(t_mesh, 2), "gpu_launches": int(launches),
           "rank0_phase_ms_per_iteration": {k: round(c[k] / max(c["evals"], 1), 3) for k in ("sort", "knn", "pairs", "clip", "clip_kernel")},
           "seeds_finite": bool(np.isfinite(x).all()), "scaling": "strong"}
    h.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--small", action="store_true", help="tiny workload for plumbing checks (not a bench value)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c3", action="store_true", help="skip the second timed leg (C3: 20 M triangles / 5 M seeds, Lloyd)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__
    if rank == 0 or not os.path.exists(os.path.join(ROOT, "graphitethree_b200", "libb200cvt.so")):
        __graft_entry__.build()
    from graphitethree_b200 import capi, sharding
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus

    V, F, X = workload(args.gpus, args.small)
    S, dim = X.shape
    stream = torch.cuda.Stream()
    h = capi.Handle(3, device=local_rank)
    h.set_stream(stream.cuda_stream)
    h.set_mesh(V, F)
    comm_setup(h, rank, world, torch, dist, capi)
    x0_dev = torch.from_numpy(X).cuda()
    x_pin = torch.empty((S, dim), dtype=torch.float64).pin_memory()
    x_pin_np = x_pin.numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    evals = {}

    def step_resident():
        h.set_seeds_device(x0_dev.data_ptr(), S)
        h.lloyd_device(LLOYD_ITERS)
        info = h.newton_device(NEWTON_ITERS, NEWTON_M)
        evals["n"] = LLOYD_ITERS + info["nfev"]
        evals["newton"] = info

    # host-pointer calls on the pinned buffer, in place (what a geogram adapter does with points_.data())
    import ctypes as C

    def lib_lloyd(buf):
        capi._check(capi.lib().b200cvt_lloyd(h._h, LLOYD_ITERS, None, buf.ctypes.data_as(C.POINTER(C.c_double)), S,
                                             capi.PROGRESS_CB(), None))
        return buf

    def lib_newton(buf):
        info = np.zeros(4, dtype=np.uint32)
        capi._check(capi.lib().b200cvt_newton(h._h, NEWTON_ITERS, NEWTON_M, None, buf.ctypes.data_as(C.POINTER(C.c_double)), S,
                                              capi.PROGRESS_CB(), None, info.ctypes.data_as(C.POINTER(C.c_uint32))))
        evals["n_e2e"] = LLOYD_ITERS + int(info[1])

    def step_e2e_host():
        x_pin_np[...] = X
        lib_lloyd(x_pin_np)
        lib_newton(x_pin_np)

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        step_resident()
    h.cumulative(reset=True)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = h.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            step_resident()
        e1.record(stream)
    barrier()
    wall = time.time() - t0
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    launches = h.launch_count() - launches0
    cum = h.cumulative(reset=True)
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    # per-rank phase sums (load balance of the Morton ranges): rank 0 reports min / max over the ranks
    by_rank = None
    if world > 1:
        mine = torch.tensor([cum[k] for k in ("sort", "knn", "pairs", "clip")] + [dev_ms], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        by_rank = torch.stack(allr).cpu().numpy()
    n_eval = evals["n"]
    value = S * n_eval * args.steps / (ms_total * 1e-3)
    x_final = h.get_seeds()

    # ---- end-to-end through the host-pointer API (single process per GPU; N = 1 only has host seeds) ----
    e2e = None
    if world == 1:
        for _ in range(2):
            step_e2e_host()
        barrier()
        t0 = time.time()
        for _ in range(args.steps):
            step_e2e_host()
        barrier()
        te = time.time() - t0
        e2e = {"value": S * evals["n_e2e"] * args.steps / te, "unit": "seed-iterations/s",
               "h2d_bytes_per_step": 2 * S * dim * 8, "d2h_bytes_per_step": 2 * S * dim * 8}
    else:
        # sharded runs: seeds enter from pinned host memory on every rank and the result is read back
        x_out_dev = torch.empty_like(x0_dev)
        x_out_pin = torch.empty((S, dim), dtype=torch.float64).pin_memory()

        def step_e2e_sharded():
            x0_dev.copy_(x_pin, non_blocking=True)
            torch.cuda.synchronize()
            step_resident()
            # the result comes back into pinned host memory (a pageable numpy array costs a staged copy at ~1/4 of the rate)
            h.get_seeds_device(x_out_dev.data_ptr())
            x_out_pin.copy_(x_out_dev, non_blocking=True)
            torch.cuda.synchronize()
        x_pin_np[...] = X
        step_e2e_sharded()
        barrier()
        t0 = time.time()
        for _ in range(args.steps):
            step_e2e_sharded()
        barrier()
        te = torch.tensor([time.time() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": S * evals["n"] * args.steps / float(te.item()), "unit": "seed-iterations/s",
               "h2d_bytes_per_step": S * dim * 8, "d2h_bytes_per_step": S * dim * 8}

    c3 = None
    if not args.no_c3:
        try:
            c3 = c3_leg(args, rank, local_rank, world, torch, dist, capi, sharding)
        except Exception as ex_:      # the headline line stands even if the second leg fails
            c3 = {"error": str(ex_)}
    if rank == 0:
        own = (S + world - 1) // world
        line = {"metric": "CVT seed-iterations/sec", "value": value, "unit": "seed-iterations/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(F.shape[0], S),
                "run": {"evaluations_per_step": n_eval, "newton": evals["newton"], "parallelism": "morton-range x%d" % world,
                        "wall_s": wall},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "c3": c3}
        # roofline of the dominant kernel (clip_win_kernel), SURVEY.md §8(d). Durations are CUDA events recorded by the
        # library on its stream around that kernel alone, accumulated over the timed region.
        try:
            fp32, fp64, copy = capi.measure_peaks(local_rank)
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except Exception:
                pass
            hbm = peaks.get("hbm_gbs", 6650.0)
            peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 GB/s (of fallback)"
            T = int(F.shape[0])
            # algorithmic bytes per seed-iteration: seed 8D + kNN list 4k + (T/S)(12 idx + 12 adj + 3*8D) + outputs 8(D+1)
            bytes_unit = 8 * dim + 4 * 20 + (T / S) * (24 + 24 * dim) + 8 * (dim + 1)
            n_launch = max(cum["evals"], 1)
            clip_ms = cum["clip_kernel"] / n_launch
            traffic, traffic_src = None, None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                traffic, traffic_src = tj.get("clip_win_kernel"), tj.get("_note")
            except Exception:
                pass
            ach = bytes_unit * own / (clip_ms * 1e-3) / 1e9
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": traffic,
                                "traffic_source": traffic_src, "kernel": "clip_win_kernel", "peak_source": peak_src, "kernel_ms_per_launch": clip_ms,
                                "algorithmic_bytes_per_seed_iteration": bytes_unit, "units_per_launch": own,
                                "share_of_step": cum["clip_kernel"] / (cum["sort"] + cum["knn"] + cum["pairs"] + cum["clip"]),
                                "note": "FP64 geometry: the kernel is bound by FP64/issue rate, not HBM; see roofline_flops"}
            states = [X, x_final]
            if world == 1 and (args.small or not args.no_cpu_baseline):      # oracle event counts: N = 1 only
                f_lloyd = algorithmic_flops_per_seed_iteration(V, F, states, False)
                f_newton = algorithmic_flops_per_seed_iteration(V, F, [x_final], True)
            else:
                f_lloyd, f_newton = 10400.0, 10670.0
            n_newton = n_eval - LLOYD_ITERS
            flops_step = own * (LLOYD_ITERS * f_lloyd + n_newton * f_newton)
            tf = flops_step / (cum["clip_kernel"] * 1e-3 / args.steps) / 1e12
            tf_phase = flops_step / (cum["clip"] * 1e-3 / args.steps) / 1e12
            line["roofline_flops"] = {"kernel": "clip_win_kernel", "achieved": tf, "unit": "TFLOP/s", "fp32_peak": fp32, "fp64_peak": fp64,
                                      "frac_fp32": tf / fp32, "frac_fp64": tf / fp64, "achieved_whole_clip_phase": tf_phase,
                                      "peak_source": "FMA microbenchmark on this GPU (b200cvt_measure_peaks), non-tensor",
                                      "algorithmic_flops_per_seed_iteration": {"lloyd": f_lloyd, "func_grad": f_newton}}
            # kNN kernel: read seed, write k indices + count, plus the bisector rows it now writes with the lists
            # (k rows of 6 doubles and 4 floats) — every rank: the seeds of its range and halo
            # units = the kNN queries one launch serves: all seeds at N = 1, the owned range + its two-cell halo at N > 1
            knn_queries = S if world == 1 else (h.stats().get("knn_queries") or own)
            knn_bytes = knn_queries * (dim * 8 + 20 * 4 + 4 + 20 * (6 * 8 + 4 * 4 if dim == 3 else 8 * 8 + 8 * 4)) * n_eval
            knn_gbs = knn_bytes / (cum["knn"] * 1e-3 / args.steps) / 1e9
            knn_traffic = None
            try:
                knn_traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("knn_kernel")
            except Exception:
                pass
            line["roofline_knn"] = {"bound": "hbm", "achieved": knn_gbs, "peak": hbm, "unit": "GB/s", "frac": knn_gbs / hbm,
                                    "peak_source": peak_src, "copy_gbs_here": copy, "traffic": knn_traffic,
                                    "units_per_launch": int(knn_queries),
                                    "note": "issue-bound (64-bit key compare-exchanges), not HBM; time = the kNN phase of bench.py's timers"}
            line["phase_ms_per_evaluation"] = {k: cum[k] / n_launch for k in ("sort", "knn", "pairs", "clip", "clip_kernel")}
            # outside the phases: the L-BFGS direction kernel (+ the push of the trial point to the peers), once per Newton iteration
            line["lbfgs_direction_ms_per_iteration"] = cum["cells"] / max(args.steps * evals["newton"]["iters"], 1)
            if by_rank is not None:
                line["phase_ms_per_evaluation_over_ranks"] = {
                    k: {"min": float(by_rank[:, i].min() / n_launch), "max": float(by_rank[:, i].max() / n_launch)}
                    for i, k in enumerate(("sort", "knn", "pairs", "clip"))}
                tot = by_rank[:, :4].sum(axis=1) / n_launch
                line["phase_ms_per_evaluation_over_ranks"]["all_phases"] = {"min": float(tot.min()), "max": float(tot.max())}
        except Exception as ex_:   # the bench value stands even if the roofline leg fails
            line["roofline"] = {"error": str(ex_)}
        # CPU baseline on the host cores, bounded sample of the same workload: rank 0 at N = 1 only
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import ref
                if ref.available():
                    r = ref.RefCVT(V, F, multithread=True)
                    r.set_points(X)
                    t = r.lloyd(1)                       # warm-up (thread partition of the mesh)
                    r.set_points(X)
                    c0 = r.counters()["funcgrad"]
                    t = r.lloyd(3) + r.newton(2, NEWTON_M)
                    ev = 3 + r.counters()["funcgrad"] - c0
                    line["cpu_baseline"] = {"value": S * ev / t, "unit": "seed-iterations/s", "cores": ref.RefCVT.nb_threads(),
                                            "kind": "reference", "sample": "3 Lloyd + Newton_iterations(2) = %d evaluations on the full input, %.1f s" % (ev, t)}
                    r.close()
                else:
                    from oracle import port
                    t0 = time.time()
                    port.lloyd(V, F, X, 2)
                    t = time.time() - t0
                    line["cpu_baseline"] = {"value": S * 2 / t, "unit": "seed-iterations/s", "cores": 1, "kind": "port",
                                            "sample": "2 Lloyd iterations on the full input, %.1f s" % t}
            except Exception as ex_:
                line["cpu_baseline"] = {"error": str(ex_)}
        emit(line)
    h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
