# usage: run_bench_n.sh <tag> <N> [extra bench args]  — weak-scaling bench line at N GPUs (torchrun for N > 1)
TAG=$1; N=$2; shift; shift
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_bench_n1.json.log 2> gpurun_out/${TAG}_bench_n1.err || tail -5 gpurun_out/${TAG}_bench_n1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/${TAG}_bench_n$N.json.log 2> gpurun_out/${TAG}_bench_n$N.err || tail -5 gpurun_out/${TAG}_bench_n$N.err
fi
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n$N.json.log").read().strip().splitlines()[-1])
print("N=$N value %.1f M  ms/step %.2f  e2e %.1f M  evals %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["run"]["evaluations_per_step"]))
print("  phases", {k: round(v, 4) for k, v in d.get("phase_ms_per_evaluation", {}).items()}, " per-eval total %.3f" % (d["ms_per_step"]/d["run"]["evaluations_per_step"]), " lbfgs direction ms/iter", round(d.get("lbfgs_direction_ms_per_iteration", 0), 4))
if d.get("c3"): print("  c3", {k: d["c3"].get(k) for k in ("ms_per_iteration", "lloyd_iterations_per_s", "seed_iterations_per_s", "error")})
PY
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n$N.json.log").read().strip().splitlines()[-1])
if d.get("phase_ms_per_evaluation_over_ranks"): print("  over ranks", {k: (round(v["min"], 3), round(v["max"], 3)) for k, v in d["phase_ms_per_evaluation_over_ranks"].items()})
PY
