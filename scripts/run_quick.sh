# quick GPU check: bit-level fingerprints against the round-1 kernels + a short bench line.  usage: run_quick.sh <tag>
TAG=$1
python scripts/gpu_hash.py > gpurun_out/hash_$TAG.json 2> gpurun_out/hash_$TAG.err || tail -5 gpurun_out/hash_$TAG.err
python scripts/hash_diff.py profiles/r2_hashes_baseline.json gpurun_out/hash_$TAG.json
python bench.py --steps 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json.log 2> gpurun_out/${TAG}_bench.err || tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json.log").read().strip().splitlines()[-1])
print("value %.1f M  ms/step %.2f  e2e %.1f M" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6))
print("phases", {k: round(v, 4) for k, v in d["phase_ms_per_evaluation"].items()})
print("frac_fp32 %.4f" % d["roofline_flops"]["frac_fp32"])
PY
