# round 2, call A: sanity of the round-1 state + ADVICE fixes on a fresh box, compute-sanitizer passes
set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_pytest.log
timeout 600 compute-sanitizer --tool memcheck --leak-check no --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r2a_memcheck_smoke.log 2>&1; tail -5 gpurun_out/r2a_memcheck_smoke.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r2a_racecheck_smoke.log 2>&1; tail -5 gpurun_out/r2a_racecheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --leak-check no --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "sharded_lloyd_two_partitions or sharded_volumetric" > gpurun_out/r2a_memcheck_sharded.log 2>&1; tail -5 gpurun_out/r2a_memcheck_sharded.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "sharded_lloyd_two_partitions" > gpurun_out/r2a_racecheck_sharded.log 2>&1; tail -5 gpurun_out/r2a_racecheck_sharded.log
python bench.py --steps 3 > gpurun_out/r2a_bench.json.log 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json.log
