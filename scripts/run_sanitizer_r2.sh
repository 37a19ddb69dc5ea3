# round 2, end state: compute-sanitizer memcheck + racecheck over the kernels added or rewritten in round 2
set -x
timeout 900 compute-sanitizer --tool memcheck --leak-check no --print-limit 20 python scripts/gpu_san_r2.py > gpurun_out/r2s_memcheck.log 2>&1; tail -4 gpurun_out/r2s_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python scripts/gpu_san_r2.py > gpurun_out/r2s_racecheck.log 2>&1; tail -4 gpurun_out/r2s_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --leak-check no --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r2s_memcheck_smoke.log 2>&1; tail -3 gpurun_out/r2s_memcheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r2s_racecheck_smoke.log 2>&1; tail -3 gpurun_out/r2s_racecheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --leak-check no --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "sharded_lloyd_two_partitions or sharded_volumetric" > gpurun_out/r2s_memcheck_sharded.log 2>&1; tail -3 gpurun_out/r2s_memcheck_sharded.log
