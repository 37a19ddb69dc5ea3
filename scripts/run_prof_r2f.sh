# round 2, final state: the command batch behind profiles/r2f_* (one B200)
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json.log 2> gpurun_out/r2f_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2f_bench_reference.json.log 2> gpurun_out/r2f_bench_reference.err
ncu --set full --clock-control none --import-source on -k regex:'clip_win_kernel|knn_kernel|facet_home|facet_task|facet_big|reduce_pairs|compact_pairs' -s 21 -c 7 -o gpurun_out/r2f_prof python scripts/gpu_prof.py 316 200000 > gpurun_out/r2f_prof.log 2>&1
ncu --set full --clock-control none -k regex:'lbfgs_direction|lbfgs_post' -s 12 -c 2 -o gpurun_out/r2f_prof_lbfgs python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/r2f_prof_lbfgs.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/r2f_launches.log 2>&1
python scripts/gpu_configs.py c4 --ref > gpurun_out/r2f_configs.log 2>&1; cut -c1-900 gpurun_out/r2f_configs.log
python scripts/gpu_volume.py 47 119 > gpurun_out/r2f_volume.log 2>&1; cut -c1-700 gpurun_out/r2f_volume.log
python scripts/gpu_dropin_c2.py > gpurun_out/r2f_dropin_c2.log 2>&1; cut -c1-600 gpurun_out/r2f_dropin_c2.log
tail -c 600 gpurun_out/r2f_bench.json.log; tail -c 400 gpurun_out/r2f_bench_reference.json.log
