set -x
python bench.py > gpurun_out/r1g_bench.json.log 2> gpurun_out/r1g_bench.err
ncu --set full --clock-control none --import-source on -k regex:'clip_win_kernel|knn_kernel|facet_home|facet_task|reduce_pairs|compact_pairs' -s 18 -c 6 -o gpurun_out/r1g_prof python scripts/gpu_prof.py 316 200000 > gpurun_out/r1g_prof.log 2>&1
ncu --set full --clock-control none -k regex:'lbfgs_direction|lbfgs_post' -s 6 -c 2 -o gpurun_out/r1g_prof_lbfgs python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_prof_lbfgs.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_launches.log 2>&1
tail -c 600 gpurun_out/r1g_bench.json.log
