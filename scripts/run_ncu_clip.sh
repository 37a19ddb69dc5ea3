# ncu --set full of the clip kernel (and neighbours) on the C2 input, 4th Lloyd evaluation.  usage: run_ncu_clip.sh <tag> [kernel regex]
TAG=$1
RE=${2:-clip_win_kernel}
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s 3 -c 2 -o gpurun_out/${TAG}_prof python scripts/gpu_prof.py 316 200000 > gpurun_out/${TAG}_prof.log 2>&1
tail -3 gpurun_out/${TAG}_prof.log
