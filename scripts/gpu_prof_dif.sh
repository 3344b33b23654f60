cd $GRAFT_REPO_ROOT
SMALL="python bench.py --variant dif --fits 8192 --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$SMALL > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lm_fit_kernel -s 3 -c 1 -o gpurun_out/prof_lm_dif_r1 $SMALL > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/ncu2.log | cut -c1-100
