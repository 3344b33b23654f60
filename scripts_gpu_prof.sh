cd $GRAFT_REPO_ROOT
SMALL="python bench.py --fits 16384 --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$SMALL > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lm_fit_kernel -s 3 -c 1 -o gpurun_out/prof_lm_der_v2 $SMALL > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log | cut -c1-200
