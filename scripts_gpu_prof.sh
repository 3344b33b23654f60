cd $GRAFT_REPO_ROOT
python scripts/ekfprof.py > gpurun_out/ekfprof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ekf_cov_dmma_kernel -s 1 -c 1 -o gpurun_out/prof_ekf_cov_dmma_r1 python scripts/ekfprof.py > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log | cut -c1-300
