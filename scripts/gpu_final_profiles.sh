# Round-end evidence run (under gpurun): tests, bench lines, ncu launch lists and --set full captures.
# Every ncu run is preceded by the same command line exiting 0 without ncu.
set -x
cd $GRAFT_REPO_ROOT
make -s -C oracle 2>&1 | tail -2
python -m pytest tests -q -m gpu 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_der.json 2> gpurun_out/final_bench_der.err
python bench.py --variant dif --steps 3 --warmup 3 --fits 65536 --no-extras > gpurun_out/final_bench_dif.json 2> gpurun_out/final_bench_dif.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference_der.json 2>/dev/null
python bench.py --impl reference --variant dif --steps 3 --warmup 1 > gpurun_out/final_bench_reference_dif.json 2>/dev/null
# staged shape (der, large batches): launch list of the bench command, then one full capture per kernel
BIG="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$BIG > gpurun_out/plain0.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_staged_der.csv $BIG > gpurun_out/ncu_s0.log 2>&1
for K in jac solve eval; do
  $BIG > gpurun_out/plain_$K.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:staged_${K}_kernel -s 2 -c 1 -o gpurun_out/final_prof_staged_$K $BIG > gpurun_out/ncu_s_$K.log 2>&1
done
# per-contour m = 8 fits, a thread per fit
I8="python scripts/image8_probe.py 1048576 thread"
$I8 > gpurun_out/plain_i8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:image8_der_kernel -s 2 -c 1 -o gpurun_out/final_prof_image8 $I8 > gpurun_out/ncu_i8.log 2>&1
I8D="python scripts/image8_probe.py 1048576 thread dif"
$I8D > gpurun_out/plain_i8d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:image8_dif_kernel -s 2 -c 1 -o gpurun_out/final_prof_image8_dif $I8D > gpurun_out/ncu_i8d.log 2>&1
# monolithic shape (a warp per fit): der at 4096 fits (below the switch-over), dif
SMALL="python bench.py --fits 4096 --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$SMALL > gpurun_out/plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/final_launches_lm_der.csv $SMALL > gpurun_out/ncu_a.log 2>&1
$SMALL > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lm_fit_kernel -s 3 -c 1 -o gpurun_out/final_prof_lm_der $SMALL > gpurun_out/ncu_b.log 2>&1
SMALLD="python bench.py --variant dif --fits 8192 --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$SMALLD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lm_fit_kernel -s 3 -c 1 -o gpurun_out/final_prof_lm_dif $SMALLD > gpurun_out/ncu_c.log 2>&1
python scripts/ekfprof.py > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ekf_cov_pipe_kernel -s 1 -c 1 -o gpurun_out/final_prof_ekf_cov python scripts/ekfprof.py > gpurun_out/ncu_d.log 2>&1
python scripts/ekfprof.py > gpurun_out/plain5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/final_launches_ekf.csv python scripts/ekfprof.py > gpurun_out/ncu_e.log 2>&1
# gpurun merges at most 64 MiB back: turn the reports into their text summaries here and keep only the jac report
mkdir -p gpurun_out/export
for K in staged_jac staged_solve staged_eval image8 image8_dif lm_der lm_dif ekf_cov; do
  rep=gpurun_out/final_prof_$K.ncu-rep
  [ -f $rep ] || continue
  ncu -i $rep --page details --csv > gpurun_out/export/${K}_details.csv 2>/dev/null
  [ $K != ekf_cov ] && python scripts/ncu_by_line.py $rep > gpurun_out/export/${K}_by_line.txt 2>/dev/null
  [ $K != staged_jac ] && rm -f $rep
done
ls -la gpurun_out gpurun_out/export | tail -50
