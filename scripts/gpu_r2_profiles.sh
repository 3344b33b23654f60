# Round-2 evidence run (under gpurun): tests, bench lines, ncu launch lists and --set full captures.
# Every ncu run is preceded by the same command line exiting 0 without ncu; numbers taken under ncu are never bench values.
set -x
cd $GRAFT_REPO_ROOT
G=gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3 > $G/r2_final_gpu_tests.txt
python bench.py --steps 5 --warmup 3 > $G/r2_final_bench_der.json 2> $G/r2_final_bench_der.err
python bench.py --variant dif --steps 3 --warmup 3 --fits 65536 --no-extras > $G/r2_final_bench_dif.json 2> $G/r2_final_bench_dif.err
python bench.py --impl reference --steps 3 --warmup 1 > $G/r2_final_bench_reference_der.json 2>/dev/null
python bench.py --impl reference --variant dif --steps 3 --warmup 1 > $G/r2_final_bench_reference_dif.json 2>/dev/null
BIG="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$BIG > $G/plain0.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $G/r2_final_launches_bench_der.csv $BIG > $G/ncu_s0.log 2>&1
for K in jac solve eval; do
  $BIG > $G/plain_$K.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:staged_${K}_kernel -s 2 -c 1 -o $G/r2_prof_staged_$K $BIG > $G/ncu_s_$K.log 2>&1
done
DIFP="python scripts/dif_staged_probe.py 65536"
$DIFP > $G/plain_dif.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dif_build_kernel -s 4 -c 1 -o $G/r2_prof_dif_build $DIFP > $G/ncu_dif.log 2>&1
$DIFP > $G/plain_dif2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $G/r2_final_launches_dif_staged.csv $DIFP > $G/ncu_dif2.log 2>&1
python scripts/ekfprof.py > $G/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ekf_cov_tma_kernel -s 1 -c 1 -o $G/r2_prof_ekf_cov python scripts/ekfprof.py > $G/ncu_d.log 2>&1
python scripts/ekfprof.py > $G/plain5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $G/r2_final_launches_ekf.csv python scripts/ekfprof.py > $G/ncu_e.log 2>&1
mkdir -p $G/export
for K in staged_jac staged_solve staged_eval dif_build ekf_cov; do
  rep=$G/r2_prof_$K.ncu-rep
  [ -f $rep ] || continue
  ncu -i $rep --page details --csv > $G/export/r2_${K}_details.csv 2>/dev/null
  ncu -i $rep --page raw --csv > $G/export/r2_${K}_raw.csv 2>/dev/null
  [ $K != ekf_cov ] && python scripts/ncu_by_line.py $rep > $G/export/r2_${K}_by_line.txt 2>/dev/null
  rm -f $rep
done
ls -la $G/export | tail -30
