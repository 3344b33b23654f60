# Round 2, second half: evidence run after the EKF work (the LM kernels are those of scripts/gpu_r2_profiles.sh).
# Every ncu run is preceded by the same command line exiting 0 without ncu; numbers taken under ncu are never bench values.
set -x
cd $GRAFT_REPO_ROOT
G=gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3 > $G/r2b_final_gpu_tests.txt
python bench.py --steps 5 --warmup 3 > $G/r2b_final_bench_der.json 2> $G/r2b_final_bench_der.err
python bench.py --impl reference --steps 3 --warmup 1 > $G/r2b_final_bench_reference_der.json 2>/dev/null
python scripts/ekf_bench_probe.py > $G/r2b_final_ekf.json 2> $G/r2b_final_ekf.err
python scripts/ekfprof.py > $G/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ekf_cov_pp_kernel -s 1 -c 1 -o $G/r2b_prof_ekf_cov python scripts/ekfprof.py > $G/ncu_d.log 2>&1
for cfg in "1000 0" "5000 0" "20000 0" "5000 10"; do
  tag=$(echo $cfg | tr ' ' '_')
  python scripts/ekfprof.py $cfg > $G/plain_ekf_$tag.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $G/r2b_final_launches_ekf_$tag.csv python scripts/ekfprof.py $cfg > $G/ncu_e_$tag.log 2>&1
done
mkdir -p $G/export
rep=$G/r2b_prof_ekf_cov.ncu-rep
ncu -i $rep --page details --csv > $G/export/r2b_ekf_cov_details.csv 2>/dev/null
ncu -i $rep --page raw --csv > $G/export/r2b_ekf_cov_raw.csv 2>/dev/null
python scripts/ncu_by_line.py $rep > $G/export/r2b_ekf_cov_by_line.txt 2>/dev/null
tail -2 $G/r2b_final_gpu_tests.txt
head -c 600 $G/r2b_final_bench_der.json
