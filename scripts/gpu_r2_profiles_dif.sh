# dif-only part of scripts/gpu_r2_profiles.sh (re-run after the last changes of the dif build kernel)
set -x
cd $GRAFT_REPO_ROOT
G=gpurun_out
python -m pytest tests/test_lm_dif_staged_gpu.py tests/test_frames_gpu.py tests/test_relink_gpu.py tests/test_lm_gpu.py -q -m gpu 2>&1 | tail -2 > $G/r2_final_gpu_tests_dif.txt
python bench.py --variant dif --steps 3 --warmup 3 --fits 65536 --no-extras > $G/r2_final_bench_dif.json 2> $G/r2_final_bench_dif.err
python bench.py --impl reference --variant dif --steps 3 --warmup 1 > $G/r2_final_bench_reference_dif.json 2>/dev/null
DIFP="python scripts/dif_staged_probe.py 65536"
$DIFP > $G/plain_dif.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dif_build_kernel -s 4 -c 1 -o $G/r2_prof_dif_build $DIFP > $G/ncu_dif.log 2>&1
$DIFP > $G/plain_dif2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $G/r2_final_launches_dif_staged.csv $DIFP > $G/ncu_dif2.log 2>&1
mkdir -p $G/export
rep=$G/r2_prof_dif_build.ncu-rep
ncu -i $rep --page details --csv > $G/export/r2_dif_build_details.csv 2>/dev/null
ncu -i $rep --page raw --csv > $G/export/r2_dif_build_raw.csv 2>/dev/null
python scripts/ncu_by_line.py $rep > $G/export/r2_dif_build_by_line.txt 2>/dev/null
rm -f $rep
tail -2 $G/plain_dif.log
