set -x
cd $GRAFT_REPO_ROOT
make -s -C oracle 2>&1 | tail -2
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_der_v2.json 2> gpurun_out/bench_der_v2.err; tail -c 600 gpurun_out/bench_der_v2.err
python bench.py --variant dif --steps 3 --warmup 3 --fits 65536 --no-extras > gpurun_out/bench_dif_v2.json 2> gpurun_out/bench_dif_v2.err
SMALL="python bench.py --fits 32768 --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$SMALL > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_v2.csv $SMALL > gpurun_out/ncu1.log 2>&1
$SMALL > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lm_fit_kernel -s 3 -c 1 -o gpurun_out/prof_lm_der_r1_v2 $SMALL > gpurun_out/ncu2.log 2>&1
