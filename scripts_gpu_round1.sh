set -x
cd $GRAFT_REPO_ROOT
make -s -C oracle 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_der.json 2> gpurun_out/bench_der.err; tail -c 2500 gpurun_out/bench_der.json; tail -5 gpurun_out/bench_der.err
python bench.py --variant dif --steps 3 --warmup 3 --fits 65536 > gpurun_out/bench_dif.json 2> gpurun_out/bench_dif.err; tail -c 2500 gpurun_out/bench_dif.json; tail -5 gpurun_out/bench_dif.err
