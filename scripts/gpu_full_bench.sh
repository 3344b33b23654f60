cd $GRAFT_REPO_ROOT
make -s -C oracle 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_der_v3.json 2> gpurun_out/bench_der_v3.err; tail -c 400 gpurun_out/bench_der_v3.err
python bench.py --variant dif --steps 3 --warmup 3 --fits 65536 --no-extras > gpurun_out/bench_dif_v3.json 2> gpurun_out/bench_dif_v3.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_der.json 2>/dev/null
python bench.py --impl reference --variant dif --steps 3 --warmup 1 > gpurun_out/bench_ref_dif.json 2>/dev/null
