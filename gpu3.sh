timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_k1d_n1.json 2> gpurun_out/bench_k1d_n1.err; tail -c 3000 gpurun_out/bench_k1d_n1.json; tail -3 gpurun_out/bench_k1d_n1.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload tkoz_test3_4096 --no-cpu-baseline > gpurun_out/bench_k1d_tkoz3.json 2>&1; tail -c 1500 gpurun_out/bench_k1d_tkoz3.json
