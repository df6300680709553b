mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/s_pytest.log 2>&1; tail -5 gpurun_out/s_pytest.log
python __graft_entry__.py smoke > gpurun_out/s_smoke.log 2>&1; tail -2 gpurun_out/s_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/s_bench_n1.json 2> gpurun_out/s_bench_n1.err; tail -c 1500 gpurun_out/s_bench_n1.json; tail -3 gpurun_out/s_bench_n1.err
for w in sierpinski_1024 barnsley_2048 tkoz_test3_4096 sierpinski3d_512; do timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s_bench_$w.json 2> gpurun_out/s_bench_$w.err; done
