mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/sc5_pytest.log 2>&1; tail -4 gpurun_out/sc5_pytest.log | cut -c1-300
python __graft_entry__.py smoke > gpurun_out/sc5_smoke.log 2>&1; tail -1 gpurun_out/sc5_smoke.log
timeout 600 python bench.py > gpurun_out/sc5_bench_n1.json 2> gpurun_out/sc5_bench_n1.err; tail -c 600 gpurun_out/sc5_bench_n1.json; tail -3 gpurun_out/sc5_bench_n1.err
for w in tkoz_test3_4096 csci6360_8192; do timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sc5_bench_$w.json 2> gpurun_out/sc5_bench_$w.err; tail -c 300 gpurun_out/sc5_bench_$w.json; done
export FFR_JIT_DUMP_DIR=/tmp/ffrjit FFR_JIT_NO_DISK_CACHE=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffr_jit_render -s 1 -c 1 -f -o gpurun_out/k1d_csci_sc python tools/prof_one.py csci 0 2 2048 2 > gpurun_out/k1d_ncu_csci_sc.log 2>&1; tail -1 gpurun_out/k1d_ncu_csci_sc.log
