mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/fin_pytest.log 2>&1; tail -4 gpurun_out/fin_pytest.log | cut -c1-200
python __graft_entry__.py smoke > gpurun_out/fin_smoke.log 2>&1; tail -2 gpurun_out/fin_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/fin_bench_n1.json 2> gpurun_out/fin_bench_n1.err; tail -c 2500 gpurun_out/fin_bench_n1.json; tail -4 gpurun_out/fin_bench_n1.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/fin_bench_ref.json 2> gpurun_out/fin_bench_ref.err; tail -c 900 gpurun_out/fin_bench_ref.json; tail -4 gpurun_out/fin_bench_ref.err
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py ) > gpurun_out/fin_memcheck.log 2>&1; echo "memcheck rc $?"; tail -25 gpurun_out/fin_memcheck.log | cut -c1-220
export JIT=2 WAVES=2 MODES=1
for i in 1 2 3; do python tools/probe.py csci 2>&1 | cut -c1-120; FFR_JIT_ROT_STATIC=1 python tools/probe.py csci 2>&1 | cut -c1-120; done > gpurun_out/fin_rot.log 2>&1; cat gpurun_out/fin_rot.log
