FFR_JIT_DUMP_DIR=$PWD/gpurun_out/jitsrc_tkoz3 timeout 600 ncu --set full --import-source on --clock-control none -k regex:ffr_jit_render -s 1 -c 1 -f -o gpurun_out/r1_k1d_tkoz3 python tools/prof_one.py tkoz3 0 1 2048 2 > gpurun_out/ncu_k1d_tkoz3.log 2>&1
tail -2 gpurun_out/ncu_k1d_tkoz3.log
FFR_JIT_DUMP_DIR=$PWD/gpurun_out/jitsrc_async timeout 600 ncu --set full --import-source on --clock-control none -k regex:ffr_jit_render -s 1 -c 1 -f -o gpurun_out/r1_k1d_csci python tools/prof_one.py csci 0 1 2048 2 > gpurun_out/ncu_k1d.log 2>&1
tail -2 gpurun_out/ncu_k1d.log
timeout 600 python bench.py --steps 3 --warmup 3 --workload csci6360_8192 --no-cpu-baseline > gpurun_out/bench_k1d_csci8192.json 2>&1; tail -c 600 gpurun_out/bench_k1d_csci8192.json
