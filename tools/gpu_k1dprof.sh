mkdir -p gpurun_out
export FFR_JIT_DUMP_DIR=/tmp/ffrjit FFR_JIT_NO_DISK_CACHE=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffr_jit_render -s 1 -c 1 -f -o gpurun_out/k1d_csci python tools/prof_one.py csci 0 2 2048 2 > gpurun_out/k1d_ncu_csci.log 2>&1; tail -1 gpurun_out/k1d_ncu_csci.log
