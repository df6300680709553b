mkdir -p gpurun_out
export FFR_JIT_DUMP_DIR=/tmp/ffrjit FFR_JIT_NO_DISK_CACHE=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffr_jit_render -s 1 -c 1 -f -o gpurun_out/k1e_sierp3d_dir python tools/prof_one.py sierp3d 0 2 8192 2 > gpurun_out/k1e_ncu_sierp3d_dir.log 2>&1; tail -1 gpurun_out/k1e_ncu_sierp3d_dir.log
