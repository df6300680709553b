mkdir -p gpurun_out
export FFR_JIT_NO_DISK_CACHE=1 FFR_ACC_MAX_MB=2048
for g in 0 2; do
FFR_ACC_GRAN=$g timeout 600 ncu --set full --clock-control none -k regex:ffr_jit_render -s 1 -c 1 -f -o gpurun_out/k1e_sierp3d_scr$g python tools/prof_one.py sierp3d 0 2 8192 2 > gpurun_out/k1e_ncu_sierp3d_scr$g.log 2>&1; tail -1 gpurun_out/k1e_ncu_sierp3d_scr$g.log
done
# barnsley with the default policy (cell scramble at 32 MiB), for profiles/
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffr_jit_render -s 1 -c 1 -f -o gpurun_out/k1e_barnsley_scr python tools/prof_one.py barnsley 0 2 8192 2 > gpurun_out/k1e_ncu_barnsley_scr.log 2>&1
