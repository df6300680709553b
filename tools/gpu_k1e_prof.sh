mkdir -p gpurun_out
export JIT=2 WAVES=4
echo "== global vs discard" > gpurun_out/k1e_probe.log
MODES=1,4 python tools/probe.py sierpinski barnsley sierp3d >> gpurun_out/k1e_probe.log 2>&1
for cfg in "192 4" "128 6" "384 2" "416 2" "512 1"; do set -- $cfg; echo "== TPB $1 MINB $2" >> gpurun_out/k1e_probe.log; FFR_JIT_TPB=$1 FFR_JIT_MINB=$2 MODES=1 python tools/probe.py sierpinski barnsley sierp3d >> gpurun_out/k1e_probe.log 2>&1; done
cat gpurun_out/k1e_probe.log
export FFR_JIT_DUMP_DIR=/tmp/ffrjit FFR_JIT_NO_DISK_CACHE=1
for w in barnsley sierp3d; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffr_jit_render -s 1 -c 1 -f -o gpurun_out/k1e_$w python tools/prof_one.py $w 0 2 8192 2 > gpurun_out/k1e_ncu_$w.log 2>&1; tail -2 gpurun_out/k1e_ncu_$w.log
done
