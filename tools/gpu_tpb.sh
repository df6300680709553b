mkdir -p gpurun_out
export JIT=2 WAVES=2 MODES=1 FFR_JIT_NO_DISK_CACHE=1
for t in 320 352 384; do echo "== TPB $t"; FFR_JIT_TPB=$t python tools/probe.py csci tkoz3 2>&1 | cut -c1-200; done > gpurun_out/tpb_probe.log 2>&1
echo "== ROT_STATIC" >> gpurun_out/tpb_probe.log; FFR_JIT_ROT_STATIC=1 python tools/probe.py csci 2>&1 | cut -c1-200 >> gpurun_out/tpb_probe.log
cat gpurun_out/tpb_probe.log
