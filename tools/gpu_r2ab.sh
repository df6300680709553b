#!/bin/bash
# round 2, call AB (1 GPU): bench lines of the 384-thread K1d against the same build forced to 320 threads
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for wl in csci6360_4096 tkoz_test3_4096; do
  for t in 384 320; do
    FFR_JIT_TPB=$t FFR_JIT_NO_DISK_CACHE=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-strong --no-configs --workload $wl > gpurun_out/r2ab_${wl}_$t.json 2> gpurun_out/r2ab_${wl}_$t.err
    python - $wl $t <<'P'
import json,sys
d=json.loads([l for l in open('gpurun_out/r2ab_%s_%s.json'%(sys.argv[1],sys.argv[2])) if l.startswith('{')][-1])
print(sys.argv[1],sys.argv[2],"value %.4e e2e %.4e ms %.2f"%(d['value'],d['e2e']['value'],d['ms_per_step']))
P
  done
done
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
python tools/probe.py csci tkoz3 2>&1 | cut -c1-200
