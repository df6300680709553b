#!/bin/bash
# round 2, call AY (1 GPU): the other variation flames of the examples under the K1d switches (is the static choice sensible beyond the two BASELINE flames?)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 FFR_JIT_NO_DISK_CACHE=1
{ echo "== default"; python tools/probe.py tkoz1 tkoz2 tkoz4 tkoz5 swv flam3 2>&1 | cut -c1-100
echo "== GEN 2"; FFR_JIT_GEN_ROLLED=2 python tools/probe.py tkoz1 tkoz2 tkoz4 tkoz5 swv flam3 2>&1 | cut -c1-100
echo "== ANG 0"; FFR_JIT_POLAR_ANG=0 python tools/probe.py tkoz1 tkoz2 tkoz4 tkoz5 swv flam3 2>&1 | cut -c1-100
echo "== P 0 S 0 COLD 0 (start of session)"; FFR_JIT_POLAR_NEED=0 FFR_JIT_SIN_VIA_SINCOS=0 FFR_JIT_COLD=0 python tools/probe.py tkoz1 tkoz2 tkoz4 tkoz5 swv flam3 2>&1 | cut -c1-100; } | tee gpurun_out/r2ay_probe.log
