#!/bin/bash
# round 2, call AG (1 GPU): the float/u32 build (num_t = float, hist_t = uint32_t) through the JIT kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export JIT=2 WAVES=4 MODES=1 ELEM=4
python tools/probe.py csci tkoz3 sierpinski barnsley sierp3d 2>&1 | cut -c1-260 | tee gpurun_out/r2ag_probe.log
JIT=1 python tools/probe.py csci sierpinski 2>&1 | cut -c1-200 | tee -a gpurun_out/r2ag_probe.log
