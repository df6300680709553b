#!/bin/bash
# round 2, call I: probe after the address laundering, then the full bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export WAVES=4 JIT=2
timeout 300 python tools/probe.py csci tkoz3 > gpurun_out/r2i_probe.log 2>&1
timeout 300 python tools/probe.py csci tkoz3 >> gpurun_out/r2i_probe.log 2>&1
cat gpurun_out/r2i_probe.log
unset WAVES JIT
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
tail -4 gpurun_out/r2i_bench.err
