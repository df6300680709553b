#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
tail -4 gpurun_out/r2v_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2v_bench_ref.json 2> gpurun_out/r2v_bench_ref.err
tail -c 600 gpurun_out/r2v_bench_ref.json
