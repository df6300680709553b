#!/bin/bash
# round 2, call U: the whole GPU test suite (no -x) + the default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
echo "bench rc=$?" >> gpurun_out/r2u_bench.err
tail -25 gpurun_out/r2u_pytest.log | cut -c1-300
tail -3 gpurun_out/r2u_bench.err
