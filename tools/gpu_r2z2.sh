#!/bin/bash
# round 2, call Z2 (1 GPU): the whole GPU suite on the final build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2z2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2z2_pytest.log
tail -8 gpurun_out/r2z2_pytest.log | cut -c1-200
