#!/bin/bash
# round 2, call AM (1 GPU): compute-sanitizer memcheck over every kernel family of the final build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py ) > gpurun_out/r2am_memcheck.log 2>&1
echo "memcheck rc $?"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2am_memcheck.log; tail -12 gpurun_out/r2am_memcheck.log | cut -c1-220
