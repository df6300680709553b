#!/bin/bash
# round 2, call T: evidence for profiles/ on the final build: 1e12-sample queue soak, launch list of
# the bench command, ncu --set full of the four kernels behind the bench lines
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time FFR_SOAK_SAMPLES=1e12 timeout 900 python -m pytest tests/test_gpu_jit.py -m gpu -q -k soak ) > gpurun_out/r2t_soak.log 2>&1
tail -6 gpurun_out/r2t_soak.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/r2t_launches_bench.log 2>&1
grep -c ffr_jit_render gpurun_out/r2t_launches_bench.csv
export FFR_JIT_DUMP_DIR=/tmp/ffrjit FFR_JIT_NO_DISK_CACHE=1
for cfg in csci tkoz3 sierp3d barnsley; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffr_jit_render -s 1 -c 1 -f -o gpurun_out/r2t_$cfg python tools/prof_one.py $cfg 0 2 2048 2 > gpurun_out/r2t_ncu_$cfg.log 2>&1; tail -1 gpurun_out/r2t_ncu_$cfg.log
done
