#!/bin/bash
# round 2, call K (2 GPUs): multi-GPU tests, CLI start-up timing (one process per GPU vs one process),
# torchrun bench at N=2, e2e check of the streaming interface on the 1 GiB config
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2k_smi.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=8 ) > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -14 gpurun_out/r2k_pytest.log
python - > /tmp/csci8k.json <<'P'
import importlib,sys
sys.path.insert(0,'.')
ex=importlib.import_module("flame-fractal-renderer_b200.examples")
print(ex.example_json("csci6360_project", size=[8192,8192]))
P
B=flame-fractal-renderer_b200/ffr-buf.out
for mode in warmcache_mp sp mp; do
  export FFR_TIMING=1
  if [ $mode = sp ]; then export FFR_SINGLE_PROCESS=1; else unset FFR_SINGLE_PROCESS; fi
  { time $B -f /tmp/csci8k.json -o /tmp/out_$mode.buf -s 25000000000 -b 8192 --gpus 2 --jit --seed 3 2> gpurun_out/r2k_cli_$mode.txt ; } 2> gpurun_out/r2k_time_$mode.txt; cat gpurun_out/r2k_time_$mode.txt
  grep -E "timing|wall|render done|samples plotted" gpurun_out/r2k_cli_$mode.txt | tr '\r' '\n' | grep -v progress
done
cmp /tmp/out_sp.buf /tmp/out_mp.buf && echo "2-GPU outputs identical (single process vs one process per GPU)"
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err
tail -c 900 gpurun_out/r2k_bench_n2.json; tail -3 gpurun_out/r2k_bench_n2.err
( timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --workload sierpinski3d_512 --no-cpu-baseline ) > gpurun_out/r2k_bench_s3d.json 2> gpurun_out/r2k_bench_s3d.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2k_bench_s3d.json') if l.startswith('{')][-1])
print("sierpinski3d value %.3e e2e %.3e ratio %.2f"%(d['value'],d['e2e']['value'],d['e2e']['value']/d['value']))
P
