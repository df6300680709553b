#!/bin/bash
# round 2, call AF (8 GPUs): the driver's scaling command on the final build (weak headline + cfg5 as stated),
# the reference arm under torchrun, and the multi-GPU tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=8
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/r2af_bench_n8.json 2> gpurun_out/r2af_bench_n8.err
echo "bench rc=$?"; tail -4 gpurun_out/r2af_bench_n8.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2af_bench_n8.json') if l.startswith('{')][-1])
print("N=8 value %.4e e2e %.4e ms/step %.2f"%(d['value'],d['e2e']['value'],d['ms_per_step']))
s=d.get('strong_scaling')
if s: print("strong:", {k:s[k] for k in s if k in ('seconds','samples_per_s','value','samples','n_gpus')})
P
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r2af_pytest.log 2>&1
tail -3 gpurun_out/r2af_pytest.log | cut -c1-200
