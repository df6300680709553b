#!/bin/bash
# round 2, call BG (2 GPUs): the driver's scaling command at N=2 on the final HEAD
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=2
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/r2bg_bench_n2.json 2> gpurun_out/r2bg_bench_n2.err
echo "bench rc=$?"; tail -3 gpurun_out/r2bg_bench_n2.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2bg_bench_n2.json') if l.startswith('{')][-1])
print("N=2 value %.4e e2e %.4e ms/step %.2f"%(d['value'],d['e2e']['value'],d['ms_per_step']))
s=d.get('strong_scaling')
if s: print("strong:", {k:s[k] for k in s if k in ('seconds','value','n_gpus')})
P
