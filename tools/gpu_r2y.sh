#!/bin/bash
# round 2, call Y (2 GPUs): multi-GPU tests incl. the torchrun NCCL reduce, bench of the colour flame at N=2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -k "multi or atan2 or torchrun or two_devices or cli or streaming" ) > gpurun_out/r2y_pytest.log 2>&1
tail -12 gpurun_out/r2y_pytest.log | cut -c1-250
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 --workload tkoz_test3_4096 ) > gpurun_out/r2y_bench_n2_tkoz3.json 2> gpurun_out/r2y_bench_n2_tkoz3.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2y_bench_n2_tkoz3.json') if l.startswith('{')][-1])
print("tkoz3 N=2 value %.4e e2e %.4e ms/step %.2f"%(d['value'],d['e2e']['value'],d['ms_per_step']))
P
tail -3 gpurun_out/r2y_bench_n2_tkoz3.err
