mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_jit.py tests/test_gpu_affine.py -m gpu -q -x ) > gpurun_out/n2b_pytest.log 2>&1; tail -3 gpurun_out/n2b_pytest.log
for w in csci6360_4096; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --workload $w > gpurun_out/n2b_bench_$w.json 2> gpurun_out/n2b_bench_$w.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/n2b_bench_$w.json').read().strip().splitlines()[-1])
    print('$w N=2', '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'])
except Exception as e:
    print('$w', 'FAILED', e); print(open('gpurun_out/n2b_bench_$w.err').read()[-2500:])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --impl reference --ref-samples 20000000 > gpurun_out/n2b_ref.json 2> gpurun_out/n2b_ref.err; tail -c 400 gpurun_out/n2b_ref.json; tail -2 gpurun_out/n2b_ref.err
export JIT=2 WAVES=2 MODES=1 FFR_JIT_NO_DISK_CACHE=1
for i in 1 2 3; do python tools/probe.py csci 2>&1 | cut -c1-100; FFR_JIT_ROT_STATIC=1 python tools/probe.py csci 2>&1 | cut -c1-100; done > gpurun_out/n2b_rot.log 2>&1; cat gpurun_out/n2b_rot.log
