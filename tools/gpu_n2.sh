mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x ) > gpurun_out/n2_pytest.log 2>&1; tail -3 gpurun_out/n2_pytest.log
for w in csci6360_4096 sierpinski_1024; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --workload $w > gpurun_out/n2_bench_$w.json 2> gpurun_out/n2_bench_$w.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/n2_bench_$w.json').read().strip().splitlines()[-1])
    print('$w N=2', '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'])
except Exception as e:
    print('$w', 'FAILED', e); print(open('gpurun_out/n2_bench_$w.err').read()[-2500:])
PY
done
