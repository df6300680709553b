mkdir -p gpurun_out
N=${1:-8}
for w in csci6360_4096 barnsley_2048; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --workload $w > gpurun_out/n${N}_bench_$w.json 2> gpurun_out/n${N}_bench_$w.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/n${N}_bench_$w.json').read().strip().splitlines()[-1])
    print('$w N=$N', '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'], 'ms/step', d['ms_per_step'])
except Exception as e:
    print('$w', 'FAILED', e); print(open('gpurun_out/n${N}_bench_$w.err').read()[-2500:])
PY
done
