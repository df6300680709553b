# K1e (pure-affine flame-specialised kernel): GPU tests + bench against K1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_affine.py tests/test_gpu_jit.py -m gpu -x -q ) > gpurun_out/k1e_pytest.log 2>&1; tail -15 gpurun_out/k1e_pytest.log
for w in sierpinski_1024 barnsley_2048 sierpinski3d_512; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --jit 2 > gpurun_out/k1e_bench_$w.json 2> gpurun_out/k1e_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/k1e_bench_$w.json').read().strip().splitlines()[-1])
    print('$w', '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], d['atomic_roofline'], d['config']['kernel'])
except Exception as e:
    print('$w', 'FAILED', e); print(open('gpurun_out/k1e_bench_$w.err').read()[-2000:])
PY
done
