mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_pytest.log 2>&1; tail -4 gpurun_out/f_pytest.log
for w in sierpinski_1024 barnsley_2048 sierpinski3d_512 tkoz_test3_4096; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_$w.json 2> gpurun_out/f_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/f_bench_$w.json').read().strip().splitlines()[-1])
    print('$w', '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'], 'frac', round(d['roofline']['frac'],3), {k: '%.3g'%v for k,v in d['atomic_roofline'].items()})
except Exception as e:
    print('$w', 'FAILED', e); print(open('gpurun_out/f_bench_$w.err').read()[-2000:])
PY
done
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; tail -c 600 gpurun_out/f_bench_n1.json
