mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/f2_pytest.log 2>&1; tail -6 gpurun_out/f2_pytest.log | cut -c1-200
for w in sierpinski3d_512 sierpinski_1024; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f2_bench_$w.json 2> gpurun_out/f2_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/f2_bench_$w.json').read().strip().splitlines()[-1])
    print('$w', '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'], 'frac', round(d['roofline']['frac'],3), {k: '%.3g'%v for k,v in d['atomic_roofline'].items()})
except Exception as e:
    print('$w', 'FAILED', e); print(open('gpurun_out/f2_bench_$w.err').read()[-2000:])
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_k1e_sierp3d.csv python bench.py --workload sierpinski3d_512 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f2_ncu_bench.log 2>&1
