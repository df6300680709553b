mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/launches_final_bench.log 2>&1; tail -c 400 gpurun_out/launches_final_bench.log; grep -c ffr_jit_render gpurun_out/launches_final.csv
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/fin5_bench.json 2>gpurun_out/fin5_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/fin5_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline'])"
