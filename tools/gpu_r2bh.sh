#!/bin/bash
# round 2, call BH (1 GPU): ncu launch list of the bench command on the final HEAD
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2bh_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong > gpurun_out/r2bh_launches_bench.log 2>&1
grep -c ffr_jit_render gpurun_out/r2bh_launches_bench.csv
python - <<'P'
import csv,collections,io
rows=[l for l in open('gpurun_out/r2bh_launches_bench.csv') if l.startswith('"')]
rd=csv.DictReader(io.StringIO(''.join(rows)))
t=collections.OrderedDict()
for r in rd:
    k=r['Kernel Name'].split('(')[0][:40]
    t.setdefault(k,[]).append(float(r['Metric Value'].replace(',',''))/1e6)
for k,v in t.items(): print('%-42s n=%3d  mean %.3f ms  total %.1f ms'%(k,len(v),sum(v)/len(v),sum(v)))
P
