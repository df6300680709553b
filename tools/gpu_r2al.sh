#!/bin/bash
# round 2, call AL (1 GPU): whole GPU suite + smoke + default bench + reference arm on the final build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2al_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2al_pytest.log
tail -6 gpurun_out/r2al_pytest.log | cut -c1-200
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/r2al_smoke.log 2>&1
tail -2 gpurun_out/r2al_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r2al_bench.json 2> gpurun_out/r2al_bench.err
echo "bench rc=$?"; tail -4 gpurun_out/r2al_bench.err
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2al_bench_ref.json 2> gpurun_out/r2al_bench_ref.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2al_bench.json') if l.startswith('{')][-1])
print("headline %.4e e2e %.4e"%(d['value'],d['e2e']['value']))
for c in d.get('configs',[]):
    print(c['config']['workload'], "%.4e %.4e frac %.2f"%(c['value'],c['e2e']['value'],c['atomic_roofline']['frac']))
r=json.loads([l for l in open('gpurun_out/r2al_bench_ref.json') if l.startswith('{')][-1])
print("reference %.4e"%r['value'])
P
