#!/bin/bash
# round 2, call Z (1 GPU): the whole GPU suite + smoke + the default bench on the final build, and
# an ncu pass over the scatter-ceiling microbenchmarks (L2 RED sector counts, DRAM bytes)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r2z_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
tail -6 gpurun_out/r2z_pytest.log | cut -c1-200
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/r2z_smoke.log 2>&1
tail -2 gpurun_out/r2z_smoke.log
M=gpu__time_duration.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_read.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_bank_conflicts_pipe_lsu.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed
for wl in barnsley sierp3d csci; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:atomic_ --csv --log-file gpurun_out/r2z_ceil_$wl.csv python tools/prof_ceilings.py $wl > gpurun_out/r2z_ceil_$wl.log 2>&1
  tail -3 gpurun_out/r2z_ceil_$wl.log
done
( time timeout 900 python bench.py ) > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
echo "bench rc=$?"; tail -4 gpurun_out/r2z_bench.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2z_bench.json') if l.startswith('{')][-1])
print("headline %.4e e2e %.4e"%(d['value'],d['e2e']['value']))
for c in d.get('configs',[]):
    print(c['workload'] if 'workload' in c else c['config']['workload'], "%.4e %.4e frac %.2f"%(c['value'],c['e2e']['value'],c['atomic_roofline']['frac']))
P
