mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1_launches_bench.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/r1_render_regroup_csci4096 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --waves 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/r1_render_affine_sierp3d python bench.py --workload sierpinski3d_512 --steps 1 --warmup 3 --no-cpu-baseline --waves 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/r1_render_direct_tkoz3 python bench.py --workload tkoz_test3_4096 --steps 1 --warmup 3 --no-cpu-baseline --waves 1 > /dev/null 2>&1
for w in sierpinski_1024 barnsley_2048 tkoz_test3_4096 sierpinski3d_512 csci6360_8192; do python bench.py --workload $w --steps 3 --warmup 3 --waves 4 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$w.json; done
