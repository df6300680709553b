mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n1.json; cat gpurun_out/bench_n1.json | cut -c1-1500
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n2.json; cat gpurun_out/bench_n2.json | cut -c1-900
