timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_k1d_n2.json 2> gpurun_out/bench_k1d_n2.err; tail -c 900 gpurun_out/bench_k1d_n2.json; tail -2 gpurun_out/bench_k1d_n2.err
