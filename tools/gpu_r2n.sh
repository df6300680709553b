#!/bin/bash
# round 2, call N (8 GPUs): BASELINE config 5 through the CLI (one process per GPU vs one process),
# then the torchrun bench at N=8 (weak headline + the fixed-size cfg5 run)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2n_smi.txt 2>&1; nproc >> gpurun_out/r2n_smi.txt
python - > /tmp/csci8k.json <<'P'
import importlib,sys
sys.path.insert(0,'.')
ex=importlib.import_module("flame-fractal-renderer_b200.examples")
print(ex.example_json("csci6360_project", size=[8192,8192]))
P
B=flame-fractal-renderer_b200/ffr-buf.out
export FFR_TIMING=1
run() {
  name=$1; shift
  { time env "$@" $B -f /tmp/csci8k.json -o /tmp/out_$name.buf -s 100000000000 -b 8192 --seed 3 --gpus 8 2> gpurun_out/r2n_cli_$name.txt ; } 2> gpurun_out/r2n_time_$name.txt
  echo "== $name: $(grep real gpurun_out/r2n_time_$name.txt)"
  tr '\r' '\n' < gpurun_out/r2n_cli_$name.txt | grep -E "^timing: |render done|ERROR|samples plotted"
  tr '\r' '\n' < gpurun_out/r2n_cli_$name.txt | grep -E "^timing\[" | sort -t' ' -k3 | tail -4
}
run mp_cold X=1
run mp_warm X=1
run mp_warm2 X=1
run sp FFR_SINGLE_PROCESS=1
cmp /tmp/out_sp.buf /tmp/out_mp_warm.buf && echo "8-GPU outputs identical (single process vs one process per GPU)"
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r2n_bench_n8.json 2> gpurun_out/r2n_bench_n8.err
tail -3 gpurun_out/r2n_bench_n8.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2n_bench_n8.json') if l.startswith('{')][-1])
print("N=8 value %.4e e2e %.4e ms/step %.2f"%(d['value'],d['e2e']['value'],d['ms_per_step']))
print("strong:", d['strong_scaling']['value'], d['strong_scaling']['seconds'], d['strong_scaling']['wall_seconds'])
P
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 6 --warmup 3 --workload tkoz_test3_4096 ) > gpurun_out/r2n_bench_n8_tkoz3.json 2> gpurun_out/r2n_bench_n8_tkoz3.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2n_bench_n8_tkoz3.json') if l.startswith('{')][-1])
print("tkoz3 N=8 value %.4e e2e %.4e ms/step %.2f"%(d['value'],d['e2e']['value'],d['ms_per_step']))
P
