#!/bin/bash
# round 2, call M (2 GPUs): where does the time to a ready context go? one process vs one per GPU
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python - > /tmp/csci8k.json <<'P'
import importlib,sys
sys.path.insert(0,'.')
ex=importlib.import_module("flame-fractal-renderer_b200.examples")
print(ex.example_json("csci6360_project", size=[8192,8192]))
P
B=flame-fractal-renderer_b200/ffr-buf.out
export FFR_TIMING=1
run() { # name, extra env..., then args
  name=$1; shift
  { time env "$@" $B -f /tmp/csci8k.json -o /tmp/out_$name.buf -s 25000000000 -b 8192 --jit --seed 3 $GP 2> gpurun_out/r2m_cli_$name.txt ; } 2> gpurun_out/r2m_time_$name.txt
  echo "== $name: $(grep real gpurun_out/r2m_time_$name.txt)"
  tr '\r' '\n' < gpurun_out/r2m_cli_$name.txt | grep -E "^timing|render done|ERROR" 
}
GP="--gpus 1" run one_cold X=1
GP="--gpus 1" run one_warm X=1
GP="--gpus 2" run mp X=1
GP="--gpus 2" run mp_mask FFR_WORKER_MASK=1
GP="--gpus 2" run sp FFR_SINGLE_PROCESS=1
GP="--gpus 2" run mp_eager CUDA_MODULE_LOADING=EAGER
cmp /tmp/out_sp.buf /tmp/out_mp.buf && cmp /tmp/out_mp.buf /tmp/out_mp_mask.buf && echo "outputs identical"
