#!/bin/bash
# round 2, call O (8 GPUs): how long does driver initialisation take with one visible GPU vs eight,
# alone and with eight processes at once?
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,persistence_mode --format=csv > gpurun_out/r2o_smi.txt 2>&1
python - > /tmp/f.json <<'P'
import importlib,sys
sys.path.insert(0,'.')
ex=importlib.import_module("flame-fractal-renderer_b200.examples")
print(ex.example_json("csci6360_project", size=[1024,1024]))
P
B=flame-fractal-renderer_b200/ffr-buf.out
export FFR_TIMING=1
one() { # name, visible
  { time env CUDA_VISIBLE_DEVICES=$2 $B -f /tmp/f.json -o /tmp/o_$1.buf -s 1000000 -b 1000 --seed 3 --no-jit 2> gpurun_out/r2o_$1.txt ; } 2> gpurun_out/r2o_time_$1.txt
  echo "== $1 (visible $2): $(grep real gpurun_out/r2o_time_$1.txt) $(grep -E 'driver initialised|device set up' gpurun_out/r2o_$1.txt | sed 's/timing\[[0-9]*\]: create: //' | tr '\n' ';')"
}
one all8_first 0,1,2,3,4,5,6,7
one all8_again 0,1,2,3,4,5,6,7
one masked_alone 3
one masked_alone2 5
echo "== eight masked processes at once"
for k in 0 1 2 3 4 5 6 7; do one par$k $k & done; wait
echo "== eight unmasked processes at once"
for k in 0 1 2 3 4 5 6 7; do one parall$k 0,1,2,3,4,5,6,7 & done; wait
cat gpurun_out/r2o_smi.txt
