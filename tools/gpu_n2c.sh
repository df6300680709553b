mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x ) > gpurun_out/n2c_pytest.log 2>&1; tail -2 gpurun_out/n2c_pytest.log
python - <<PY
import importlib, sys
sys.path.insert(0, '.')
ex = importlib.import_module("flame-fractal-renderer_b200.examples")
open('/dev/shm/csci8k.json','w').write(ex.example_json("csci6360_project", size=[8192,8192]))
PY
for g in 1 2; do
( time FFR_TIMING=1 ./flame-fractal-renderer_b200/ffr-buf.out -f /dev/shm/csci8k.json -s 20000000000 -b 8192 --seed 1 --gpus $g --jit -o /dev/shm/csci8k_$g.buf ) > gpurun_out/n2c_cli_$g.log 2>&1; echo "== gpus $g"; tr '\r' '\n' < gpurun_out/n2c_cli_$g.log | grep -a "timing\|render done\|real" | cut -c1-160
done
cmp /dev/shm/csci8k_1.buf /dev/shm/csci8k_2.buf && echo "buffers identical"
