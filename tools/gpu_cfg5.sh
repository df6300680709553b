mkdir -p gpurun_out
# BASELINE config 5 through the CLI: one process, 8 devices, peer-memory reduce, 1e11 samples at 8192^2
python - <<PY
import importlib, sys
sys.path.insert(0, '.')
ex = importlib.import_module("flame-fractal-renderer_b200.examples")
open('/dev/shm/csci8k.json','w').write(ex.example_json("csci6360_project", size=[8192,8192]))
PY
( time FFR_TIMING=1 ./flame-fractal-renderer_b200/ffr-buf.out -f /dev/shm/csci8k.json -s 100000000000 -b 8192 --seed 1 --gpus 8 --jit -o /dev/shm/csci8k.buf ) > gpurun_out/cfg5_cli.log 2>&1; tr "\r" "\n" < gpurun_out/cfg5_cli.log | grep -a "timing\|render done\|samples plotted\|real" | cut -c1-200
python - <<PY
import numpy as np, hashlib
b = np.fromfile('/dev/shm/csci8k.buf', dtype=np.uint64)
print('cells', b.size, 'sum', int(b.sum()), 'max', int(b.max()), 'nonzero', int((b>0).sum()), 'sha256', hashlib.sha256(b.tobytes()).hexdigest()[:16])
PY
