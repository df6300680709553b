mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/t3_pytest.log 2>&1; tail -6 gpurun_out/t3_pytest.log | cut -c1-300
# CLI: second run of the same flame takes the cached kernel in auto mode
cd flame-fractal-renderer_b200
python - <<'PY'
import importlib, sys
sys.path.insert(0, '..')
ex = importlib.import_module("flame-fractal-renderer_b200.examples")
open('/tmp/fern.json','w').write(ex.example_json("barnsley_fern", size=[2048,2048]))
PY
for i in 1 2 3; do ./ffr-buf.out -f /tmp/fern.json -o /tmp/o$i.buf -s 300000000000 -b 1000000 --seed 1 2>&1 | grep -i "samples/sec\|kernel\|jit" | head -3; done
for i in 1 2; do ./ffr-buf.out -f /tmp/fern.json -o /tmp/p$i.buf -s 20000000000 -b 1000000 --seed 1 2>&1 | grep -i "samples/sec\|kernel\|jit" | head -3; done
cmp /tmp/o1.buf /tmp/o2.buf && echo same
