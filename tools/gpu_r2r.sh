#!/bin/bash
# round 2, call R: compact tile: shared-memory directory cache, parity + A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_affine.py tests/test_gpu_fullsize.py -m gpu -x -q -k "not statistical and not size_independent" ) > gpurun_out/r2r_pytest.log 2>&1
tail -3 gpurun_out/r2r_pytest.log
python - <<'P' > gpurun_out/r2r_probe.log 2>&1
import importlib, os, sys, time
sys.path.insert(0,'.')
ffr=importlib.import_module("flame-fractal-renderer_b200")
ex=importlib.import_module("flame-fractal-renderer_b200.examples")
import torch
def run(name,size,blocked,L=8192,waves=4,reps=3):
    os.environ["FFR_DIR_BLOCKED"]=str(blocked)
    fl=ffr.Flame(ex.example_json(name,size=size))
    r=ffr.BufferRenderer(fl,jit=ffr.JIT_ON)
    chains=r.resident_chains*waves
    r.render_chains(0,chains,L,base_seed=1)
    best=0
    for k in range(reps):
        t=time.time(); r.render_chains((k+1)*chains,chains,L,base_seed=1); dt=time.time()-t
        best=max(best,chains*L/dt)
    msg=r.jit_info["message"]
    r.close()
    print("%-24s %-18s blocked=%d  %.3e samples/s  | %s"%(name,size,blocked,best,msg),flush=True)
for blocked, cache in ((1,1),(1,0)):
    os.environ["FFR_DIR_CACHE"]=str(cache)
    run("sierpinski_triangle_3d",[512,512,512],blocked)
    run("barnsley_fern",[8192,8192],blocked)
    run("sierpinski_triangle",[8192,8192],blocked)
    run("sierpinski_triangle_3d",[1024,512,512],blocked)
P
cat gpurun_out/r2r_probe.log
