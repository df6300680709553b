timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
REGROUP=1,2 timeout 600 python tools/probe.py csci tkoz3 2>&1 | tail -5
timeout 600 python tools/probe.py sierpinski barnsley sierp3d 2>&1 | tail -5
cd tools
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o ../gpurun_out/prof_csci_regroup3 python prof_one.py csci 2 > /dev/null 2>&1
