#!/bin/bash
# quick GPU check of the forward attention kernels: A/B timings + the kernel parity tests   (usage: tools/quick_fwd.sh <tag>)
tag=$1
timeout 120 python tools/ab_fwd.py > gpurun_out/${tag}_ab.txt 2>&1; echo "ab rc=$?"; grep -E "kernel=pair|rc=" gpurun_out/${tag}_ab.txt | head -8
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
