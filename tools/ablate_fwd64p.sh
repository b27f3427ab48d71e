#!/bin/bash
# Ablation builds of the paired-chunk forward kernel (results are wrong by construction: timing only).
#   build here:  tools/ablate_fwd64p.sh build      run on the GPU box:  tools/ablate_fwd64p.sh run <tag>
set -e
V="NOEXP NOMASK NOSTORE NOCOPY"
if [ "$1" = build ]; then
  for v in $V; do RTTS_LIB_NAME=libreformer_b200_x$v.so RTTS_DEFS=-DRTTS_X_$v python reformer_tts_b200/csrc/build.py | tail -1; done
else
  for v in $V; do echo "== $v"; RTTS_LIB=$PWD/reformer_tts_b200/libreformer_b200_x$v.so timeout 100 python tools/ab_fwd.py 2>&1 | grep "kernel=pair" | head -2; done > gpurun_out/$2_ablate.txt 2>&1
  cat gpurun_out/$2_ablate.txt
fi
