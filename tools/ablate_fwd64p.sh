#!/bin/bash
# Ablation builds of the paired-chunk forward kernel (results are wrong by construction: timing only).
#   build here:  tools/ablate_fwd64p.sh build      run on the GPU box:  tools/ablate_fwd64p.sh run <tag>
# A variant is a '+'-joined list of RTTS_X_* switches (NOEXP NOMASK NOSOFT NOEPI NOSTORE NOCOPY NOSUM NOPV NOS).
V=${VARIANTS:-"NOSUM+NOSOFT+NOEPI NOSOFT+NOEPI NOSUM+NOSOFT+NOEPI+NOPV+NOS NOSUM+NOSOFT+NOEPI+NOPV+NOCOPY NOSUM+NOSOFT+NOEPI+NOPV+NOS+NOCOPY"}
if [ "$1" = build ]; then
  for v in $V; do RTTS_LIB_NAME=libreformer_b200_x$v.so RTTS_DEFS="$(echo $v | sed 's/\([A-Z]*\)/-DRTTS_X_\1/g; s/+/ /g')" python reformer_tts_b200/csrc/build.py 2>/dev/null | tail -1; done
else
  for v in $V; do echo "== $v"; RTTS_LIB=$PWD/reformer_tts_b200/libreformer_b200_x$v.so timeout 100 python tools/ab_fwd.py 2>&1 | grep "kernel=pair" | head -1; done > gpurun_out/$2_ablate.txt 2>&1
  cat gpurun_out/$2_ablate.txt
fi
