#!/bin/bash
# A/B of the fused forward's helper-warp run length and L2 hints (diagnostic libraries built with
# make DIAG=1 VARIANT=runN EXTRA_DEFS=-DAB_FWD_RUN=N); prints ms per launch of tools/fwd_probe.py
cd "$(dirname "$0")/.."
export ARCFACE_B200_DIAG=1
for lib in diag diag_run16 diag_run8; do
  for ev in 0 1; do
    for pf in 0; do
      echo -n "lib=$lib evict_first=$ev pf=$pf: "
      ARCFACE_B200_DIAG_LIB=$PWD/multimodalsimilar_b200/libarcface_b200_$lib.so ARCFACE_B200_FWD_EVICT=$ev ARCFACE_B200_FWD_PF=$pf python tools/fwd_probe.py 2>&1 | tail -1
    done
  done
done
