#!/bin/bash
# A/B of the fused forward's helper-warp configuration (diagnostic libraries built with
# make DIAG=1 VARIANT=<name> EXTRA_DEFS=...); prints ms per launch of tools/fwd_probe.py
cd "$(dirname "$0")/.."
export ARCFACE_B200_DIAG=1
for lib in "$@"; do
  echo -n "lib=$lib: "
  ARCFACE_B200_DIAG_LIB=$PWD/multimodalsimilar_b200/libarcface_b200_$lib.so timeout 120 python tools/fwd_probe.py 2>&1 | tail -1
done
