#!/bin/bash
# Timing experiments on the dW epilogue of the single-launch backward (diagnostic variant libraries built with
# make DIAG=1 VARIANT=<name> EXTRA_DEFS=...).  usage: tools/dw_exp.sh "<splits...>" lib1 lib2 ...
cd "$(dirname "$0")/.."
export ARCFACE_B200_DIAG=1
SPLITS="$1"; shift
for lib in "$@"; do
  echo "== lib=$lib"
  ARCFACE_B200_DIAG_LIB=$PWD/multimodalsimilar_b200/libarcface_b200_$lib.so timeout 300 python tools/bwd_probe.py $SPLITS 2>&1 | grep -v Warning
done
