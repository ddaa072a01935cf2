#!/bin/bash
# Fused forward under the measurement modes (0 = normal, 1 = GEMM does not wait for the helper warps, 2 = helper warps
# alone) for a list of diagnostic variant libraries.  usage: tools/fwd_exp.sh lib1 lib2 ...
cd "$(dirname "$0")/.."
export ARCFACE_B200_DIAG=1
for lib in "$@"; do
  for mode in 0 2 1 0; do
    echo -n "lib=$lib "
    ARCFACE_B200_FWD_DEBUG=$mode ARCFACE_B200_DIAG_LIB=$PWD/multimodalsimilar_b200/libarcface_b200_$lib.so timeout 120 python tools/fwd_probe.py 2>&1 | tail -1
  done
done
