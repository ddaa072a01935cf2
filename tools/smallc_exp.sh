#!/bin/bash
# K3 under the three ways of combining the class splits' dX tiles (ARCFACE_B200_BWD_DXSUM: 0 reduce-add, 1 sum kernel,
# 2 in-kernel sum by the last arrival) at one rank's shard of an 8-GPU job and at the north-star size.
cd "$(dirname "$0")/.."
export ARCFACE_B200_DIAG=1
for C in 125000 1000000; do
for i in 1 2; do
for m in 0 1 2; do
echo -n "C=$C dxsum=$m  "; PROBE_C=$C ARCFACE_B200_BWD_DXSUM=$m python tools/bwd_probe.py 2>&1 | grep "fused default"
done; done; done
