#!/bin/bash
# Single-launch backward with the dX role on CTA pairs (default) against single CTAs (ARCFACE_B200_BWD_DX=cta), per role split.
cd "$(dirname "$0")/.."
export ARCFACE_B200_DIAG=1
for i in 1 2; do
echo "== dX on CTA pairs"; timeout 300 python tools/bwd_probe.py $1 2>&1 | grep -v Warning
echo "== dX on single CTAs"; ARCFACE_B200_BWD_DX=cta timeout 300 python tools/bwd_probe.py $2 2>&1 | grep -v Warning
done
