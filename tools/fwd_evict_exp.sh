#!/bin/bash
# Forward with the normalised rows stored under an L2 evict-last policy (diag_wel) against the default: time and DRAM bytes.
cd "$(dirname "$0")/.."
export ARCFACE_B200_DIAG=1
for i in 1 2; do for lib in diag diag_wel; do
echo -n "lib=$lib "; ARCFACE_B200_DIAG_LIB=$PWD/multimodalsimilar_b200/libarcface_b200_$lib.so python tools/fwd_probe.py 2>&1 | tail -1
done; done
for lib in diag diag_wel; do
echo "== ncu lib=$lib"
ARCFACE_B200_DIAG_LIB=$PWD/multimodalsimilar_b200/libarcface_b200_$lib.so ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pair_gemm_kernel --launch-skip 3 -c 2 python tools/fwd_probe.py 2>&1 | grep -E "dram__bytes|gpu__time"
done
