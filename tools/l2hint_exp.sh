cd /root/repo 2>/dev/null || cd $GRAFT_REPO_ROOT
export ARCFACE_B200_DIAG=1
for i in 1 2; do
for lib in diag diag_l2h; do
for ev in 0 1; do
echo -n "lib=$lib evict_first_dw=$ev: "
ARCFACE_B200_BWD_EVICT=$ev ARCFACE_B200_DIAG_LIB=$PWD/multimodalsimilar_b200/libarcface_b200_$lib.so timeout 200 python tools/bwd_probe.py 2>&1 | grep "fused default"
done; done; done
