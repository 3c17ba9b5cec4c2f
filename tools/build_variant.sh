#!/bin/bash
# developer tool: build libbisip_b200_<tag>.so with extra compile flags for an A/B timing on the GPU box
#   tools/build_variant.sh noexptab -DBISIP_EXP_TABLE=0
# (selected at run time with BISIP_B200_LIB=bisip_b200/csrc/libbisip_b200_<tag>.so; git-ignored, travels with gpurun)
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
C=bisip_b200/csrc; T=$(mktemp -d)
UNITS="api ens_wp_collapsed ens_wp_vec ens_wp_dmma ens_dmma ens_rc ens_umma ens_collapsed ens_colecole ens_dias_shin batch_decomp batch_vec"
for u in $UNITS; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c -o $T/$u.o $C/$u.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $C/libbisip_b200_$tag.so $(for u in $UNITS; do echo $T/$u.o; done)
rm -rf $T
ls -la $C/libbisip_b200_$tag.so
