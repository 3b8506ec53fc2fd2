#!/bin/bash
# Variant builds of libmoldyn_b200.so for A/B runs on one box (selected with MOLDYN_B200_LIBRARY, see scripts/gpu_round1_final.sh):
#   libmd_r1base.so        the kernels of an earlier commit (default: 380bc4a, the tree before the closing session)
#   libmd_dilute_minb5.so  this tree with 5 resident blocks/SM for the dilute force kernel (96 registers, small spills)
# Outputs are git-ignored (*.so) but travel to the GPU box with the snapshot.
set -e
cd "$(dirname "$0")/.."
V=moldyn_b200/lib/variants
BASE=${1:-380bc4a}
mkdir -p $V
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
env -u CC -u CXX $NVCC $FLAGS -DMD_FORCE_MINB_DILUTE=5 -o $V/libmd_dilute_minb5.so moldyn_b200/csrc/moldyn_b200.cu -ldl
T=$(mktemp -d)
mkdir -p $T/moldyn_b200/csrc $T/include
for f in moldyn_b200/csrc/md_dist.inc moldyn_b200/csrc/md_kernels.cuh moldyn_b200/csrc/moldyn_b200.cu include/moldyn_b200.h; do git show $BASE:$f > $T/$f; done
(cd $T && env -u CC -u CXX $NVCC $FLAGS -o "$OLDPWD/$V/libmd_r1base.so" moldyn_b200/csrc/moldyn_b200.cu -ldl)
rm -rf $T
ls -la $V
