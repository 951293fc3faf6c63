#!/bin/bash
# usage: mkvariant.sh NAME SOURCE.cu  -> tools/scratch/libs/libadsel_NAME.so (greedy_warp.o replaced by SOURCE)
set -e
PKG=/root/repo/exploring-diversity-based-active-learning-for-3d-object-detection-in-autonomous-driving_b200
cd $PKG/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I$PKG/csrc -c -o /tmp/variant_$1.o $2 2>&1 | grep -v "warning\|^$\|Remark\|\^\|constexpr" || true
OBJS=$(ls build/*.o | grep -v greedy_warp)
nvcc -shared -o /root/repo/tools/scratch/libs/libadsel_$1.so $OBJS /tmp/variant_$1.o -lcuda 2>&1 | grep -v warning || true
ls -la /root/repo/tools/scratch/libs/libadsel_$1.so
