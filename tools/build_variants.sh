#!/bin/bash
# developer experiment: the library built with different arithmetic-form switches (B2G_FUSE_ADD / B2G_FUSE_ACC / B2G_RAW_RSQRT)
cd "$(dirname "$0")/.."
mkdir -p variants
for v in "0 0 0" "0 0 1" "1 0 1" "1 1 1"; do
  set -- $v
  ( nvcc -DB2G_FUSE_ADD=$1 -DB2G_FUSE_ACC=$2 -DB2G_RAW_RSQRT=$3 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -shared -o variants/libb200gym_$1$2$3.so isaacgymenvs_b200/csrc/b200gym.cu ) &
done
wait
ls -la variants/
