#!/bin/bash
# Build kernel-experiment variants of libsfw_b200.so into build/variants/ (git-ignored, travels with gpurun):
#   scripts/build_variants.sh name "<extra nvcc flags>" [source root]
set -e
name=$1; flags=$2; root=${3:-/root/repo}
src=$root/social_force_window_planner_b200/csrc
mkdir -p /root/repo/build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --shared -cudart static \
  -diag-suppress 177 $flags $src/sfw_kernels.cu $src/sfw_crowd.cu $src/sfw_abi.cu $src/sfw_sensor.cu $src/sfw_exchange.cu \
  -o /root/repo/build/variants/libsfw_$name.so
echo built $name
