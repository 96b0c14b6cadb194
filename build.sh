#!/bin/bash
# Builds libsggan_sm100.so in-tree (same command __graft_entry__.build() runs).
set -e
cd "$(dirname "$0")/sg-gan-tf2_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
  -o libsggan_sm100.so csrc/*.cu "$@"
