#!/bin/bash
# Builds liborbx_b200.so (sm_100a only) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")/multi_orbslam3_b200"
SRCS=$(ls csrc/*.cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --threads 0 \
     -Xcompiler -fPIC,-Wall -shared -o liborbx_b200.so $SRCS -lcudart "$@"
echo "built $(pwd)/liborbx_b200.so"
