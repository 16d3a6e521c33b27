#!/bin/sh
# builds tools/ubench (sm_100a micro-benchmarks); run it on a B200: gpurun -- ./tools/ubench
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o ubench ubench.cu
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o ubench_mma ubench_mma.cu
