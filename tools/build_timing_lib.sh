#!/bin/bash
# instrumented K3 build (phase timers) next to the product library; used by tools/gpu_timing.sh only
set -e
cd "$(dirname "$0")/../aeroapps_b200/csrc"
make -s
nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -DAB200_BVP_TIMING -c vlidort.cu -o /tmp/vlidort_timing.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libaeroapps_b200_timing.so mie_aop.o /tmp/vlidort_timing.o -lcudart
echo built ../libaeroapps_b200_timing.so
