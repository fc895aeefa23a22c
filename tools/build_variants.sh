#!/bin/bash
# kernel-tuning builds: default, P=640 split, traced
set -e
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared"
nvcc $F -o metasbt_b200/libmsbt.so metasbt_b200/csrc/msbt_kernels.cu &
nvcc $F -DMSBT_FU_P_THREADS=640 -o metasbt_b200/libmsbt_p640.so metasbt_b200/csrc/msbt_kernels.cu &
nvcc $F -DMSBT_FU_TRACE -o metasbt_b200/libmsbt_trace.so metasbt_b200/csrc/msbt_kernels.cu &
wait
