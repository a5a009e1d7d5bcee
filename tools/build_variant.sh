#!/bin/bash
# tools/build_variant.sh name -DFLAG=..  -> qppvm_b200/variants/libqppvm_b200_name.so (kernel A/B experiments; select with QPPVM_B200_LIB)
name=$1; shift
mkdir -p qppvm_b200/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -ccbin /usr/bin/g++ "$@" \
  -o qppvm_b200/variants/libqppvm_b200_$name.so qppvm_b200/csrc/qppvm_capi.cu qppvm_b200/csrc/qppvm_multi.cu -ldl
