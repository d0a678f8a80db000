#!/bin/bash
# build/lib_<name>.so with extra -D flags (A/B variants while tuning; see tools/ab_run.sh)
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -I include -Xptxas -v "$@" \
  onekapy_b200/csrc/oneka_api.cu -o build/lib_$name.so 2>&1 | grep -A2 "Compiling entry function '_Z12track_kernelILb[01]ELi[01]" | grep -v "^--" | paste - - - | sed -e 's/ptxas info    ://g' -e "s/Compiling entry function//" -e 's/Function properties for [^ ]*//' | cut -c1-230
