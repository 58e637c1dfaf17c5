#!/bin/bash
# Build a variant of libvboc_b200.so with extra -D flags into vboc_b200/variants/<name>.so (A/B experiments:
# VBOC_LIB=<path> selects the library at run time).
set -e
name=$1; shift
cd "$(dirname "$0")/../vboc_b200/csrc"
mkdir -p ../variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v -Xcompiler -fPIC -shared -DVB_TUNE_BUILD "$@" \
    -o ../variants/$name.so vboc_cuda.cu > /tmp/vboc_build_$name.log 2>&1 || { grep -E "error" -A2 /tmp/vboc_build_$name.log | head -40; echo BUILD FAILED; exit 1; }
grep -E "solve_kernelILi3ELi0ELi5ELb0" -A2 /tmp/vboc_build_$name.log | grep -E "registers|spill" | tr '\n' ' '; echo
