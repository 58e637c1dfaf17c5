#!/bin/bash
# Build libvboc_b200.so for sm_100a and report registers / spills / code size of the solve kernels.
set -e
cd "$(dirname "$0")/../vboc_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v -Xcompiler -fPIC -shared \
    -o ../libvboc_b200.so vboc_cuda.cu > /tmp/vboc_build.log 2>&1 || { grep -E "error" -A2 /tmp/vboc_build.log | head -40; echo BUILD FAILED; exit 1; }
grep -E "warning|solve_kernelILi3ELi0" -A2 /tmp/vboc_build.log | grep -E "warning|registers|spill" || true
cuobjdump -sass ../libvboc_b200.so | awk '/Function :/ {name=$3} /^ +\/\*[0-9a-f]+\*\/ / {cnt[name]++} END {for (n in cnt) print cnt[n], cnt[n]*16/1024 " KB", n}' | sort -rn | head -2
