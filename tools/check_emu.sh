#!/bin/bash
# Rebuild the host emulation of the warp solver and compare it with the oracle (dev loop without a GPU).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
g++ -O2 -std=c++17 -fopenmp -fPIC -shared -Wall -Wno-unknown-pragmas -Wno-maybe-uninitialized \
    -o "$ROOT/tools/emu/libemu.so" "$ROOT/tools/emu/emu.cpp"
python "$ROOT/tools/emu/compare.py" "${1:-8}" "${2:-warp}" | grep -E "^n |agree"
