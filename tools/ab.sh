#!/bin/bash
# A/B timing of library variants on the bounded profile run: tools/ab.sh <B> <cap> name1 name2 ...
B=$1; cap=$2; shift 2
for v in "$@"; do
  echo "== $v"; VBOC_LIB=$PWD/vboc_b200/variants/$v.so python tools/prof_run.py $B $cap | tail -2
done
