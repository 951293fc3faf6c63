#!/bin/bash
# A/B timing of library variants in one session: tools/scratch/ab.sh v0 v1 ...
for rep in 1 2; do
for v in "$@"; do
  echo -n "$v: "; ADSEL_LIB=/root/repo/tools/scratch/libs/libadsel_$v.so timeout 120 python tools/greedy_perf.py P28K 4800 "" 2>&1 | grep -i "variant\|error\|fail" | sed "s/stages.*golden/golden/" | cut -c1-110
done; done
