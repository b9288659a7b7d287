#!/bin/bash
# usage: ab.sh OUT ROUNDS "ENV1" "ENV2" ...   (each ENV is a space-separated list of VAR=val, or "-")
out=$1; rounds=$2; shift 2
: > $out
for r in $(seq $rounds); do
  for e in "$@"; do
    if [ "$e" = "-" ]; then e=""; fi
    env $e python tools/sweep_kernel_ms.py 1024 20 2>&1 | tail -1 >> $out
  done
done
