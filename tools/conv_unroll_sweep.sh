#!/bin/bash
# issue-loop unroll factor sweep for the tcgen05 conv kernels (0 = fully unrolled)
for u in 1 2 4 0; do
  echo "== AVS_CONV_UNROLL=$u"
  AVS_CONV_UNROLL=$u MB_CLIPS=64 MB_QUICK=1 timeout 300 python tools/conv_microbench.py 2>&1 | grep -E "dbg=0"
done
