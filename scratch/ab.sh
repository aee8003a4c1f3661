#!/bin/bash
# A/B of compile-time variants of the attention kernels: scratch/ab.sh "<flags A>" "<flags B>" ...
for f in "$@"; do
  SPT_NVCC_EXTRA="$f" python -m spt_proto_b200.build --force > /dev/null 2>&1 || { echo "build failed: $f"; continue; }
  echo "== [$f] $(python scratch/attn128_prof.py 2>&1 | grep 'fwd ms')"
done
