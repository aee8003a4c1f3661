#!/bin/bash
for f in "$@"; do
  SPT_NVCC_EXTRA="$f" python -m spt_proto_b200.build --force > /dev/null 2>&1 || { echo "build failed: $f"; continue; }
  echo "== [$f] $(python scratch/lk_prof.py 2>&1 | tail -1)"
done
