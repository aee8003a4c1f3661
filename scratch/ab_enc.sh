#!/bin/bash
for f in "$@"; do
  SPT_NVCC_EXTRA="$f" python -m spt_proto_b200.build --force > /dev/null 2>&1 || { echo "build failed: $f"; continue; }
  echo "== [$f] $(python scratch/attn_layout_prof.py 2>&1 | sed -n 2p | grep -o 'encode ms [0-9.]*')"
done
