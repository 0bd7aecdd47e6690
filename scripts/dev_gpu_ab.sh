#!/bin/bash
# A/B of kernel variants inside ONE gpurun call (boxes differ in clocks): each variant twice, interleaved
# a variant = comma-separated VAR=VALUE list
for rep in 1 2; do
for v in "$@"; do
  r=$(env ${v//,/ } timeout 300 python scripts/dev_gpu_fused16.py 8192 ${REGIME:-stationary} 3 2>&1 | tail -2 | head -1 | awk '{print $6, $7, $8}')
  echo "$v rep$rep: $r"
done
done
