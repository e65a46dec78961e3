#!/bin/bash
# attention: mbarrier polling (SLIME_ATTN_SPIN) and L2 prefetch distance (SLIME_ATTN_PREFETCH) sweep;
# variant 32 = no softmax (floor of the TMA + MMA side), 5 = column split, 21 = kv split
for cfg in "0 0" "1 0" "0 2" "1 2" "1 4"; do set -- $cfg
  echo "=== spin=$1 prefetch=$2"
  SLIME_ATTN_SPIN=$1 SLIME_ATTN_PREFETCH=$2 AB_VARIANTS=5,21,32 timeout 120 python tools/ab_kernels.py attn 2>&1 | grep -v "peaked\|NVIDIA" | cut -c1-100
done
