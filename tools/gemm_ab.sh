#!/bin/bash
# A/B timings of the per-tap tcgen05 kernel variants on the benchmark shapes (run on the GPU box).
shapes=("4352 1 1 640 1920 1 1 1 7 0" "4352 1 1 640 1280 1 1 1 7 0" "4352 1 1 1280 640 1 1 1 7 1 1" "4352 1 1 640 640 1 1 1 7 1 1" \
        "21760 1 1 128 384 1 1 1 7 0" "21760 1 1 256 128 1 1 1 7 1 1" \
        "256 64 64 64 256 1 1 1 7 1" "256 64 64 64 256 1 1 1 7 0" "256 64 64 256 64 1 1 1 7 0" \
        "256 16 16 128 128 3 1 1 7 1" "256 8 8 256 256 3 1 1 7 1" "256 64 64 256 32 3 1 1 7 0" "256 64 64 32 64 3 2 1 7 0" "256 32 32 64 32 1 1 1 7 0")
for cfg in "1 1" "2 1" "0 1" "0 0"; do
  set -- $cfg
  echo "== CAPF_TC_MSUB=$1 (0 = auto)  CAPF_TC2=$2 (2-CTA kernel for wide Linears)"
  for s in "${shapes[@]}"; do CAPF_TC_MSUB=$1 CAPF_TC2=$2 python tools/one_conv.py $s; done
done
