#!/bin/bash
# round-2 GPU session 2: per-class ncu --set full captures (original mode, B=16) of the conv_tc variants
bash tools/ncu_capture.sh r2a 16 original \
  "128 2 0 1 0 0 55 xf128_d2u1conv1" \
  "128 2 0 1 0 0 62 xf128_u3dense0conv1" \
  "64 2 0 1 0 0 2 xf64_d0u1conv1" \
  "64 2 1 0 1 0 10 rt64_d0u1conv3" \
  "64 2 1 0 1 0 15 rt64_d2u1conv3" \
  "128 3 1 0 0 0 2 res128_d3u1conv3" \
  "32 8 0 0 0 1 36 halo32_u3dense0conv2" \
  "128 3 0 0 0 1 6 halo128_u3conva"
