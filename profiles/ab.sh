#!/bin/bash
# A/B two builds of the library in ONE gpurun call (boxes differ by ~10%): profiles/ab.sh libA.so libB.so N batch [c]
# the variants live in stablelinearalgebra.jl_b200/csrc/abl/ (git-ignored)
D=stablelinearalgebra.jl_b200/csrc/abl
for rep in 1 2; do
  for v in "$1" "$2"; do
    echo -n "$v: "; SLA_B200_LIB=$PWD/$D/$v SLA_LDR_ALGO=${SLA_LDR_ALGO:-reg} timeout 120 python profiles/ldr_bench.py $3 $4 $5
  done
done
